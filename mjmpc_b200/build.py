"""In-tree build of the CUDA extension (``libmjmpc_b200.so``) for sm_100a.

``python -m mjmpc_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import concurrent.futures
import glob
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmjmpc_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "--expt-relaxed-constexpr",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(HERE, "..", "include", "mjmpc_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, out: str = None, defines=()) -> str:
    """out / defines: build a kernel variant next to the default library (experiments; see tools/)."""
    if out is None and not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    target = out or LIB
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else [])
    # one nvcc per translation unit, in parallel (they are independent: no relocatable device code), then one link
    with tempfile.TemporaryDirectory(prefix="mjb_build_") as tmp:
        def compile_one(src):
            obj = os.path.join(tmp, os.path.basename(src)[:-3] + ".o")
            r = subprocess.run([nvcc] + flags + ["-c", "-o", obj, src], capture_output=True, text=True)
            return src, obj, r
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
            results = list(ex.map(compile_one, sources()))
        for src, _, r in results:
            if verbose:
                sys.stderr.write(r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed on %s:\n%s" % (os.path.basename(src), r.stderr))
        r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", target] + [o for _, o, _ in results],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr)
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
