"""TEST HARNESS (CPU): turns mjmpc_b200/csrc/update.cu into a host translation unit for the block emulator
(block_emu.h) -- kernel<<<grid, block, smem, stream>>>(args) becomes emu::launch(grid, block, smem, stream,
[&]{ kernel(args); }), `extern __shared__ T x[];` becomes a pointer to the emulated dynamic shared memory --
and builds libupdate_emu.so, which exports the same extern "C" entry points as the product library but takes
HOST pointers.  The source text is otherwise untouched: the kernels and their launch code are the product's."""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def transform(src: str) -> str:
    src = re.sub(r"extern\s+__shared__\s+(\w+)\s+(\w+)\[\];", r"\1* \2 = (\1*)emu::g_dyn_smem;", src)
    out, pos = "", 0
    pat = re.compile(r"([A-Za-z_][\w:]*(?:<[^<>;()]*>)?)\s*<<<")
    while True:
        m = pat.search(src, pos)
        if not m:
            out += src[pos:]
            break
        end_cfg = src.index(">>>", m.end())
        cfg = _split_top(src[m.end():end_cfg])
        while len(cfg) < 4:
            cfg.append("0")
        i = src.index("(", end_cfg)
        depth, j = 0, i
        while True:
            depth += src[j] == "("
            depth -= src[j] == ")"
            if depth == 0:
                break
            j += 1
        args = src[i + 1:j].replace("\\\n", " ")
        out += src[pos:m.start()]
        out += "emu::launch(dim3(%s), dim3(%s), (size_t)(%s), (cudaStream_t)(%s), [&] { %s(%s); })" % (
            cfg[0], cfg[1], cfg[2], cfg[3], m.group(1), args)
        pos = j + 1
    return src if not out else out


def build() -> str:
    gen = os.path.join(HERE, "update_emu_gen.cpp")
    so = os.path.join(HERE, "libupdate_emu.so")
    src = open(os.path.join(ROOT, "mjmpc_b200", "csrc", "update.cu")).read()
    body = transform(src).replace('#include "common.h"', '#include "../../mjmpc_b200/csrc/common.h"')
    with open(gen, "w") as f:
        f.write('// GENERATED from mjmpc_b200/csrc/update.cu by gen_update_emu.py -- do not edit\n')
        f.write('#include "block_emu.h"\n#include <stdarg.h>\n#include <stdio.h>\n')
        f.write('#include "../../include/mjmpc_b200.h"\n')
        f.write('namespace mjb { static thread_local char g_err[512]; int set_error(int code, const char* fmt, ...) {\n'
                '    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap); return code; } }\n')
        f.write('extern "C" const char* mjb_last_error(void) { return mjb::g_err; }\n')
        f.write(body)
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-I", cuda_inc,
                           "-o", so, gen])
    return so


if __name__ == "__main__":
    print(build())
