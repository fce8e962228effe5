"""Development aid (CPU): factor/solve trips of the constraint solver per substep, per lane and per warp (max
over 32 consecutive particles), from the product device math compiled for the host (tests/hostcheck).  The
warp-level trip count is what the GPU pays: one lane that needs a further pass keeps its whole warp in the loop.

usage: python tools/host_trip_stats.py [DEFINE ...]        e.g.  MJB_NO_REPAIR
"""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import reference_noise, synthetic_state                      # noqa: E402
from mjmpc_b200.envs.model import compile_model, reacher7dof_spec          # noqa: E402


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def main():
    defs = sys.argv[1:]
    cm = compile_model(reacher7dof_spec())
    P = cm.chain.params
    so = os.path.join(tempfile.mkdtemp(), "libtrips.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", so,
                           os.path.join(ROOT, "tests", "hostcheck", "hostcheck.cpp"), "-lm"] + ["-D" + d for d in defs])
    lib = C.CDLL(so)
    K, H, seeds = 2048, 32, 4
    tot = np.zeros(6)
    for seed in range(seeds):
        st = synthetic_state(cm, seed)
        noise = reference_noise(K, H, 7, seed + 10)
        mean = np.zeros((H, 7))
        trips = np.zeros(K * H * 2, dtype=np.int32)
        lib.hostcheck_record_trips(trips.ctypes.data_as(C.POINTER(C.c_int)))
        costs = np.zeros((K, H)); qv = np.zeros((K, H, 14))
        lib.hostcheck_rollout(_p(P), 0, _p(st["qp"]), _p(st["qv"]), _p(st["target_pos"]), K, H, _p(mean), _p(noise),
                              _p(costs), _p(qv))
        lib.hostcheck_record_trips(None)
        t = trips.reshape(K, H * 2)                              # particle-major, substep order
        w = t.reshape(K // 32, 32, H * 2).max(axis=1)            # what a warp of 32 consecutive particles pays
        tot += [(t > 1).mean(), t[t > 1].mean(), (t >= 3).mean(), w.mean(), (w >= 3).mean(), (w >= 4).mean()]
    tot /= seeds
    stats = (C.c_longlong * 6)()
    lib.hostcheck_stats(stats, 0)
    print("%-20s constrained substeps %.3f  trips per constrained substep %.3f  lanes with >= 3 trips %.4f"
          % ((",".join(defs) or "default",) + tuple(tot[:3])))
    print("%-20s warp level: mean trips %.3f  >= 3 trips %.3f  >= 4 trips %.3f" % (("",) + tuple(tot[3:])))
    print("%-20s slow-path substeps %d of %d; rank-one repairs tried %d, confirmed %d"
          % ("", stats[3], stats[0], stats[4], stats[5]))


if __name__ == "__main__":
    main()
