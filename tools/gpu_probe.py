"""Ad-hoc GPU probe: FP64 peak microbenchmark and raw K1 timing (development aid, not the bench)."""
import ctypes as C
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mjmpc_b200 import _lib
from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
from mjmpc_b200.envs.model import compile_model, reacher7dof_spec

L = _lib.lib()
KS = [int(x) for x in os.environ.get("PROBE_K", "8192,65536,262144").split(",")]
for bps, iters in ([(8, 2048)] if "PROBE_K" not in os.environ else []):
    tf = C.c_double(); ms = C.c_double()
    _lib.check(L.mjb_fp64_peak(0, bps, iters, C.byref(tf), C.byref(ms)))
    print("fp64 peak blocks/SM=%d: %.2f TFLOP/s (%.3f ms)" % (bps, tf.value, ms.value))

cm = compile_model(reacher7dof_spec())
env = GpuReacherVecEnv(cm)
rng = np.random.default_rng(0)
lo, hi = cm.tree.jnt_range[:, 0], cm.tree.jnt_range[:, 1]
st = dict(qp=rng.uniform(lo + 0.1 * (hi - lo), hi - 0.1 * (hi - lo)), qv=rng.normal(0, .5, 7), qa=np.zeros(7),
          target_pos=rng.uniform([-.3, -.2, -.25], [.3, .2, .25]), timestep=0)
env.set_env_state(st)
for K in KS:
    H = 32
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    noise = torch.randn((H, 7, K), dtype=torch.float64, device="cuda", generator=g)
    for t in range(2, H):
        noise[t] = 0.25 * noise[t] + 0.8 * noise[t - 1]
    noise = noise.permute(2, 0, 1)
    mean = torch.zeros(H, 7, dtype=torch.float64, device="cuda")
    out = env.rollout_device(K, H, mean, noise, want_ncon=True)
    ncon = out['ncon'].double().mean().item() / (2 * H)
    out = env.rollout_device(K, H, mean, noise)      # production instantiation (no EXTRA outputs)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    ts = []
    for i in range(5):
        e0.record(); env.rollout_device(K, H, mean, noise, costs=out["costs"], actions=out["actions"]); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    ps = K * H / (ms * 1e-3)
    print("K=%d H=%d rollout %.3f ms  %.3e particle-steps/s  alg %.2f TFLOP/s  ncon frac %.3f"
          % (K, H, ms, ps, ps * 5340 / 1e12, ncon))
