"""Secondary measurement (not the driver's bench contract): the tree rollout kernel (K11) and the MPPI
step on the reference's Swimmer-v0 model.  One JSON line per size: kernel time by CUDA events, particle-steps/s
(one particle-step = frame_skip 4 substeps + reward), the MPC step through the controller (CUDA graph), and the CPU
oracle (oracle/tree_step.c, all host threads) on a bounded sample of the same workload.
    python tools/bench_tree.py [--sizes 4096,65536] [--horizon 32]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mjmpc_b200.control import MPPI
from mjmpc_b200.envs import mjcf_tree as T
from mjmpc_b200.envs.gpu_tree_env import GpuTreeVecEnv

sizes = [int(x) for x in (sys.argv[sys.argv.index("--sizes") + 1] if "--sizes" in sys.argv else "4096,65536").split(",")]
H = int(sys.argv[sys.argv.index("--horizon") + 1]) if "--horizon" in sys.argv else 32
rng = np.random.default_rng(0)
MODEL = sys.argv[sys.argv.index("--model") + 1] if "--model" in sys.argv else "swimmer"      # swimmer | swimmer-nocontact | cheetah
env = GpuTreeVecEnv.half_cheetah() if MODEL == "cheetah" else GpuTreeVecEnv.swimmer(contacts=MODEL == "swimmer")
if "--no-planar" in sys.argv:           # the general 3-D instantiation instead of the planar one
    from mjmpc_b200 import _lib
    _lib.lib().mjb_tree_use_planar(0)
INST = "general" if "--no-planar" in sys.argv else "planar"
nv, nu = env.nv, env.d_action
state = dict(qpos=rng.uniform(-.1, .1, nv), qvel=rng.uniform(-.1, .1, nv))
env.set_env_state(state)
for K in sizes:
    mean = torch.as_tensor(rng.normal(0, 0.3, (H, nu)), device=env.device)
    noise = torch.as_tensor(rng.normal(0, 0.5, (K, H, nu)), device=env.device)
    for _ in range(3):
        out = env.rollout_device(K, H, mean, noise)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        out = env.rollout_device(K, H, mean, noise)
    e1.record()
    torch.cuda.synchronize()
    k_ms = e0.elapsed_time(e1) / reps
    c = MPPI(d_state=env.d_state, d_obs=env.d_obs, d_action=nu, action_lows=env.action_lows, action_highs=env.action_highs, horizon=H,
             init_cov=0.4, base_action="null", num_particles=K, lam=0.1, step_size=1.0, alpha=1, gamma=1.0, n_iters=1,
             set_sim_state_fn=env.set_env_state, rollout_fn=env.rollout_fn, seed=3, filter_coeffs=[0.25, 0.8, 0.0])
    c.enable_cuda_graph(state)
    for _ in range(5):
        c.step_device(state)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        c.step_device(state)
    e1.record()
    torch.cuda.synchronize()
    step_ms = e0.elapsed_time(e1) / 50
    t0 = time.perf_counter()
    for _ in range(50):
        c.optimize(state)
    e2e_ms = (time.perf_counter() - t0) / 50 * 1e3
    # CPU oracle on a bounded sample
    from oracle.tree_step import TreeOracle
    o = TreeOracle(env.model, T.solref_to_kb)
    ks = min(K, 2048)
    nz = noise[:ks].cpu().numpy()
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    ref = o.rollout(np.concatenate([state["qpos"], state["qvel"]]), mean.cpu().numpy(), nz, env.frame_skip, env.fwd_dof, env.w_fwd,
                    env.w_ctrl, nthreads=min(cores, 64))
    cpu_s = time.perf_counter() - t0
    err = float(np.abs(out["costs"][:ks].cpu().numpy() - ref["costs"]).max() / (1 + np.abs(ref["costs"]).max()))
    print(json.dumps(dict(config="%s tree rollout + MPPI" % {"cheetah": "HalfCheetah-v0"}.get(MODEL, "Swimmer-v0"), model=MODEL, instantiation=INST, num_particles=K, horizon=H, frame_skip=env.frame_skip,
                          rollout_kernel_ms=k_ms, particle_steps_per_s=K * H / (k_ms * 1e-3), mpc_step_ms=step_ms,
                          mpc_hz=1e3 / step_ms, e2e_ms=e2e_ms, rel_err_vs_oracle=err,
                          cpu_oracle=dict(particle_steps_per_s=ks * H / cpu_s, cores=min(cores, 64), sample="%d particles x %d steps" % (ks, H)))),
          flush=True)
