"""Secondary measurements for the other BASELINE.json configs (not the driver's bench contract): one JSON
line per config, device-timed MPC steps through the public controller API on one GPU.

  configs[0]  MPPI reacher_7dof, shipped sizes K=32 H=16
  configs[1]  CEM diag-cov, SimplePendulum, K=4096 H=64
  configs[3]  DMD-MPC full covariance, 7-DOF arm, K=65536 H=32
  configs[4]  1024 independent MPPI / PFMPC instances (K=32, H=16) with per-instance randomised dynamics, one launch
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import synthetic_states
from mjmpc_b200.control import CEM, DMDMPC, MPPI, PFMPC
from mjmpc_b200.envs.gpu_pendulum import GpuPendulumVecEnv
from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
from mjmpc_b200.envs.model import compile_model, reacher7dof_spec

compiled = compile_model(reacher7dof_spec())
R7 = dict(d_state=25, d_obs=20, d_action=7, action_lows=-np.ones(7), action_highs=np.ones(7))


WORLD, RANK, LOCAL = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
if WORLD > 1:
    # torchrun: the full-size lines with the particles sharded over the ranks (device time, max over ranks)
    import torch.distributed as dist
    torch.cuda.set_device(LOCAL)
    dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))


def timed(ctrl, states, steps=200, warmup=10, graph=True, set_state=None):
    if graph:
        ctrl.enable_cuda_graph(states[0])
    for i in range(warmup):
        ctrl.step_device(states[i % len(states)])
    torch.cuda.synchronize()
    if WORLD > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        ctrl.step_device(states[(warmup + i) % len(states)])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if WORLD > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def emit(name, ms, K, H, n_ctrl=1, **kw):
    if RANK != 0:
        return
    kw.setdefault("n_gpus", WORLD)
    print(json.dumps(dict(config=name, ms_per_step=ms, mpc_hz=1e3 / ms * n_ctrl, particle_steps_per_s=K * H * n_ctrl / (ms * 1e-3),
                          num_particles=K, horizon=H, instances=n_ctrl, **kw)), flush=True)


states = synthetic_states(compiled, 16, seed=1)
ONLY = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else ""
if WORLD > 1:
    ONLY = "fullsize"
if ONLY in ("", "fullsize"):
    # the other controllers at the north-star size K=65536, H=32: one GPU, eager + graph
    from mjmpc_b200.control import RandomShooting
    K, H = 65536, 32
    from mjmpc_b200.utils.shard import ShardContext
    kw = dict(horizon=H, num_particles=K, gamma=1.0, n_iters=1, filter_coeffs=[0.25, 0.8, 0.0], seed=123,
              shard=ShardContext(RANK, WORLD), device=LOCAL, **R7)
    mk = {
        "MPPI": lambda: MPPI(init_cov=1.0, base_action='null', lam=0.2, step_size=1.0, alpha=1, **kw),
        "CEM full-cov": lambda: CEM(init_cov=1.0, base_action='null', elite_frac=0.2, step_size=1.0, beta=0.0, cov_type='full', **kw),
        "CEM diag-cov": lambda: CEM(init_cov=1.0, base_action='null', elite_frac=0.2, step_size=1.0, beta=0.0, cov_type='diagonal', **kw),
        "RandomShooting": lambda: RandomShooting(init_cov=1.0, base_action='null', step_size=1.0, **kw),
        "PFMPC": lambda: PFMPC(cov_shift=0.05, cov_resample=1.0, base_action='null', lam=0.2, **kw),
        "DMD-MPC diag-cov": lambda: DMDMPC(init_cov=0.1, beta=0.3, base_action='null', lam=0.2, step_size=1.0, update_cov=True,
                                           cov_type='diagonal', **kw),
        "DMD-MPC full-cov": lambda: DMDMPC(init_cov=0.1, beta=0.3, base_action='null', lam=0.2, step_size=1.0, update_cov=True,
                                           cov_type='full', **kw),
    }
    for name, f in mk.items():
        env = GpuReacherVecEnv(compiled, device=LOCAL)
        c = f()
        c.set_sim_state_fn, c.rollout_fn = env.set_env_state, env.rollout_fn
        for graph in (False, True):
            if graph and (name == "PFMPC" or WORLD > 1):
                continue                                 # PFMPC: r comes from Python's random module every step; sharded steps launch eagerly
            try:
                ms = timed(c, states, steps=60, warmup=5, graph=graph)
                emit("fullsize %s K=65536 H=32 (%s)" % (name, "cuda graph" if graph else "eager"), ms, K, H)
            except Exception as e:              # a config that cannot run is a finding, not a crash of the sweep
                print(json.dumps(dict(config="fullsize " + name, graph=graph, error=repr(e)[:300])), flush=True)
        env.close()
        del c, env
        torch.cuda.empty_cache()
    if WORLD > 1:
        dist.barrier()
        dist.destroy_process_group()
    if ONLY:
        sys.exit(0)
# configs[0]
env = GpuReacherVecEnv(compiled)
c = MPPI(horizon=16, init_cov=1.0, base_action='null', lam=0.2, num_particles=32, step_size=1.0, alpha=1, gamma=1.0, n_iters=1,
         filter_coeffs=[0.25, 0.8, 0.0], seed=123, **R7)
c.set_sim_state_fn, c.rollout_fn = env.set_env_state, env.rollout_fn
emit("configs[0] MPPI reacher_7dof shipped sizes", timed(c, states), 32, 16)
env.close()
# configs[1]
penv = GpuPendulumVecEnv()
c = CEM(d_state=2, d_obs=3, d_action=1, horizon=64, init_cov=3.0, base_action='null', elite_frac=0.2, num_particles=4096,
        step_size=1.0, gamma=1.0, n_iters=1, action_lows=penv.action_lows, action_highs=penv.action_highs,
        cov_type='diagonal', filter_coeffs=[0.6, 0.5, 0.0], seed=0)
c.set_sim_state_fn, c.rollout_fn = penv.set_env_state, penv.rollout_fn
emit("configs[1] CEM diag pendulum", timed(c, [{"state": np.array([np.pi, 0.0])}, {"state": np.array([1.0, -0.5])}]), 4096, 64)
# configs[3]
env = GpuReacherVecEnv(compiled)
c = DMDMPC(horizon=32, init_cov=0.1, beta=0.3, base_action='null', lam=0.2, num_particles=65536, step_size=1.0, gamma=1.0,
           n_iters=1, update_cov=True, cov_type='full', filter_coeffs=[0.25, 0.8, 0.0], seed=5, **R7)
c.set_sim_state_fn, c.rollout_fn = env.set_env_state, env.rollout_fn
emit("configs[3] DMD-MPC full covariance 7-DOF arm", timed(c, states, steps=100), 65536, 32)
env.close()
# configs[4]
B = 1024
env = GpuReacherVecEnv(compiled, n_workers=B)
env.randomize_dynamics(dict(body_mass={"r_forearm_link": [0.3, 0.0], "r_wrist_roll_link": [0.3, 0.0]},
                            body_inertia={"r_upper_arm_link": [0.1, 0.1]},
                            dof_damping={"r_elbow_flex_joint": [0.1, 0.1]}), base_seed=123)
c = MPPI(horizon=16, init_cov=1.0, base_action='null', lam=0.2, num_particles=32, step_size=1.0, alpha=1, gamma=1.0, n_iters=1,
         filter_coeffs=[0.25, 0.8, 0.0], seed=123, batch_size=B, **R7)
c.set_sim_state_fn, c.rollout_fn = env.set_env_state, env.rollout_fn
bstates = [synthetic_states(compiled, B, seed=10 + s) for s in range(2)]
emit("configs[4] 1024 independent MPPI instances, randomised dynamics, one GPU", timed(c, bstates, steps=100), 32, 16, n_ctrl=B,
     note="mpc_hz = controller-steps per second summed over the 1024 instances; host->device copy of 1024 states included")
brows = [np.stack([np.concatenate([s["qp"], s["qv"], s["target_pos"]]) for s in bs]) for bs in bstates]
emit("configs[4] 1024 independent MPPI instances, states handed over as one (1024,17) array", timed(c, brows, steps=100), 32, 16,
     n_ctrl=B, note="same step; the per-dict Python work of 1024 state dicts is what the line above mostly measures")
c = PFMPC(horizon=16, cov_shift=0.05, cov_resample=1.0, base_action='null', lam=0.2, num_particles=32, gamma=1.0, n_iters=1,
          filter_coeffs=[0.25, 0.8, 0.0], seed=123, batch_size=B, **R7)
c.set_sim_state_fn, c.rollout_fn = env.set_env_state, env.rollout_fn
emit("configs[4] 1024 independent PFMPC instances, randomised dynamics, one GPU", timed(c, bstates, steps=100, graph=False), 32, 16,
     n_ctrl=B, note="eager launches (the resampler's r comes from Python's random module every step)")
env.close()
