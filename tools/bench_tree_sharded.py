"""Secondary measurement: MPPI on the MJCF-tree backend with the particles SHARDED over the ranks of one node
(torchrun, NCCL, the same ShardContext as the headline bench): K = 65536, H = 32 on HalfCheetah-v0 and Swimmer-v0.
Device-timed MPC steps, max over ranks; rank 0 prints one JSON line per model with the step time, the sharded-vs-
unsharded difference of the first action (rank 0 also runs the unsharded controller once) and the 1-GPU time.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_tree_sharded.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from mjmpc_b200.control import MPPI
from mjmpc_b200.envs.gpu_tree_env import GpuTreeVecEnv
from mjmpc_b200.utils.shard import ShardContext

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
K, H = 65536, 32


def controller(env, shard):
    return MPPI(d_state=env.d_state, d_obs=env.d_obs, d_action=env.d_action, action_lows=env.action_lows,
                action_highs=env.action_highs, horizon=H, init_cov=0.4, base_action="null", num_particles=K, lam=0.1,
                step_size=1.0, alpha=1, gamma=1.0, n_iters=1, set_sim_state_fn=env.set_env_state, rollout_fn=env.rollout_fn,
                seed=3, filter_coeffs=[0.25, 0.8, 0.0], device=local, shard=shard)


def timed(c, state, steps=30, warmup=4):
    for _ in range(warmup):
        c.step_device(state)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        c.step_device(state)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


for name, make in (("HalfCheetah-v0", GpuTreeVecEnv.half_cheetah), ("Swimmer-v0", GpuTreeVecEnv.swimmer)):
    rng = np.random.default_rng(0)
    env = make(device=local)
    state = dict(qpos=rng.uniform(-.1, .1, env.nv), qvel=rng.uniform(-.1, .1, env.nv))
    c = controller(env, ShardContext(rank, world))
    first = c.optimize(state)[0]
    ms = timed(c, state)
    line = dict(config="%s MPPI K=%d H=%d, particles sharded over %d GPU(s)" % (name, K, H, world), n_gpus=world, ms_per_step=ms,
                mpc_hz=1e3 / ms, particle_steps_per_s=K * H / (ms * 1e-3))
    if world > 1:
        dist.barrier()
    if rank == 0 and world > 1:
        env1 = make(device=local)
        c1 = controller(env1, ShardContext())
        ref = c1.optimize(state)[0]
        line["first_action_max_abs_diff_vs_unsharded"] = float(np.abs(first - ref).max())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            c1.step_device(state)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            c1.step_device(state)
        e1.record()
        torch.cuda.synchronize()
        line["ms_per_step_one_gpu_eager"] = e0.elapsed_time(e1) / 10
        line["efficiency"] = line["ms_per_step_one_gpu_eager"] / (world * ms)
        env1.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
    env.close()
if world > 1:
    dist.destroy_process_group()
