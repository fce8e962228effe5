#!/usr/bin/env python
"""A sweep of independent MPPI (or PFMPC) controllers with per-instance randomised dynamics (BASELINE.json configs[4]; the
reference runs such sweeps as separate jobs, examples/job_script.py:186-217, and randomises per worker,
subproc_vec_env.py:304-312), partitioned over the GPUs of one box WITHOUT any collective in the data path:

    python examples/run_sweep.py --instances 1024                                  # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
           examples/run_sweep.py --instances 1024                                  # 128 instances per GPU

Rank r advances instances [r*B/N, (r+1)*B/N) in lock step: one rollout launch over B/N x K particles with one
model per instance, one thread block per instance for the update.  Models, start states and noise are keyed by the
GLOBAL instance index, so every instance behaves the same whatever N is.  The only communication is the timing
reduction (max over ranks) and the gather of the result rows at the end."""
import argparse
import hashlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mjmpc_b200.control import MPPI, PFMPC                           # noqa: E402
from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv             # noqa: E402
from mjmpc_b200.envs.model import compile_model, reacher7dof_spec    # noqa: E402

RANDOMIZE = dict(body_mass={"r_forearm_link": [0.3, 0.0], "r_wrist_roll_link": [0.3, 0.0]},
                 body_inertia={"r_upper_arm_link": [0.1, 0.1]}, dof_damping={"r_elbow_flex_joint": [0.1, 0.1]})


def start_state(compiled, instance, step):
    """Synthetic start state of one instance, a function of the global instance index."""
    rng = np.random.default_rng([instance, step])
    lo, hi = compiled.tree.jnt_range[:, 0], compiled.tree.jnt_range[:, 1]
    return np.concatenate([rng.uniform(lo + 0.2 * (hi - lo), hi - 0.2 * (hi - lo)), rng.normal(0, 0.5, 7),
                           rng.uniform([-.3, -.2, -.25], [.3, .2, .25])])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--instances", type=int, default=1024)
    ap.add_argument("--particles", type=int, default=32)       # reacher_7dof-v0.yml: 8 workers x 4 particles
    ap.add_argument("--horizon", type=int, default=16)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--seed", type=int, default=123)
    ap.add_argument("--controller", default="mppi", choices=["mppi", "pfmpc"])
    ap.add_argument("--backend", default="nccl", help="torch.distributed backend when launched with torchrun")
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group(args.backend)
    if args.instances % world != 0:
        raise AssertionError("Number of instances must be divisible by number of GPUs")
    B = args.instances // world
    first = rank * B
    compiled = compile_model(reacher7dof_spec())
    env = GpuReacherVecEnv(compiled, n_workers=B, device=local)
    env.randomize_dynamics(RANDOMIZE, base_seed=args.seed, worker_offset=first)
    common = dict(d_state=env.d_state, d_obs=env.d_obs, d_action=7, action_lows=env.action_lows, action_highs=env.action_highs,
                  horizon=args.horizon, base_action='null', num_particles=args.particles, gamma=1.0, n_iters=1,
                  filter_coeffs=[0.25, 0.8, 0.0], seed=args.seed, batch_size=B, device=local)
    if args.controller == "mppi":
        ctrl = MPPI(init_cov=1.0, lam=0.2, step_size=1.0, alpha=1, **common)
    else:
        ctrl = PFMPC(cov_shift=0.05, cov_resample=1.0, lam=0.2, **common)
    ctrl.set_instance_offset(first)
    ctrl.set_sim_state_fn, ctrl.rollout_fn = env.set_env_state, env.rollout_fn
    states = [torch.from_numpy(np.stack([start_state(compiled, first + b, s) for b in range(B)])).cuda() for s in range(4)]
    ctrl.enable_cuda_graph(states[0])
    for s in range(3):
        ctrl.step_device(states[s % 4])
    ctrl.reset()
    ctrl.set_instance_offset(first)
    ctrl.enable_cuda_graph(states[0])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    actions = None
    for s in range(args.steps):
        actions = ctrl.step_device(states[s % 4])
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    rows = np.concatenate([actions.cpu().numpy().reshape(B, 7), ctrl.mean_action.reshape(B, -1)], axis=1)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        gathered = [None] * world
        dist.all_gather_object(gathered, rows)
        rows = np.concatenate(gathered, axis=0)
    if rank == 0:
        t = float(ms.item()) * 1e-3
        print(json.dumps(dict(config="independent %s instances, per-instance randomised dynamics, no collective" % args.controller.upper(),
                              instances=args.instances, n_gpus=world, instances_per_gpu=B, num_particles=args.particles,
                              horizon=args.horizon, steps=args.steps, ms_per_sweep_step=t / args.steps * 1e3,
                              controller_steps_per_s=args.instances * args.steps / t,
                              particle_steps_per_s=args.instances * args.particles * args.horizon * args.steps / t,
                              result_sha1=hashlib.sha1(np.ascontiguousarray(rows).tobytes()).hexdigest())))
    if world > 1:
        dist.destroy_process_group()
    env.close()


if __name__ == "__main__":
    main()
