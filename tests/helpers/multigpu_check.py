"""Run under torchrun (one rank per GPU): sharded controllers must reproduce the unsharded ones.

Every rank builds (a) a controller that owns only its K/N particle block and exchanges partial vectors
over NCCL, and (b) an unsharded controller on its own GPU.  Philox counters are keyed by the global
particle index, so both see the same noise; after a few MPC steps the distributions must agree (softmax
sums: 1e-10; elite / argmin / resampling indices: exactly) and be identical across ranks."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    emu = os.environ.get("MJB_TEST_EMU") == "1"
    if emu:
        # CPU dry run (tests/test_shard_gloo.py): the same controllers on the host build of the kernels
        # (tests/helpers/emu_device.py), ranks connected by gloo, every rank on "device 0"
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import emu_device
        emu_device.install()
        local = 0
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from mjmpc_b200.control import CEM, DMDMPC, MPPI, PFMPC, RandomShooting
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    from mjmpc_b200.envs.model import compile_model, reacher7dof_spec
    from mjmpc_b200.utils.shard import ShardContext
    compiled = compile_model(reacher7dof_spec())
    shard = ShardContext(rank, world)
    K, H = int(os.environ.get("MJB_CHECK_K", "2048")), 16
    common = dict(d_state=25, d_obs=20, d_action=7, action_lows=-np.ones(7), action_highs=np.ones(7), horizon=H,
                  num_particles=K, gamma=0.98, n_iters=1, filter_coeffs=[0.25, 0.8, 0.0], seed=5, device=local)
    cases = {
        "mppi": (MPPI, dict(init_cov=1.0, base_action='null', lam=0.2, step_size=0.9, alpha=0)),
        "dmd": (DMDMPC, dict(init_cov=0.5, beta=0.1, base_action='null', lam=0.2, step_size=0.8, update_cov=True, cov_type='full')),
        "cem": (CEM, dict(init_cov=1.0, base_action='repeat', elite_frac=0.2, step_size=0.7, beta=0.1, cov_type='full')),
        "rs": (RandomShooting, dict(init_cov=1.0, base_action='null', step_size=1.0)),
        "pfmpc": (PFMPC, dict(cov_shift=0.1, cov_resample=1.0, base_action='null', lam=0.6)),
    }
    state = dict(qp=np.array([0.2, 0.3, -0.1, -1.0, 0.2, -0.4, 0.1]), qv=np.zeros(7), qa=np.zeros(7),
                 target_pos=np.array([0.1, 0.1, 0.1]), timestep=0)
    ok = True
    for name, (cls, kw) in cases.items():
        res = []
        for sh in (shard, ShardContext()):
            env = GpuReacherVecEnv(compiled, device=local)
            kw2 = dict(common); kw2.update(kw)
            if cls is PFMPC:
                kw2.pop("num_particles"); kw2["num_particles"] = K
            c = cls(shard=sh, **kw2)
            c.set_sim_state_fn = env.set_env_state
            c.rollout_fn = env.rollout_fn
            acts = [c.optimize(state)[0] for _ in range(3)]
            extra = None
            if cls is CEM:
                extra = c.elite_ids.cpu().numpy()
            elif cls is RandomShooting:
                extra = c.best_id.cpu().numpy()
            elif cls is PFMPC:
                extra = c.resample_ids.cpu().numpy()
            res.append((np.stack(acts), c.mean_action, getattr(c, "cov_action", None) if cls is not PFMPC else None, extra))
            env.close()
        (a_s, m_s, c_s, e_s), (a_u, m_u, c_u, e_u) = res
        try:
            np.testing.assert_allclose(a_s, a_u, rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(m_s, m_u, rtol=1e-9, atol=1e-12)
            if c_s is not None:
                np.testing.assert_allclose(c_s, c_u, rtol=1e-9, atol=1e-12)
            if e_s is not None:
                np.testing.assert_array_equal(e_s, e_u)
            g = [None] * world
            dist.all_gather_object(g, m_s.tobytes())
            assert all(x == g[0] for x in g), "mean differs across ranks"
            if rank == 0:
                print("multigpu %-6s ok (N=%d)" % (name, world), flush=True)
        except AssertionError as e:
            ok = False
            print("multigpu %-6s FAIL on rank %d: %s" % (name, rank, str(e)[:300]), flush=True)
    # the MJCF-tree backend under the same sharding: MPPI on Swimmer-v0, contacts on
    from mjmpc_b200.envs.gpu_tree_env import GpuTreeVecEnv
    rng = np.random.default_rng(3)
    sw_state = dict(qpos=rng.uniform(-.1, .1, 7), qvel=rng.uniform(-.1, .1, 7))
    res = []
    for sh in (shard, ShardContext()):
        env = GpuTreeVecEnv.swimmer(device=local)
        c = MPPI(d_state=14, d_obs=12, d_action=4, action_lows=env.action_lows, action_highs=env.action_highs, horizon=H,
                 num_particles=K, gamma=0.98, n_iters=1, filter_coeffs=[0.25, 0.8, 0.0], seed=5, device=local, init_cov=0.5,
                 base_action='null', lam=0.1, step_size=0.9, alpha=0, shard=sh)
        c.set_sim_state_fn, c.rollout_fn = env.set_env_state, env.rollout_fn
        res.append((np.stack([c.optimize(sw_state)[0] for _ in range(3)]), c.mean_action))
        env.close()
    try:
        np.testing.assert_allclose(res[0][0], res[1][0], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(res[0][1], res[1][1], rtol=1e-9, atol=1e-12)
        if rank == 0:
            print("multigpu %-6s ok (N=%d)" % ("tree", world), flush=True)
    except AssertionError as e:
        ok = False
        print("multigpu %-6s FAIL on rank %d: %s" % ("tree", rank, str(e)[:300]), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
