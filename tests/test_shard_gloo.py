"""CPU, world_size 2, gloo: the host-side sharding protocol of the multi-GPU path.

What runs on the GPUs is tested on the box (`-m gpu`: N logical shards on one device, and the 2-GPU
bench); here two real processes exercise the torch.distributed plumbing the controllers use
(`ShardContext`: contiguous particle blocks, rank-ordered all_gather) and the partial-vector protocol of
the softmax update: every rank reduces ITS particles to [m | S, sum w a] (the layout of
mjb_softmax_partials, here produced by the numpy oracle), the partials are all-gathered and combined in
rank order with the rescaling mjb_softmax_combine applies.  The result must equal the reference's
unsharded MPPI update and be bit-identical on both ranks."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, K, H, d, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mjmpc_b200.utils.shard import ShardContext
        from oracle import control_np as O
        shard = ShardContext.from_env()
        assert (shard.rank, shard.world_size) == (rank, world)
        k0, kl = shard.local_range(K)
        assert (k0, kl) == (rank * K // world, K // world)
        with pytest.raises(AssertionError):
            shard.local_range(K + 1)
        rng = np.random.RandomState(7)            # same data on every rank, each takes its block
        mean = rng.normal(0, 0.3, (H, d))
        actions = mean[None] + rng.normal(0, 1, (K, H, d))
        costs = np.abs(rng.normal(3, 1, (K, H)))
        lam, step, gamma = 0.3, 0.8, 0.97
        gs = O.gamma_seq(gamma, H)
        # phase 1 on this rank's particles
        ctg0 = O.cost_to_go(costs[k0:k0 + kl].copy(), gs)[:, 0]
        m = ctg0.min()
        w = np.exp((-1.0 / lam) * ctg0 - (-1.0 / lam) * m)
        partial = np.concatenate([[m], np.stack([np.concatenate([[w.sum()], (w[:, None] * actions[k0:k0 + kl, t]).sum(0)])
                                                 for t in range(H)]).reshape(-1)])
        allp = shard.all_gather(torch.from_numpy(partial)).numpy()
        assert allp.shape == (world, 1 + H * (1 + d))
        np.testing.assert_array_equal(allp[rank], partial)               # rank order preserved
        # phase 2: rank-ordered combine with the shard-minimum rescaling
        mstar = allp[:, 0].min()
        comb = np.zeros((H, 1 + d))
        for r in range(world):
            comb += allp[r, 1:].reshape(H, 1 + d) * np.exp((-1.0 / lam) * allp[r, 0] - (-1.0 / lam) * mstar)
        new_mean = (1.0 - step) * mean + step * comb[:, 1:] / comb[:, :1]
        want, _ = O.mppi_update(mean, np.eye(d), costs, actions, gs, lam, 1, step)
        np.testing.assert_allclose(new_mean, want, rtol=1e-12, atol=1e-14)
        both = shard.all_gather(torch.from_numpy(new_mean)).numpy()
        np.testing.assert_array_equal(both[0], both[1])                   # identical on every rank
        q.put((rank, "ok"))
    except Exception as e:                                                 # pragma: no cover
        q.put((rank, "FAIL: %r" % (e,)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_softmax_update():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 64, 6, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == {0: "ok", 1: "ok"}, res


def test_two_rank_gloo_sharded_controllers_match_unsharded_on_the_host_emulation():
    """The N > 1 path end to end on CPU: tests/helpers/multigpu_check.py (the script the 2-GPU test launches
    with NCCL) under torchrun with two gloo ranks, the controllers running on the host build of the product's
    kernels (tests/helpers/emu_device.py).  All five controllers: a sharded controller that owns K/2 particles
    and exchanges partials / costs / elite moments / particles must reproduce the unsharded one -- actions,
    mean and covariance to 1e-9, elite, argmin and resampling indices exactly -- and every rank must end with
    bit-identical parameters."""
    import subprocess
    sys.path.insert(0, os.path.join(ROOT, "tests", "helpers"))
    import emu_device
    emu_device.build_lib()
    env = dict(os.environ, MJB_TEST_EMU="1", MJB_P2P="0", MJB_CHECK_K="512", OMP_NUM_THREADS="1")
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "helpers", "multigpu_check.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for name in ("mppi", "dmd", "cem", "rs", "pfmpc", "tree"):
        assert "multigpu %-6s ok" % name in r.stdout, r.stdout[-2000:]


@pytest.mark.parametrize("controller", ["mppi", "pfmpc"])
def test_partitioned_sweep_is_independent_of_the_number_of_ranks(controller):
    """BASELINE configs[4] shape: examples/run_sweep.py (independent MPPI / PFMPC instances with per-instance randomised
    dynamics, partitioned over ranks with no data-path collective) on the host emulation with one rank and with
    two gloo ranks: models, start states and noise are keyed by the global instance index, so the gathered
    results must be bit-identical."""
    import json
    import subprocess
    sys.path.insert(0, os.path.join(ROOT, "tests", "helpers"))
    import emu_device
    emu_device.build_lib()
    run = os.path.join(ROOT, "tests", "helpers", "run_on_emu.py")
    script = [run, "script", os.path.join(ROOT, "examples", "run_sweep.py"), "--instances", "8", "--steps", "3",
              "--controller", controller]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    one = subprocess.run([sys.executable] + script, env=env, capture_output=True, text=True, timeout=600)
    assert one.returncode == 0, one.stdout[-2000:] + one.stderr[-2000:]
    port = 29650 + (os.getpid() % 300)
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port)] + script + ["--backend", "gloo"],
                         env=env, capture_output=True, text=True, timeout=600)
    assert two.returncode == 0, two.stdout[-2000:] + two.stderr[-2000:]
    a = json.loads([l for l in one.stdout.splitlines() if l.startswith("{")][-1])
    b = json.loads([l for l in two.stdout.splitlines() if l.startswith("{")][-1])
    assert (a["n_gpus"], b["n_gpus"], b["instances_per_gpu"]) == (1, 2, 4)
    assert a["result_sha1"] == b["result_sha1"]


def test_single_process_shard_context():
    from mjmpc_b200.utils.shard import ShardContext
    s = ShardContext()
    assert s.local_range(10) == (0, 10)
    t = torch.arange(6.0).reshape(2, 3)
    assert tuple(s.all_gather(t).shape) == (1, 2, 3)
    s4 = ShardContext(3, 4)
    assert s4.local_range(64) == (48, 16)
