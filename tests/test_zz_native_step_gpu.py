"""mjb_softmax_mpc_step: the whole MPC step of MPPI / DMD-MPC behind one native call must be bit-identical to the
step-by-step path (the same kernels launched one by one from Python, MJB_FUSED_STEP=0) -- actions, mean and
covariance over several hot-started steps, for every option that changes the launch sequence.
(Written without GPU access: sorts after the suites that have run on hardware, so that a surprise here cannot hide
them behind `-x`.)"""
import os
import threading

import numpy as np
import pytest

from conftest import synthetic_state

pytestmark = pytest.mark.gpu

COMMON = dict(d_state=25, d_obs=20, d_action=7, action_lows=-np.ones(7), action_highs=np.ones(7))
CASES = {
    "mppi": ("MPPI", dict(init_cov=1.0, base_action='null', lam=0.2, step_size=1.0, alpha=1, gamma=1.0, n_iters=1)),
    "mppi_ctrlcost_repeat": ("MPPI", dict(init_cov=0.6, base_action='repeat', lam=0.4, step_size=0.8, alpha=0, gamma=0.97, n_iters=1)),
    "mppi_time_based_2iters": ("MPPI", dict(init_cov=0.8, base_action='null', lam=0.3, step_size=0.9, alpha=0, gamma=0.99,
                                            n_iters=2, time_based_weights=True)),
    "mppi_zero_seq_2iters": ("MPPI", dict(init_cov=0.8, base_action='null', lam=0.3, step_size=0.9, alpha=1, gamma=0.99,
                                          n_iters=2, use_zero_control_seq=True)),
    "dmd_full_cov": ("DMDMPC", dict(init_cov=0.5, beta=0.1, base_action='null', lam=0.2, step_size=0.8, gamma=0.98, n_iters=1,
                                    update_cov=True, cov_type='full')),
    "dmd_diag_cov_repeat": ("DMDMPC", dict(init_cov=0.5, beta=0.05, base_action='repeat', lam=0.2, step_size=0.7, gamma=1.0,
                                           n_iters=2, update_cov=True, cov_type='diagonal')),
    "dmd_fixed_cov": ("DMDMPC", dict(init_cov=0.5, beta=0.3, base_action='null', lam=0.2, step_size=1.0, gamma=1.0, n_iters=1)),
}


def _run(compiled_model, name, fused, monkeypatch, K=1024, H=12, steps=4, hotstart=True):
    import mjmpc_b200.control as ctl
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    monkeypatch.setenv("MJB_FUSED_STEP", "1" if fused else "0")
    cls, kw = CASES[name]
    env = GpuReacherVecEnv(compiled_model)
    c = getattr(ctl, cls)(horizon=H, num_particles=K, filter_coeffs=[0.25, 0.8, 0.0], seed=11, **kw, **COMMON)
    c.set_sim_state_fn = env.set_env_state
    c.rollout_fn = env.rollout_fn
    used = []
    real = c._fused_step
    c._fused_step = lambda *a, **k: _tap(real, used, *a, **k)
    acts = [c.optimize(synthetic_state(compiled_model, 70 + s), hotstart=hotstart)[0] for s in range(steps)]
    out = (np.stack(acts), c.mean_action, c.cov_action, c.num_steps, used)
    env.close()
    return out


def _tap(real, used, *a, **k):
    r = real(*a, **k)
    used.append(r is not None)
    return r


@pytest.mark.parametrize("name", sorted(CASES))
def test_native_step_is_bit_identical_to_step_by_step(compiled_model, name, monkeypatch):
    a_f, m_f, c_f, n_f, used_f = _run(compiled_model, name, True, monkeypatch)
    a_s, m_s, c_s, n_s, used_s = _run(compiled_model, name, False, monkeypatch)
    assert all(used_f) and not any(used_s)            # the two runs really took the two paths
    np.testing.assert_array_equal(a_f, a_s)
    np.testing.assert_array_equal(m_f, m_s)
    np.testing.assert_array_equal(c_f, c_s)
    assert n_f == n_s == 4


def test_native_step_without_hotstart_and_fallbacks(compiled_model, monkeypatch):
    """hotstart=False leaves the mean unshifted; options that need host work inside the step (sample_mode='sample',
    base_action='random', a wrapped rollout_fn) silently take the step-by-step path."""
    import mjmpc_b200.control as ctl
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    a_f, m_f, _, _, used = _run(compiled_model, "mppi", True, monkeypatch, steps=2, hotstart=False)
    a_s, m_s, _, _, _ = _run(compiled_model, "mppi", False, monkeypatch, steps=2, hotstart=False)
    assert all(used)
    np.testing.assert_array_equal(a_f, a_s)
    np.testing.assert_array_equal(m_f, m_s)
    monkeypatch.setenv("MJB_FUSED_STEP", "1")
    st = synthetic_state(compiled_model, 3)
    for kw, wrap in ((dict(sample_mode='sample'), False), (dict(base_action='random'), False), (dict(), True)):
        env = GpuReacherVecEnv(compiled_model)
        params = dict(init_cov=1.0, base_action='null', lam=0.2, step_size=1.0, alpha=1, gamma=1.0, n_iters=1)
        params.update(kw)
        c = ctl.MPPI(horizon=8, num_particles=256, filter_coeffs=[0.25, 0.8, 0.0], seed=1, **params, **COMMON)
        c.set_sim_state_fn = env.set_env_state
        fn = env.rollout_fn
        c.rollout_fn = (lambda *a, **k: fn(*a, **k)) if wrap else fn
        assert c._fused_step(st) is None
        action, _ = c.optimize(st)
        assert action.shape == (7,) and np.all(np.isfinite(action))
        env.close()


@pytest.mark.parametrize("cls", ["MPPI", "DMDMPC"])
def test_native_step_with_peer_exchange_two_logical_ranks(compiled_model, cls, monkeypatch):
    """Host emulation only: two logical ranks of one sharded controller run in two threads of this process,
    their symmetric exchange buffers being plain tensors both can address -- the fused exchange+combine kernel
    (stores into the peer's buffer, sequence flags, spin wait) called from mjb_softmax_mpc_step with the
    sequence numbers the controller hands it, n_iters = 2.  Both ranks must end bit-identical and agree with the
    unsharded controller.  On real GPUs this path is covered by tests/test_multigpu_gpu.py."""
    if os.environ.get("MJB_TEST_EMU") != "1":
        pytest.skip("host emulation only (needs peers in one address space)")
    import torch
    import mjmpc_b200.control as ctl
    from mjmpc_b200 import _lib
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    from mjmpc_b200.utils.shard import ShardContext
    monkeypatch.setenv("MJB_FUSED_STEP", "1")
    K, H, W, steps = 512, 8, 2, 3
    kw = dict(init_cov=0.7, base_action='null', lam=0.3, step_size=0.9, gamma=0.98, n_iters=2)
    kw.update(dict(alpha=0) if cls == "MPPI" else dict(beta=0.1, update_cov=True, cov_type='full'))
    states = [synthetic_state(compiled_model, 90 + s) for s in range(steps)]

    def make(shard):
        env = GpuReacherVecEnv(compiled_model)
        c = getattr(ctl, cls)(horizon=H, num_particles=K, filter_coeffs=[0.25, 0.8, 0.0], seed=5, shard=shard, **kw, **COMMON)
        c.set_sim_state_fn = env.set_env_state
        c.rollout_fn = env.rollout_fn
        c.overlap_noise = shard.world_size > 1 and cls == "MPPI"       # the sharded MPPI ranks also prefetch their noise
        return c, env

    ref, env0 = make(ShardContext())
    want = np.stack([ref.optimize(s)[0] for s in states])
    ranks = [make(ShardContext(r, W)) for r in range(W)]
    spec = ranks[0][0]._softmax_spec()
    P = _lib.lib().mjb_softmax_partial_doubles(H, 7, int(spec["time_based"]), spec["cov_mode"])
    bufs = [torch.zeros(2 * W * P + 2 * W, dtype=torch.float64) for _ in range(W)]
    ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64)

    class FakePeerExchange:                       # what utils.shard.PeerExchange provides, without symmetric memory
        def __init__(self):
            self.peer_ptrs_dev, self.seq = ptrs.data_ptr(), 0

        def next_seq(self):
            self.seq += 1
            return self.seq

    got, errs = {}, []

    def work(r):
        try:
            c = ranks[r][0]
            c.__dict__["_px"] = {P: FakePeerExchange()}
            got[r] = (np.stack([c.optimize(s)[0] for s in states]), c.mean_action, c.cov_action)
        except Exception as e:                    # pragma: no cover
            errs.append(repr(e))

    th = [threading.Thread(target=work, args=(r,)) for r in range(W)]
    for t in th:
        t.start()
    for t in th:
        t.join(300)
    assert not errs and len(got) == W, errs
    for k in range(3):
        np.testing.assert_array_equal(got[0][k], got[1][k])            # bit-identical on every rank
    np.testing.assert_allclose(got[0][0], want, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(got[0][1], ref.mean_action, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(got[0][2], ref.cov_action, rtol=1e-9, atol=1e-12)
    for _, e in ranks:
        e.close()
    env0.close()


def test_kernel_launches_per_mppi_step_match_the_bench_claim(compiled_model, monkeypatch):
    """bench.py reports gpu_launches = KERNELS_PER_STEP x steps.  Host emulation only: count the kernel launches of
    one hot-started MPPI step (n_iters = 1): 4 through the native step (noise, rollout, trajectory costs, weighted
    reduction whose last block runs the whole tail), 7 kernel by kernel (noise, rollout, trajectory costs, reduction,
    finalize, combine, shift)."""
    if os.environ.get("MJB_TEST_EMU") != "1":
        pytest.skip("host emulation only (counts launches inside the emulator)")
    import ctypes
    import sys
    from conftest import ROOT
    from mjmpc_b200 import _lib
    sys.path.insert(0, ROOT)
    import bench
    import mjmpc_b200.control as ctl
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    L = _lib.lib()
    L.emu_launch_count.restype = ctypes.c_ulonglong
    for fused in ("1", "0"):
        monkeypatch.setenv("MJB_FUSED_STEP", fused)
        env = GpuReacherVecEnv(compiled_model)
        params = dict(bench.MPPI_PARAMS)
        params.update(horizon=8)
        c = ctl.MPPI(num_particles=256, seed=1, **params, **COMMON)
        c.set_sim_state_fn, c.rollout_fn = env.set_env_state, env.rollout_fn
        st = synthetic_state(compiled_model, 1)
        c.optimize(st)                                   # first step: buffers, model upload
        env.set_env_state(st)
        n0 = L.emu_launch_count()
        c.step_device(None)
        assert L.emu_launch_count() - n0 == (bench.KERNELS_PER_STEP if fused == "1" else 7), (fused, L.emu_launch_count() - n0)
        env.close()


@pytest.mark.parametrize("name", ["mppi", "mppi_ctrlcost_repeat", "mppi_time_based_2iters", "dmd_fixed_cov"])
def test_overlapped_noise_is_bit_identical(compiled_model, name, monkeypatch):
    """overlap_noise: the next step's noise drawn on a side stream during this step's rollout, into the second of
    two noise tensors -- same Philox counters, so actions and distribution must not change by a bit; a jump in
    num_steps (stale prefetch) and a step without hotstart are handled."""
    import mjmpc_b200.control as ctl
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    monkeypatch.setenv("MJB_FUSED_STEP", "1")
    cls, kw = CASES[name]
    outs = []
    for overlap in (False, True):
        env = GpuReacherVecEnv(compiled_model)
        c = getattr(ctl, cls)(horizon=10, num_particles=512, filter_coeffs=[0.25, 0.8, 0.0], seed=11, **kw, **COMMON)
        c.set_sim_state_fn, c.rollout_fn = env.set_env_state, env.rollout_fn
        c.overlap_noise = overlap
        acts = [c.optimize(synthetic_state(compiled_model, 70 + s))[0] for s in range(3)]
        c.num_steps += 5                                              # the prefetched tensor no longer matches
        acts += [c.optimize(synthetic_state(compiled_model, 80 + s), hotstart=(s != 1))[0] for s in range(3)]
        if overlap:
            assert c._fused_blocks["ready"] == (c._fused_blocks["ready"][0], c.num_steps)
        outs.append((np.stack(acts), c.mean_action, c.cov_action))
        env.close()
    for k in range(3):
        np.testing.assert_array_equal(outs[0][k], outs[1][k])


def test_overlapped_noise_is_refused_when_the_noise_depends_on_the_step(compiled_model, monkeypatch):
    """Adapted covariance (DMD-MPC update_cov) or the zero control sequence make the next step's noise a function
    of this step's result: the controller silently keeps the in-line noise kernel."""
    import mjmpc_b200.control as ctl
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    monkeypatch.setenv("MJB_FUSED_STEP", "1")
    for name in ("dmd_full_cov", "mppi_zero_seq_2iters"):
        cls, kw = CASES[name]
        outs = []
        for overlap in (False, True):
            env = GpuReacherVecEnv(compiled_model)
            c = getattr(ctl, cls)(horizon=8, num_particles=256, filter_coeffs=[0.25, 0.8, 0.0], seed=11, **kw, **COMMON)
            c.set_sim_state_fn, c.rollout_fn = env.set_env_state, env.rollout_fn
            c.overlap_noise = overlap
            outs.append(np.stack([c.optimize(synthetic_state(compiled_model, 70 + s))[0] for s in range(3)]))
            assert "ready" not in c._fused_blocks
            env.close()
        np.testing.assert_array_equal(outs[0], outs[1])
