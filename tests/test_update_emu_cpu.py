"""The reduction / selection / resampling kernels of update.cu (K3-K8) and their launch code executed on the
host by the block emulator (tests/hostcheck/block_emu.h: CUDA threads as fibers, cooperative barriers, warp
shuffles, real atomics across concurrently running blocks), through the same extern "C" entry points and argument structs as the product library,
against the numpy restatement of the reference (oracle/control_np.py, itself pinned to the reference goldens).
Sizes are small (thousands of OS threads per launch) but span two particle chunks and ragged tails."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from mjmpc_b200 import _lib
from oracle import control_np as O


@pytest.fixture(scope="module")
def L():
    """The whole C ABI built for the host (tests/hostcheck/gen_lib_emu.py); this module uses its update.cu part."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "helpers"))
    import emu_device
    lib = C.CDLL(emu_device.build_lib())
    lib.mjb_last_error.restype = C.c_char_p
    lib.mjb_softmax_scratch_doubles.restype = C.c_longlong
    lib.mjb_elite_scratch_doubles.restype = C.c_longlong
    return lib


def vp(a):
    return C.c_void_p(a.ctypes.data)


def ok(L, rc):
    assert rc == 0, L.mjb_last_error().decode()


def _problem(K, H, d, seed):
    rng = np.random.RandomState(seed)
    costs = np.abs(rng.normal(2.0, 1.0, (K, H)))
    mean = rng.normal(0, 0.3, (H, d))
    actions = mean[None] + rng.normal(0, 1.0, (K, H, d))
    return costs, mean, actions


def _softmax(L, costs, actions, mean, cov, gseq, lam, step_size, control_cost=0, time_based=0, cov_mode=0, n_shards=1,
             wrap=lambda a: a):
    """mjb_softmax_partials per shard + mjb_softmax_combine, like OLGaussianMPC._softmax_update.
    wrap: applied to every array handed to the library (the fuzz tests embed them in guard bands)."""
    K, H, d = actions.shape
    T = H if time_based else 1
    P = L.mjb_softmax_partial_doubles(H, d, time_based, cov_mode)
    g = np.ascontiguousarray(gseq.reshape(-1))
    mean = wrap(mean.copy()); cov = wrap(cov.copy())
    parts = wrap(np.zeros((n_shards, P)))
    kl = K // n_shards
    for r in range(n_shards):
        c = wrap(np.ascontiguousarray(costs[r * kl:(r + 1) * kl])); a_ = wrap(np.ascontiguousarray(actions[r * kl:(r + 1) * kl]))
        a = _lib.SoftmaxArgs()
        a.K, a.H, a.d = kl, H, d
        a.costs = c.ctypes.data; a.costs_sk, a.costs_st = H, 1
        a.actions = a_.ctypes.data; a.act_sk, a.act_st, a.act_sj = H * d, d, 1
        a.mean, a.cov, a.gamma_seq = mean.ctypes.data, cov.ctypes.data, g.ctypes.data
        a.lam, a.control_cost, a.time_based, a.cov_mode = lam, control_cost, time_based, cov_mode
        total = wrap(np.zeros((T, kl))); scratch = wrap(np.zeros(int(L.mjb_softmax_scratch_doubles(kl, H, d, cov_mode))))
        a.total, a.scratch, a.partials = total.ctypes.data, scratch.ctypes.data, parts[r].ctypes.data
        ok(L, L.mjb_softmax_partials(C.byref(a), None))
    stats = wrap(np.zeros(2 + 2 * T))
    cb = _lib.CombineArgs()
    cb.H, cb.d, cb.n_shards, cb.K_global = H, d, n_shards, K
    cb.partials, cb.lam, cb.step_size = parts.ctypes.data, lam, step_size
    cb.time_based, cb.cov_mode = time_based, cov_mode
    cb.mean, cb.cov, cb.stats = mean.ctypes.data, cov.ctypes.data, stats.ctypes.data
    ok(L, L.mjb_softmax_combine(C.byref(cb), None))
    return mean, cov, stats


@pytest.mark.parametrize("gamma", [1.0, 0.9, 0.0])
def test_cost_to_go_bit_exact(L, gamma):
    K, H = 300, 17
    costs = np.abs(np.random.RandomState(1).normal(2, 1, (K, H)))
    gs = O.gamma_seq(gamma, H)
    out = np.zeros((K, H))
    g = np.ascontiguousarray(gs.reshape(-1))
    ok(L, L.mjb_cost_to_go(vp(costs), C.c_longlong(H), C.c_longlong(1), vp(g), K, H, vp(out), C.c_longlong(H), C.c_longlong(1), None))
    np.testing.assert_array_equal(out, O.cost_to_go(costs.copy(), gs))


@pytest.mark.parametrize("alpha,time_based,shards", [(1, 0, 1), (0, 0, 1), (1, 1, 1), (1, 0, 3)])
def test_mppi_update_through_the_emulated_kernels(L, alpha, time_based, shards):
    K, H, d, lam, step = 3000, 6, 7, 0.3, 0.8          # two 2048-particle chunks, ragged tail; 3 shards of 1000
    costs, mean, actions = _problem(K, H, d, 3)
    cov = np.diag(np.linspace(0.5, 1.5, d))
    gs = O.gamma_seq(0.95, H)
    want, w = O.mppi_update(mean, cov, costs, actions, gs, lam, alpha, step, time_based_weights=bool(time_based))
    got, _, stats = _softmax(L, costs, actions, mean, cov, gs, lam, step, control_cost=int(alpha == 0),
                             time_based=time_based, n_shards=shards)
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-12)
    if not time_based:
        np.testing.assert_allclose(stats[0], O.mppi_value(mean, cov, costs, actions, gs, lam, alpha), rtol=1e-10)


@pytest.mark.parametrize("cov_type,mode", [("diagonal", 1), ("full", 2)])
def test_dmd_update_through_the_emulated_kernels(L, cov_type, mode):
    K, H, d, lam, step = 2500, 5, 7, 0.2, 0.6
    costs, mean, actions = _problem(K, H, d, 4)
    cov = np.diag(np.linspace(0.5, 1.5, d))
    gs = O.gamma_seq(1.0, H)
    wm, wc, _ = O.dmd_update(mean, cov, costs, actions, gs, lam, step, True, cov_type)
    gm, gc, stats = _softmax(L, costs, actions, mean, cov, gs, lam, step, cov_mode=mode)
    np.testing.assert_allclose(gm, wm, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(gc, wc, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(stats[0], O.logsumexp_value(costs, gs, lam), rtol=1e-10)


@pytest.mark.parametrize("cov_type", ["diagonal", "full"])
def test_cem_update_through_the_emulated_kernels(L, cov_type):
    K, H, d, step = 2300, 5, 7, 0.7
    costs, mean, actions = _problem(K, H, d, 5)
    costs[7, 0] = costs[3, 0]; costs[7, 1:] = costs[3, 1:]           # an exact tie inside the elite set
    cov = np.diag(np.linspace(0.5, 1.5, d))
    gs = O.gamma_seq(1.0, H)
    E = int(K * 0.2)
    wm, wc, ids = O.cem_update(mean, cov, costs, actions, gs, E, step, cov_type)
    ctg0 = np.ascontiguousarray(O.cost_to_go(costs.copy(), gs)[:, 0])
    flags = np.zeros(K, np.uint8); sel = np.zeros(E, np.int64); scr = np.zeros(4096, np.uint8)
    ok(L, L.mjb_select_elites(vp(ctg0), C.c_longlong(K), C.c_longlong(E), vp(flags), vp(sel), vp(scr), None))
    order = np.lexsort((np.arange(K), ctg0))[:E]                      # ties -> lower index
    np.testing.assert_array_equal(np.sort(sel), np.sort(order))
    np.testing.assert_array_equal(np.flatnonzero(flags), np.sort(order))
    assert set(sel.tolist()) == set(ids.tolist()) or np.isclose(ctg0[sorted(set(ids) ^ set(sel))], ctg0[order[-1]]).all()
    a = _lib.EliteArgs()
    a.K, a.H, a.d = K, H, d
    a.flags, a.actions = flags.ctypes.data, actions.ctypes.data
    a.act_sk, a.act_st, a.act_sj = H * d, d, 1
    m = mean.copy(); cv = cov.copy()
    scratch = np.zeros(int(L.mjb_elite_scratch_doubles(K, H, d)))
    p1 = np.zeros(1 + H * d + d); p2 = np.zeros(d * (d + 1) // 2); mu = np.zeros(d)
    a.mean, a.scratch, a.partial = m.ctypes.data, scratch.ctypes.data, p1.ctypes.data
    ok(L, L.mjb_elite_moments1(C.byref(a), None))
    cb = _lib.EliteCombineArgs()
    cb.H, cb.d, cb.n_shards, cb.full_cov = H, d, 1, int(cov_type == "full")
    cb.partial1, cb.step_size, cb.mu = p1.ctypes.data, step, mu.ctypes.data
    ok(L, L.mjb_elite_combine(C.byref(cb), None))                     # pooled mean of the elite deltas
    a.mu, a.partial = mu.ctypes.data, p2.ctypes.data
    ok(L, L.mjb_elite_moments2(C.byref(a), None))
    cb.partial2, cb.mean, cb.cov = p2.ctypes.data, m.ctypes.data, cv.ctypes.data
    ok(L, L.mjb_elite_combine(C.byref(cb), None))
    np.testing.assert_allclose(m, wm, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(cv, wc, rtol=1e-9, atol=1e-12)


def test_argmin_blend_resample_gather_mean(L):
    K, H, d = 2200, 4, 7
    costs, mean, actions = _problem(K, H, d, 6)
    gs = O.gamma_seq(1.0, H)
    ctg0 = np.ascontiguousarray(O.cost_to_go(costs.copy(), gs)[:, 0])
    ctg0[1500] = ctg0[200] = ctg0.min() - 1.0                         # np.argmin: first occurrence
    idx = np.zeros(1, np.int64); val = np.zeros(1)
    ok(L, L.mjb_argmin(vp(ctg0), C.c_longlong(K), vp(idx), vp(val), None))
    assert idx[0] == 200 == np.argmin(ctg0) and val[0] == ctg0[200]
    # -0.0 and +0.0 compare equal in numpy: the tie goes to the lower index (costs = -rewards makes -0.0 out of 0.0)
    z = np.abs(ctg0) + 1.0; z[40] = 0.0; z[90] = -0.0; z[7] = 0.0
    idx2 = np.zeros(1, np.int64); val2 = np.zeros(1)
    ok(L, L.mjb_argmin(vp(z), C.c_longlong(K), vp(idx2), vp(val2), None))
    assert idx2[0] == 7 == np.argmin(z)
    flags = np.zeros(K, np.uint8); sel = np.zeros(2, np.int64); scr = np.zeros(4096, np.uint8)
    ok(L, L.mjb_select_elites(vp(z), C.c_longlong(K), C.c_longlong(2), vp(flags), vp(sel), vp(scr), None))
    np.testing.assert_array_equal(np.flatnonzero(flags), [7, 40])
    m = mean.copy()
    ok(L, L.mjb_blend_best(vp(actions), C.c_longlong(H * d), C.c_longlong(d), C.c_longlong(1), vp(idx), C.c_longlong(0),
                           K, H, d, C.c_double(0.6), vp(m), None))
    np.testing.assert_allclose(m, 0.4 * mean + 0.6 * actions[200], rtol=1e-15)
    # PFMPC: weights -> systematic resampling (sequential cumulative sum) -> gather -> mean
    w = O.pf_weights(costs, gs, 0.5)
    r = 0.37 / K
    want = O.pf_resample_with_r(w, r)
    cs = np.zeros(K); got = np.zeros(K, np.int64)
    ok(L, L.mjb_resample_indices(vp(w), C.c_longlong(K), C.c_double(r), vp(cs), vp(got), None))
    np.testing.assert_array_equal(got, want)
    out = np.zeros_like(actions)
    s3 = (C.c_longlong(H * d), C.c_longlong(d), C.c_longlong(1))
    ok(L, L.mjb_gather_particles(vp(actions), *s3, vp(got), K, H, d, vp(out), *s3, None))
    np.testing.assert_array_equal(out, actions[want])
    pm = np.zeros((H, d)); scratch = np.zeros(int(L.mjb_elite_scratch_doubles(K, H, d)))
    ok(L, L.mjb_particle_mean(vp(out), *s3, K, H, d, vp(scratch), vp(pm), None))
    np.testing.assert_allclose(pm, out.mean(0), rtol=1e-12)
    sub = np.zeros_like(actions)
    ok(L, L.mjb_particle_sub_mean(vp(out), *s3, vp(pm), K, H, d, vp(sub), *s3, None))
    np.testing.assert_array_equal(sub, out - pm[None])


def test_shifts(L):
    H, d = 6, 7
    rng = np.random.RandomState(2)
    mean = rng.normal(0, 1, (H, d)); row = rng.normal(0, 1, d)
    for name, code in (("null", 0), ("repeat", 1), ("random", 2)):
        m = mean.copy()
        ok(L, L.mjb_shift_mean(vp(m), H, d, code, vp(row), None))
        want = mean.copy(); want[:-1] = mean[1:]
        want[-1] = {"null": np.zeros(d), "repeat": mean[-1], "random": row}[name]
        np.testing.assert_array_equal(m, want)
    cov = np.eye(d) * 0.5; v = np.linspace(1, 2, d)
    ok(L, L.mjb_cov_add_diag(vp(cov), d, C.c_double(0.3), vp(v), None))
    np.testing.assert_allclose(cov, np.eye(d) * 0.5 + 0.3 * np.diag(v), rtol=1e-15)
    assert L.mjb_shift_mean(vp(mean), H, d, 7, None, None) == _lib.MJB_ENOTIMPL


def test_batched_instances_through_the_emulated_kernels(L):
    B, K, H, d, lam = 3, 300, 5, 7, 0.4
    rng = np.random.RandomState(9)
    costs = np.abs(rng.normal(2, 1, (B * K, H))); samples = rng.normal(0, 1, (B * K, H, d))
    means = rng.normal(0, 0.2, (B, H, d)); cov = np.eye(d)
    gs = O.gamma_seq(0.97, H); g = np.ascontiguousarray(gs.reshape(-1))
    a = _lib.MppiBatchedArgs()
    a.n_ctrl, a.K, a.H, a.d = B, K, H, d
    a.costs = costs.ctypes.data; a.costs_sk, a.costs_st = H, 1
    a.actions = samples.ctypes.data; a.act_sk, a.act_st, a.act_sj = H * d, d, 1
    m = means.copy(); val = np.zeros(B)
    a.mean, a.cov, a.gamma_seq, a.lam, a.step_size, a.value = m.ctypes.data, cov.ctypes.data, g.ctypes.data, lam, 0.9, val.ctypes.data
    ok(L, L.mjb_mppi_update_batched(C.byref(a), None))
    for b in range(B):
        sl = slice(b * K, (b + 1) * K)
        want, _ = O.mppi_update(means[b], cov, costs[sl], samples[sl], gs, lam, 1, 0.9)
        np.testing.assert_allclose(m[b], want, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(val[b], O.mppi_value(means[b], cov, costs[sl], samples[sl], gs, lam, 1), rtol=1e-10)
    p = _lib.PfBatchedArgs()
    p.n_ctrl, p.K, p.H, p.d = B, K, H, d
    p.costs = costs.ctypes.data; p.costs_sk, p.costs_st = H, 1
    p.samples = samples.ctypes.data; p.s_sk, p.s_st, p.s_sj = H * d, d, 1
    r = rng.uniform(0, 1.0 / K, B); r[1] = 0.0
    w = np.zeros(B * K); idx = np.zeros(B * K, np.int64); out = np.zeros_like(samples); pmean = np.zeros((B, H, d))
    p.gamma_seq, p.lam, p.r = g.ctypes.data, lam, r.ctypes.data
    p.weights, p.idx, p.out, p.mean = w.ctypes.data, idx.ctypes.data, out.ctypes.data, pmean.ctypes.data
    p.o_sk, p.o_st, p.o_sj = H * d, d, 1
    ok(L, L.mjb_pf_update_batched(C.byref(p), None))
    for b in range(B):
        sl = slice(b * K, (b + 1) * K)
        np.testing.assert_allclose(w[sl], O.pf_weights(costs[sl], gs, lam), rtol=1e-10)
        want = O.pf_resample_with_r(w[sl], r[b]) % K           # r = 0: the reference's index -1 is the last particle
        np.testing.assert_array_equal(idx[sl], want)
        np.testing.assert_array_equal(out[sl], samples[sl][want])
        np.testing.assert_allclose(pmean[b], out[sl].mean(0), rtol=1e-12)
