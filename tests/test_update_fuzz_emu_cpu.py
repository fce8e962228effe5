"""Randomised shapes through the update kernels (K3-K7, the product's update.cu on the block emulator) against the
numpy restatement of the reference: particle counts that are not multiples of the chunk / block / warp size down
to a single particle, H from 1, every d_action the kernels are instantiated for, ties in the cost-to-go (elite
set and argmin must break them towards the lower index, as numpy's stable order does), weights that underflow to
zero (resampling), shards of ragged size.  Every array handed to the kernels sits between bands of sentinels, so
an out-of-bounds write (or a read that reaches a result) at a ragged tail fails the case.  hypothesis draws the
cases; a fixed seed keeps the suite deterministic."""
import ctypes as C

import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from mjmpc_b200 import _lib
from oracle import control_np as O
from test_update_emu_cpu import L, _softmax, ok, vp   # noqa: F401  (L is the fixture: the emulated library)

class Guard:
    """Embeds every array handed to the kernels in bands of sentinels: an out-of-bounds WRITE destroys a sentinel
    (check()), an out-of-bounds READ of a float array pulls a NaN into a result the test compares."""

    def __init__(self):
        self.bufs = []

    def __call__(self, a, pad=96):
        a = np.ascontiguousarray(a)
        sentinel = np.nan if a.dtype.kind == "f" else np.iinfo(a.dtype).min + 7
        flat = np.full(a.size + 2 * pad, sentinel, dtype=a.dtype)
        view = flat[pad:pad + a.size].reshape(a.shape)
        view[...] = a
        self.bufs.append((flat, pad, a.size, sentinel))
        return view

    def check(self):
        for flat, pad, n, sentinel in self.bufs:
            band = np.concatenate([flat[:pad], flat[pad + n:]])
            assert (np.isnan(band).all() if flat.dtype.kind == "f" else (band == sentinel).all()), "out-of-bounds write"


SETTINGS = dict(deadline=None, max_examples=25, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow],
                derandomize=True)


def _problem(rng, K, H, d, quantum=None):
    costs = np.abs(rng.normal(2.0, 1.0, (K, H)))
    if quantum:
        costs = np.round(costs / quantum) * quantum        # many exact ties
    mean = rng.normal(0, 0.3, (H, d))
    actions = mean[None] + rng.normal(0, 1.0, (K, H, d))
    return costs, mean, actions


@settings(**SETTINGS)
@given(K=st.one_of(st.integers(1, 70), st.integers(1, 4500)), H=st.integers(1, 9), d=st.integers(1, 8),
       gamma=st.sampled_from([1.0, 0.97, 0.5]), lam=st.sampled_from([0.05, 0.3, 2.0]), alpha=st.integers(0, 1),
       time_based=st.integers(0, 1), shards=st.integers(1, 3), s=st.integers(0, 2 ** 31 - 1))
def test_mppi_update_random_shapes(L, K, H, d, gamma, lam, alpha, time_based, shards, s):
    K = max(shards, K - K % shards)
    rng = np.random.RandomState(s)
    costs, mean, actions = _problem(rng, K, H, d)
    cov = np.diag(rng.uniform(0.3, 2.0, d))
    gs = O.gamma_seq(gamma, H)
    want, _ = O.mppi_update(mean, cov, costs, actions, gs, lam, alpha, 0.8, time_based_weights=bool(time_based))
    guard = Guard()
    got, _, stats = _softmax(L, costs, actions, mean, cov, gs, lam, 0.8, control_cost=int(alpha == 0),
                             time_based=time_based, n_shards=shards, wrap=guard)
    guard.check()
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-12)
    if not time_based:
        np.testing.assert_allclose(stats[0], O.mppi_value(mean, cov, costs, actions, gs, lam, alpha), rtol=1e-10, atol=1e-12)


@settings(**SETTINGS)
@given(K=st.one_of(st.integers(1, 70), st.integers(1, 4500)), H=st.integers(1, 9), d=st.integers(1, 8),
       mode=st.integers(0, 2), shards=st.integers(1, 3), s=st.integers(0, 2 ** 31 - 1))
def test_dmd_update_random_shapes(L, K, H, d, mode, shards, s):
    K = max(shards, K - K % shards)
    rng = np.random.RandomState(s)
    costs, mean, actions = _problem(rng, K, H, d)
    cov = np.diag(rng.uniform(0.3, 2.0, d))
    gs = O.gamma_seq(0.99, H)
    wm, wc, _ = O.dmd_update(mean, cov, costs, actions, gs, 0.2, 0.6, mode != 0, "full" if mode == 2 else "diagonal")
    guard = Guard()
    gm, gc, _ = _softmax(L, costs, actions, mean, cov, gs, 0.2, 0.6, cov_mode=mode, n_shards=shards, wrap=guard)
    guard.check()
    np.testing.assert_allclose(gm, wm, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(gc, wc, rtol=1e-10, atol=1e-12)


@settings(**SETTINGS)
@given(K=st.one_of(st.integers(1, 70), st.integers(1, 5000)), frac=st.floats(0.0, 1.0), quantum=st.sampled_from([None, 0.5, 0.05]),
       s=st.integers(0, 2 ** 31 - 1))
def test_elite_set_and_argmin_with_ties(L, K, frac, quantum, s):
    rng = np.random.RandomState(s)
    ctg0 = np.abs(rng.normal(2.0, 1.0, K))
    if quantum:
        ctg0 = np.round(ctg0 / quantum) * quantum
    if K > 3 and s % 3 == 0:
        ctg0[rng.randint(K)] = -ctg0[rng.randint(K)]       # mixed signs: the order-preserving key must handle them
    guard = Guard()
    ctg0 = guard(ctg0)
    E = min(K, max(1, int(K * frac)))
    flags = guard(np.zeros(K, np.uint8)); sel = guard(np.zeros(E, np.int64)); scr = guard(np.zeros(4096, np.uint8))
    ok(L, L.mjb_select_elites(vp(ctg0), C.c_longlong(K), C.c_longlong(E), vp(flags), vp(sel), vp(scr), None))
    order = np.lexsort((np.arange(K), ctg0))[:E]           # == np.argsort(kind="stable")[:E]: ties -> lower index
    np.testing.assert_array_equal(np.flatnonzero(flags), np.sort(order))
    np.testing.assert_array_equal(np.sort(sel), np.sort(order))
    idx = guard(np.zeros(1, np.int64)); val = guard(np.zeros(1))
    ok(L, L.mjb_argmin(vp(ctg0), C.c_longlong(K), vp(idx), vp(val), None))
    guard.check()
    assert idx[0] == np.argmin(ctg0) and val[0] == ctg0.min()


@settings(**SETTINGS)
@given(K=st.one_of(st.integers(1, 70), st.integers(1, 5000)), lam=st.sampled_from([0.002, 0.05, 1.0]), rfrac=st.floats(0.0, 0.999999),
       s=st.integers(0, 2 ** 31 - 1))
def test_systematic_resampling_indices_bit_exact(L, K, lam, rfrac, s):
    rng = np.random.RandomState(s)
    costs = np.abs(rng.normal(2.0, 1.0, (K, 3)))
    w = O.pf_weights(costs, O.gamma_seq(1.0, 3), lam)       # small lam: most weights underflow to exactly 0
    r = rfrac / K
    want = O.pf_resample_with_r(w, r)
    guard = Guard()
    w = guard(w); cs = guard(np.zeros(K + 2)); got = guard(np.zeros(K, np.int64))
    ok(L, L.mjb_resample_indices(vp(w), C.c_longlong(K), C.c_double(r), vp(cs), vp(got), None))
    guard.check()
    np.testing.assert_array_equal(got, want % K)            # the reference's index -1 (r = 0) is the last particle


@settings(**SETTINGS)
@given(K=st.one_of(st.integers(1, 70), st.integers(1, 3000)), H=st.integers(1, 40), gamma=st.sampled_from([1.0, 0.9, 0.3, 0.0]),
       s=st.integers(0, 2 ** 31 - 1))
def test_cost_to_go_bit_exact_random_shapes(L, K, H, gamma, s):
    costs = np.abs(np.random.RandomState(s).normal(2, 1, (K, H)))
    gs = O.gamma_seq(gamma, H)
    guard = Guard()
    costs = guard(costs); out = guard(np.zeros((K, H)))
    g = np.ascontiguousarray(gs.reshape(-1))
    ok(L, L.mjb_cost_to_go(vp(costs), C.c_longlong(H), C.c_longlong(1), vp(g), K, H, vp(out), C.c_longlong(H), C.c_longlong(1), None))
    guard.check()
    np.testing.assert_array_equal(out, O.cost_to_go(costs.copy(), gs))


def _ll(*xs):
    return [C.c_longlong(int(x)) for x in xs]


@settings(**SETTINGS)
@given(K=st.one_of(st.integers(1, 70), st.integers(1, 1500)), H=st.integers(1, 9), d=st.integers(1, 8),
       base=st.integers(0, 2), layout=st.sampled_from(["row", "particle_minor"]), s=st.integers(0, 2 ** 31 - 1))
def test_particle_set_kernels_random_shapes(L, K, H, d, base, layout, s):
    """PFMPC's particle-set kernels (shift + noise, gather, mean, deviation) and RandomShooting's blend at random
    shapes in both memory layouts, against numpy restatements of particle_filter_controller.py:87,102,127-150,168
    and random_shooting.py:61-62."""
    rng = np.random.RandomState(s)
    guard = Guard()

    def alloc(a):       # -> (logical (K,H,d) view, the physical array to point at, element strides), inside guard bands
        if layout == "row":
            phys = guard(a)
            return phys, phys, (H * d, d, 1)
        phys = guard(np.ascontiguousarray(a.transpose(1, 2, 0)))          # (H,d,K)
        return phys.transpose(2, 0, 1), phys, (1, d * K, K)

    samples0 = rng.normal(0, 1, (K, H, d)); delta0 = rng.normal(0, 0.3, (K, H, d)); row = guard(rng.normal(0, 1, d))
    samples, samples_p, ss = alloc(samples0)
    delta, delta_p, ds = alloc(delta0)
    ok(L, L.mjb_pf_shift(vp(samples_p), *_ll(*ss), vp(delta_p), *_ll(*ds), K, H, d, base, vp(row), None))
    want = samples0.copy()
    want[:, :-1] = want[:, 1:]
    want = want + delta0
    if base == 0:
        want[:, -1] = 0.0
    elif base == 1:
        if H >= 2:
            want[:, -1] = want[:, -2]
    else:
        want[:, -1] = row
    np.testing.assert_array_equal(samples, want)
    # gather + mean + deviation
    idx = guard(rng.randint(0, K, K).astype(np.int64))
    out, out_p, os_ = alloc(np.zeros((K, H, d)))
    ok(L, L.mjb_gather_particles(vp(samples_p), *_ll(*ss), vp(idx), K, H, d, vp(out_p), *_ll(*os_), None))
    np.testing.assert_array_equal(out, want[idx])
    pm = guard(np.zeros((H, d))); scratch = guard(np.zeros(int(L.mjb_elite_scratch_doubles(K, H, d))))
    ok(L, L.mjb_particle_mean(vp(out_p), *_ll(*os_), K, H, d, vp(scratch), vp(pm), None))
    np.testing.assert_allclose(pm, want[idx].mean(0), rtol=1e-12, atol=1e-14)
    sub, sub_p, sub_s = alloc(np.zeros((K, H, d)))
    ok(L, L.mjb_particle_sub_mean(vp(out_p), *_ll(*os_), vp(pm), K, H, d, vp(sub_p), *_ll(*sub_s), None))
    np.testing.assert_array_equal(sub, want[idx] - pm[None])
    # RandomShooting: mean <- (1-step)*mean + step*actions[best]
    best = guard(np.array([rng.randint(K)], np.int64)); mean = guard(rng.normal(0, 1, (H, d))); m0 = mean.copy()
    ok(L, L.mjb_blend_best(vp(samples_p), *_ll(*ss), vp(best), C.c_longlong(0), K, H, d, C.c_double(0.7), vp(mean), None))
    np.testing.assert_allclose(mean, (1 - 0.7) * m0 + 0.7 * want[best[0]], rtol=1e-14, atol=1e-15)
    guard.check()


@settings(**SETTINGS)
@given(B=st.integers(1, 6), K=st.one_of(st.integers(1, 40), st.integers(1, 700)), H=st.integers(1, 9), d=st.integers(1, 8),
       alpha=st.integers(0, 1), base=st.integers(0, 1), s=st.integers(0, 2 ** 31 - 1))
def test_batched_instance_kernels_random_shapes(L, B, K, H, d, alpha, base, s):
    """One thread block per controller instance (mjb_mppi_update_batched, mjb_pf_update_batched, the batched shift
    and deviation): every instance must equal the single-controller numpy restatement."""
    rng = np.random.RandomState(s)
    guard = Guard()
    lam = 0.4
    costs = guard(np.abs(rng.normal(2, 1, (B * K, H)))); samples = guard(rng.normal(0, 1, (B * K, H, d)))
    means = rng.normal(0, 0.2, (B, H, d)); cov = guard(np.diag(rng.uniform(0.4, 1.5, d)))
    gs = O.gamma_seq(0.97, H); g = np.ascontiguousarray(gs.reshape(-1))
    a = _lib.MppiBatchedArgs()
    a.n_ctrl, a.K, a.H, a.d = B, K, H, d
    a.costs = costs.ctypes.data; a.costs_sk, a.costs_st = H, 1
    a.actions = samples.ctypes.data; a.act_sk, a.act_st, a.act_sj = H * d, d, 1
    m = guard(means.copy()); val = guard(np.zeros(B))
    a.mean, a.cov, a.gamma_seq, a.lam, a.step_size, a.value = m.ctypes.data, cov.ctypes.data, g.ctypes.data, lam, 0.9, val.ctypes.data
    a.control_cost = int(alpha == 0)
    ok(L, L.mjb_mppi_update_batched(C.byref(a), None))
    for b in range(B):
        sl = slice(b * K, (b + 1) * K)
        want, _ = O.mppi_update(means[b], cov, costs[sl], samples[sl], gs, lam, alpha, 0.9)
        np.testing.assert_allclose(m[b], want, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(val[b], O.mppi_value(means[b], cov, costs[sl], samples[sl], gs, lam, alpha), rtol=1e-10, atol=1e-12)
    m1 = m.copy()
    ok(L, L.mjb_shift_mean_batched(vp(m), B, H, d, base, None, None))
    for b in range(B):
        want = m1[b].copy(); want[:-1] = m1[b][1:]
        want[-1] = 0.0 if base == 0 else (m1[b][-1] if H >= 2 else m1[b][-1])
        np.testing.assert_array_equal(m[b], want)
    p = _lib.PfBatchedArgs()
    p.n_ctrl, p.K, p.H, p.d = B, K, H, d
    p.costs = costs.ctypes.data; p.costs_sk, p.costs_st = H, 1
    p.samples = samples.ctypes.data; p.s_sk, p.s_st, p.s_sj = H * d, d, 1
    r = guard(rng.uniform(0, 1.0 / K, B))
    w = guard(np.zeros(B * K)); idx = guard(np.zeros(B * K, np.int64)); out = guard(np.zeros((B * K, H, d))); pmean = guard(np.zeros((B, H, d)))
    p.gamma_seq, p.lam, p.r = g.ctypes.data, lam, r.ctypes.data
    p.weights, p.idx, p.out, p.mean = w.ctypes.data, idx.ctypes.data, out.ctypes.data, pmean.ctypes.data
    p.o_sk, p.o_st, p.o_sj = H * d, d, 1
    ok(L, L.mjb_pf_update_batched(C.byref(p), None))
    for b in range(B):
        sl = slice(b * K, (b + 1) * K)
        np.testing.assert_allclose(w[sl], O.pf_weights(costs[sl], gs, lam), rtol=1e-10, atol=1e-300)
        want = O.pf_resample_with_r(w[sl], r[b]) % K
        np.testing.assert_array_equal(idx[sl], want)
        np.testing.assert_array_equal(out[sl], samples[sl][want])
        np.testing.assert_allclose(pmean[b], out[sl].mean(0), rtol=1e-12, atol=1e-14)
    sub = guard(np.zeros((B * K, H, d)))
    ok(L, L.mjb_particle_sub_mean_batched(vp(out), *_ll(H * d, d, 1), vp(pmean), B, K, H, d, vp(sub), *_ll(H * d, d, 1), None))
    np.testing.assert_array_equal(sub, out - np.repeat(pmean, K, axis=0))
    guard.check()
