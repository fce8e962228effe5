"""Randomised shapes through the update kernels (K3-K7, the product's update.cu on the block emulator) against the
numpy restatement of the reference: particle counts that are not multiples of the chunk / block / warp size down
to a single particle, H from 1, every d_action the kernels are instantiated for, ties in the cost-to-go (elite
set and argmin must break them towards the lower index, as numpy's stable order does), weights that underflow to
zero (resampling), shards of ragged size.  Every array handed to the kernels sits between bands of sentinels, so
an out-of-bounds write (or a read that reaches a result) at a ragged tail fails the case.  hypothesis draws the
cases; a fixed seed keeps the suite deterministic."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import HealthCheck, given, seed, settings
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
    w = guard(w); cs = guard(np.zeros(K)); got = guard(np.zeros(K, np.int64))
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
