"""Controller classes at hypothesis-drawn shapes and options against the numpy restatement of the reference
(oracle/control_np.py, itself pinned to the reference goldens): K down to 2 particles and not a multiple of any
block size, H from 1, d_action 1..8, both input layouts (row-major numpy as an unmodified reference rollout_fn
returns it; particle-minor device tensors as the GPU rollout returns them).  Tolerances as everywhere: sums 1e-10,
elite set / argmin exact.  Sorted last on purpose; derandomised, so the GPU box runs the cases the CPU suite ran."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

pytestmark = pytest.mark.gpu
SET = dict(deadline=None, max_examples=12, derandomize=True, suppress_health_check=list(HealthCheck))
RTOL = 1e-10


def _common(d):
    return dict(d_state=2 * d, d_obs=2 * d, d_action=d, action_lows=-np.ones(d), action_highs=np.ones(d))


def _problem(s, K, H, d, layout):
    import torch
    rng = np.random.RandomState(s)
    costs = np.abs(rng.normal(2.0, 1.0, (K, H)))
    mean = rng.normal(0, 0.3, (H, d))
    actions = mean[None] + rng.normal(0, 1.0, (K, H, d))
    if layout == "numpy":
        traj = dict(costs=costs, actions=actions)
    else:
        traj = dict(costs=torch.from_numpy(costs).cuda().t().contiguous().t(),
                    actions=torch.from_numpy(actions).cuda().permute(1, 2, 0).contiguous().permute(2, 0, 1))
    return costs, mean, actions, traj


shapes = dict(K=st.one_of(st.integers(2, 40), st.integers(2, 3000)), H=st.integers(1, 12), d=st.integers(1, 8),
              layout=st.sampled_from(["numpy", "device"]), s=st.integers(0, 2 ** 31 - 1))


@settings(**SET)
@given(gamma=st.sampled_from([1.0, 0.95]), lam=st.sampled_from([0.1, 1.0]), alpha=st.integers(0, 1), tb=st.booleans(),
       step=st.sampled_from([1.0, 0.6]), base=st.sampled_from(["null", "repeat"]), **shapes)
def test_mppi_class_random_shapes(K, H, d, layout, s, gamma, lam, alpha, tb, step, base):
    from mjmpc_b200.control import MPPI
    from oracle import control_np as O
    costs, mean, actions, traj = _problem(s, K, H, d, layout)
    c = MPPI(horizon=H, init_cov=0.7, base_action=base, lam=lam, num_particles=K, step_size=step, alpha=alpha, gamma=gamma,
             n_iters=1, time_based_weights=tb, filter_coeffs=[0.25, 0.8, 0.0], seed=1, **_common(d))
    c.mean_action = mean
    cov = np.diag([0.7] * d)
    gs = O.gamma_seq(gamma, H)
    want, _ = O.mppi_update(mean, cov, costs, actions, gs, lam, alpha, step, time_based_weights=tb)
    if not tb:
        assert c._calc_val(traj) == pytest.approx(O.mppi_value(mean, cov, costs, actions, gs, lam, alpha), rel=RTOL, abs=1e-12)
    c._update_distribution(traj)
    np.testing.assert_allclose(c.mean_action, want, rtol=RTOL, atol=1e-12)
    c._shift()
    if H == 1 and base == "repeat":
        # the reference indexes mean[-2] and raises for H = 1 (olgaussian_mpc.py:125); the kernel keeps the row
        np.testing.assert_allclose(c.mean_action, want, rtol=RTOL, atol=1e-12)
    else:
        np.testing.assert_allclose(c.mean_action, O.shift_mean(want, base), rtol=RTOL, atol=1e-12)


@settings(**SET)
@given(cov_type=st.sampled_from(["diagonal", "full"]), update_cov=st.booleans(), step=st.sampled_from([1.0, 0.6]), **shapes)
def test_dmd_class_random_shapes(K, H, d, layout, s, cov_type, update_cov, step):
    from mjmpc_b200.control import DMDMPC
    from oracle import control_np as O
    costs, mean, actions, traj = _problem(s, K, H, d, layout)
    c = DMDMPC(horizon=H, init_cov=0.5, beta=0.2, base_action='null', lam=0.3, num_particles=K, step_size=step, gamma=0.99,
               n_iters=1, update_cov=update_cov, cov_type=cov_type, filter_coeffs=[0.25, 0.8, 0.0], seed=1, **_common(d))
    c.mean_action = mean
    cov = np.diag([0.5] * d)
    gs = O.gamma_seq(0.99, H)
    wm, wc, _ = O.dmd_update(mean, cov, costs, actions, gs, 0.3, step, update_cov, cov_type)
    assert c._calc_val(traj) == pytest.approx(O.logsumexp_value(costs, gs, 0.3), rel=RTOL, abs=1e-12)
    c._update_distribution(traj)
    np.testing.assert_allclose(c.mean_action, wm, rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(c.cov_action, wc, rtol=RTOL, atol=1e-12)
    c._shift()
    np.testing.assert_allclose(c.cov_action, wc + (0.2 * np.eye(d) if update_cov else 0.0), rtol=RTOL, atol=1e-12)


@settings(**SET)
@given(cov_type=st.sampled_from(["diagonal", "full"]), frac=st.sampled_from([0.05, 0.2, 0.5, 1.0]), step=st.sampled_from([1.0, 0.7]),
       **shapes)
def test_cem_and_random_shooting_classes_random_shapes(K, H, d, layout, s, cov_type, frac, step):
    from mjmpc_b200.control import CEM, RandomShooting
    from oracle import control_np as O
    costs, mean, actions, traj = _problem(s, K, H, d, layout)
    gs = O.gamma_seq(0.97, H)
    E = int(K * frac)
    if E * H >= 2:                    # np.cov of a single pooled row is undefined in the reference too
        c = CEM(horizon=H, init_cov=0.9, base_action='repeat', elite_frac=frac, num_particles=K, step_size=step, beta=0.1,
                gamma=0.97, n_iters=1, cov_type=cov_type, filter_coeffs=[0.25, 0.8, 0.0], seed=1, **_common(d))
        c.mean_action = mean
        wm, wc, ids = O.cem_update(mean, np.diag([0.9] * d), costs, actions, gs, E, step, cov_type)
        c._update_distribution(traj)
        np.testing.assert_array_equal(np.sort(c.elite_ids.cpu().numpy()), np.sort(ids))      # continuous costs: no ties
        np.testing.assert_allclose(c.mean_action, wm, rtol=RTOL, atol=1e-12)
        np.testing.assert_allclose(c.cov_action, wc, rtol=1e-9, atol=1e-12)
    r = RandomShooting(horizon=H, init_cov=0.9, base_action='null', num_particles=K, step_size=step, gamma=0.97, n_iters=1,
                       filter_coeffs=[0.25, 0.8, 0.0], seed=1, **_common(d))
    r.mean_action = mean
    want = O.rs_update(mean, costs, actions, gs, step)
    r._update_distribution(traj)
    assert int(r.best_id.cpu().numpy().reshape(-1)[0]) == int(np.argmin(O.cost_to_go(costs.copy(), gs)[:, 0]))
    np.testing.assert_allclose(r.mean_action, want if not isinstance(want, tuple) else want[0], rtol=RTOL, atol=1e-12)


@settings(**SET)
@given(lam=st.sampled_from([0.02, 0.3, 2.0]), nsteps=st.integers(0, 5), **shapes)
def test_pfmpc_class_random_shapes(K, H, d, layout, s, lam, nsteps):
    """Weights, the systematic resampler's indices (bit-exact, r drawn from Python's random after
    random.seed(seed_val + num_steps) like the reference), the gathered particle set and its mean."""
    from mjmpc_b200.control import PFMPC
    from oracle import control_np as O
    costs, _, samples, traj = _problem(s, K, H, d, layout)
    c = PFMPC(horizon=H, cov_shift=0.1, cov_resample=1.0, base_action='null', lam=lam, num_particles=K, gamma=0.98, n_iters=1,
              filter_coeffs=[0.25, 0.8, 0.0], seed=7, **_common(d))
    c.action_samples = samples
    c.num_steps = nsteps
    w = O.pf_weights(costs, O.gamma_seq(0.98, H), lam)
    ids, _ = O.pf_resample_indices(w, c.seed_val + nsteps)
    c._update_distribution(dict(costs=traj["costs"]))
    np.testing.assert_array_equal(c.resample_ids.cpu().numpy().reshape(-1), ids % K)
    np.testing.assert_array_equal(c.action_samples, samples[ids])
    np.testing.assert_allclose(c.mean_action, samples[ids].mean(0), rtol=1e-12, atol=1e-14)
