"""CPU: the numpy oracle (oracle/control_np.py) against the reference's own outputs -- the golden
vectors generated from /root/reference (tests/golden/gen_golden.py) and, when the reference tree is
present (this container), the live reference modules."""
import os
import sys

import numpy as np
import pytest

from golden_util import load
from oracle import control_np as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import refload  # noqa: E402


def test_generate_noise_golden():
    g = load("noise")
    eps = O.generate_noise(np.diag([1.0] * 7), [0.25, 0.8, 0.0], (16, 8), int(g["seed"]))
    np.testing.assert_array_equal(eps, g["eps"])


@pytest.mark.parametrize("name", ["ctg_g1", "ctg_g099", "ctg_g05", "ctg_g0"])
def test_cost_to_go_golden(name):
    g = load(name)
    np.testing.assert_array_equal(O.cost_to_go(g["costs"].copy(), g["gamma_seq"]), g["ctg"])


@pytest.mark.parametrize("name", ["mppi_basic", "mppi_ctrlcost", "mppi_timebased", "mppi_tb_ctrlcost"])
def test_mppi_golden(name):
    g = load(name)
    H = g["mean0"].shape[0]
    gs = O.gamma_seq(g["gamma"], H)
    m1, w = O.mppi_update(g["mean0"], g["cov0"], g["costs"], g["actions"], gs, g["lam"], g["alpha"], g["step_size"],
                          bool(g["time_based"]))
    np.testing.assert_array_equal(m1, g["mean1"])
    np.testing.assert_array_equal(w, g["w"])
    if not g["time_based"]:
        assert O.mppi_value(g["mean0"], g["cov0"], g["costs"], g["actions"], gs, g["lam"], g["alpha"]) == g["value"]
    np.testing.assert_array_equal(O.shift_mean(m1, 'null'), g["shifted"])


MPPIQ_CASES = ["mppiq_basic", "mppiq_td", "mppiq_qvals", "mppiq_q_notb", "mppiq_lam0"]


@pytest.mark.parametrize("name", MPPIQ_CASES)
def test_mppiq_golden(name):
    g = load(name)
    q = g["qvals"] if g["qvals"].size else None
    m1, w, q_hat = O.mppiq_update(g["mean0"], g["cov0"], g["costs"], g["actions"], q, g["gamma"], g["td_lam"],
                                  g["beta"], g["alpha"], g["step_size"], bool(g["time_based"]))
    np.testing.assert_array_equal(q_hat, g["q_hat"])
    np.testing.assert_array_equal(w, g["w"])
    np.testing.assert_array_equal(m1, g["mean1"])
    assert O.mppiq_value(g["mean0"], g["cov0"], g["costs"], g["actions"], q, g["gamma"], g["td_lam"], g["beta"],
                         g["alpha"]) == g["value"]


def test_mppiq_td1_equals_discounted_cost_to_go():
    """With td_lam = 1 and no Q estimates the TD(lambda) return telescopes to the discounted cost-to-go."""
    g = load("mppiq_td")
    H = g["costs"].shape[1]
    q_hat = O.mppiq_returns(g["costs"], None, g["gamma"], 1.0, H)
    ctg = O.cost_to_go(g["costs"].copy(), O.gamma_seq(g["gamma"], H))
    np.testing.assert_allclose(q_hat, ctg, rtol=1e-12)


@pytest.mark.parametrize("name", ["cem_diag", "cem_full"])
def test_cem_golden(name):
    g = load(name)
    gs = O.gamma_seq(g["gamma"], g["mean0"].shape[0])
    m1, c1, ids = O.cem_update(g["mean0"], g["cov0"], g["costs"], g["actions"], gs, int(g["num_elite"]), g["step_size"],
                               'full' if g["full"] else 'diagonal')
    np.testing.assert_array_equal(np.sort(ids), g["elite_ids"])
    np.testing.assert_array_equal(m1, g["mean1"])
    np.testing.assert_array_equal(c1, g["cov1"])
    assert O.mean_value(g["costs"], gs) == g["value"]
    np.testing.assert_array_equal(O.shift_mean(m1, 'repeat'), g["shifted"])


@pytest.mark.parametrize("name", ["dmd_nocov", "dmd_diag", "dmd_full"])
def test_dmd_golden(name):
    g = load(name)
    gs = O.gamma_seq(g["gamma"], g["mean0"].shape[0])
    m1, c1, w = O.dmd_update(g["mean0"], g["cov0"], g["costs"], g["actions"], gs, g["lam"], g["step_size"],
                             bool(g["update_cov"]), 'full' if g["full"] else 'diagonal')
    np.testing.assert_array_equal(w, g["w"])
    np.testing.assert_array_equal(m1, g["mean1"])
    np.testing.assert_array_equal(c1, g["cov1"])
    assert O.logsumexp_value(g["costs"], gs, g["lam"]) == g["value"]


def test_random_shooting_golden():
    g = load("rs")
    gs = O.gamma_seq(g["gamma"], g["mean0"].shape[0])
    m1, best = O.rs_update(g["mean0"], g["costs"], g["actions"], gs, g["step_size"])
    assert best == g["best_id"]
    np.testing.assert_array_equal(m1, g["mean1"])


def test_pf_golden():
    g = load("pf")
    gs = O.gamma_seq(1.0, g["costs"].shape[1])
    w = O.pf_weights(g["costs"], gs, g["lam"])
    np.testing.assert_array_equal(w, g["w"])
    ids, r = O.pf_resample_indices(w, int(g["seed"]) + int(g["num_steps"]))
    assert r == g["r"]
    np.testing.assert_array_equal(ids, g["ids"])
    np.testing.assert_array_equal(g["samples0"][ids], g["samples1"])
    g2 = load("pf_skewed")
    ids2, r2 = O.pf_resample_indices(g2["w"], 77)
    assert r2 == g2["r"]
    np.testing.assert_array_equal(ids2, g2["ids"])


def test_pendulum_golden():
    from oracle import mjstep
    g = load("pendulum")
    out = mjstep.pendulum_rollout(g["state0"][0], g["state0"][1], g["mean"], g["noise"][:, :, 0])
    # sin() of glibc (C oracle) vs numpy's: allow an ulp-level drift over the 64-step horizon
    np.testing.assert_allclose(out["states"], g["states"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(out["costs"], g["costs"], rtol=1e-12, atol=1e-12)


@pytest.mark.skipif(not refload.available(), reason="reference tree not mounted (GPU box)")
def test_oracle_matches_live_reference():
    """Fresh random cases through the live reference modules (only where /root/reference exists)."""
    R = refload.load()
    rng = np.random.RandomState(42)
    K, H, d = 96, 12, 7
    common = dict(d_state=25, d_obs=20, action_lows=-np.ones(d), action_highs=np.ones(d))
    for trial in range(3):
        mean = rng.normal(0, 0.3, (H, d))
        actions = mean[None] + rng.normal(0, 1, (K, H, d))
        costs = np.abs(rng.normal(2, 1, (K, H)))
        gamma = [1.0, 0.97, 0.9][trial]
        gs = O.gamma_seq(gamma, H)
        c = R.mppi.MPPI(d_action=d, horizon=H, init_cov=0.7, base_action='null', lam=0.3, num_particles=K,
                        step_size=0.8, alpha=0, gamma=gamma, n_iters=1, **common)
        c.mean_action = mean.copy()
        c._update_distribution(dict(costs=costs.copy(), actions=actions.copy()))
        m1, _ = O.mppi_update(mean, np.diag([0.7] * d), costs, actions, gs, 0.3, 0, 0.8)
        np.testing.assert_array_equal(m1, c.mean_action)
        e = R.cem.CEM(d_action=d, horizon=H, init_cov=1.0, base_action='null', elite_frac=0.25, num_particles=K,
                      step_size=0.5, gamma=gamma, n_iters=1, cov_type='full', **common)
        e.mean_action = mean.copy()
        e._update_distribution(dict(costs=costs.copy(), actions=actions.copy()))
        m2, c2, _ = O.cem_update(mean, np.diag([1.0] * d), costs, actions, gs, e.num_elite, 0.5, 'full')
        np.testing.assert_array_equal(m2, e.mean_action)
        np.testing.assert_array_equal(c2, e.cov_action)


def test_lqr_oracle_matches_reference_golden():
    """oracle/lqr_np.py against trajectories of the unmodified LQREnv (mjmpc/envs/basic/lqr.py)."""
    from golden_util import load
    from oracle import lqr_np
    g = load("lqr")
    out = lqr_np.rollout(g["A"], g["B"], g["Q"], g["R"], g["state0"], g["mean"], g["noise"])
    np.testing.assert_array_equal(out["costs"], g["costs"])
    np.testing.assert_array_equal(out["states"], g["states"])
