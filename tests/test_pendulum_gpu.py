"""K9 parity: pendulum rollout vs the reference env's own trajectories (golden) and the C oracle."""
import numpy as np
import pytest

from golden_util import load

pytestmark = pytest.mark.gpu


def test_pendulum_matches_reference_golden():
    import torch
    from mjmpc_b200.envs.gpu_pendulum import GpuPendulumVecEnv
    g = load("pendulum")
    env = GpuPendulumVecEnv()
    env.set_env_state({"state": g["state0"]})
    K, H = g["noise"].shape[:2]
    out = env.rollout_device(K, H, torch.from_numpy(g["mean"]).cuda(), torch.from_numpy(g["noise"]).cuda(),
                             want_states=True)
    # only sin() may differ from numpy by an ulp; everything else follows the reference op by op
    np.testing.assert_allclose(out["states"].cpu().numpy(), g["states"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(out["costs"].cpu().numpy(), g["costs"], rtol=1e-12, atol=1e-12)
    np.testing.assert_array_equal(out["actions"].cpu().numpy(), g["mean"][None] + g["noise"])


def test_cem_pendulum_config2_runs_and_improves():
    """configs[1]: CEM diag-cov, K=4096, H=64 on SimplePendulum (simple_pendulum-v0.yml:34-45 hyper-parameters)."""
    from mjmpc_b200.control import CEM
    from mjmpc_b200.envs.gpu_pendulum import GpuPendulumVecEnv
    env = GpuPendulumVecEnv()
    c = CEM(d_state=2, d_obs=3, d_action=1, horizon=64, init_cov=3.0, base_action='null', elite_frac=0.2,
            num_particles=4096, step_size=1.0, gamma=1.0, n_iters=1, action_lows=env.action_lows,
            action_highs=env.action_highs, cov_type='diagonal', filter_coeffs=[0.6, 0.5, 0.0], seed=0)
    c.set_sim_state_fn = env.set_env_state
    c.rollout_fn = env.rollout_fn
    state = np.array([np.pi, 0.0])          # hanging down
    th, thdot = state
    total = 0.0
    for step in range(120):
        a, _ = c.optimize({"state": np.array([th, thdot])})
        u = float(np.clip(a[0], -2, 2))
        x = ((th + np.pi) % (2 * np.pi)) - np.pi
        total += x ** 2 + .1 * thdot ** 2 + .001 * u ** 2
        thdot = thdot + (-15.0 * np.sin(th + np.pi) + 3.0 * u) * .05
        th = th + thdot * .05
        thdot = np.clip(thdot, -8, 8)
    x = ((th + np.pi) % (2 * np.pi)) - np.pi
    # doing nothing (hanging) costs 120 * pi^2 = 1184; the controller must do clearly better than that
    assert np.isfinite(total) and total < 0.8 * 120 * np.pi ** 2, (total, x, thdot)
