"""The controller-level API the north star names besides optimize / get_action: `get_optimal_value`
(reference: mjmpc/control/controller.py:259-275) and `set_params` (not in the reference -- it rebuilds the controller
per episode, examples/example_mpc.py:152-153 -- defined here as: a controller after set_params behaves exactly like a
freshly built one with those parameters)."""
import numpy as np
import pytest

from conftest import synthetic_state

pytestmark = pytest.mark.gpu

COMMON = dict(d_state=25, d_obs=20, d_action=7, action_lows=-np.ones(7), action_highs=np.ones(7))


def _mppi(env, **kw):
    from mjmpc_b200.control import MPPI
    p = dict(horizon=10, num_particles=256, init_cov=0.8, base_action='null', lam=0.3, step_size=0.9, alpha=0, gamma=0.97,
             n_iters=2, filter_coeffs=[0.25, 0.8, 0.0], seed=17)
    p.update(kw)
    c = MPPI(**p, **COMMON)
    c.set_sim_state_fn, c.rollout_fn = env.set_env_state, env.rollout_fn
    return c


@pytest.mark.parametrize("alpha", [0, 1])
def test_get_optimal_value_matches_the_reference_formula(compiled_model, oracle_model, alpha):
    """controller.py:259-275: reset, n_iters x (rollout, update), then ONE more rollout under the updated mean and
    value = -lam * logsumexp(-total / lam, b = 1/K) (mppi.py:113-131).  Restated with the CPU oracle rollout and the
    numpy controller math on the noise the GPU drew (same (seed, step) for every rollout of the call)."""
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    from oracle import control_np as O, mjstep
    env = GpuReacherVecEnv(compiled_model)
    c = _mppi(env, alpha=alpha)
    st = synthetic_state(compiled_model, 8)
    c.optimize(synthetic_state(compiled_model, 9))            # leave the start: get_optimal_value must reset first
    value = c.get_optimal_value(st)
    assert c.num_steps == 1                                   # controller.py:264-271: reset, then one counted call
    K, H, lam = c.num_particles, c.horizon, c.lam
    c2 = _mppi(env, alpha=alpha)
    noise = c2.sample_noise().cpu().numpy()                   # (seed, step 0)
    gs = O.gamma_seq(c.gamma, H)
    mean, cov = np.zeros((H, 7)), np.diag([0.8] * 7)
    roll = lambda m: mjstep.rollout(oracle_model, st["qp"], st["qv"], st["target_pos"], m, np.ascontiguousarray(noise), nthreads=4)
    for _ in range(c.n_iters):
        r = roll(mean)
        mean, _ = O.mppi_update(mean, cov, r["costs"], r["actions"], gs, lam, alpha, c.step_size)
    r = roll(mean)
    want = O.mppi_value(mean, cov, r["costs"], r["actions"], gs, lam, alpha)
    assert value == pytest.approx(want, rel=1e-8)
    np.testing.assert_allclose(c.mean_action, mean, rtol=1e-8, atol=1e-12)      # hotstart=False: not shifted
    env.close()


def test_get_optimal_value_of_the_other_controllers_runs_and_is_finite(compiled_model):
    """CEM / RandomShooting: mean cost-to-go (cem.py:107-113, random_shooting.py:65-69); DMD-MPC: the log-sum-exp
    value (gaussian_dmd.py:126-139); PFMPC raises NotImplementedError exactly like the reference
    (particle_filter_controller.py:176-177)."""
    from mjmpc_b200.control import CEM, DMDMPC, PFMPC, RandomShooting
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    kw = dict(horizon=8, num_particles=128, gamma=0.98, n_iters=1, filter_coeffs=[0.25, 0.8, 0.0], seed=3, **COMMON)
    st = synthetic_state(compiled_model, 2)
    for c in (CEM(init_cov=1.0, base_action='null', elite_frac=0.2, step_size=0.8, beta=0.1, cov_type='full', **kw),
              DMDMPC(init_cov=0.5, beta=0.1, base_action='null', lam=0.2, step_size=0.8, update_cov=False, **kw),
              RandomShooting(init_cov=1.0, base_action='null', step_size=1.0, **kw),
              PFMPC(cov_shift=0.1, cov_resample=1.0, base_action='null', lam=0.5, **kw)):
        env = GpuReacherVecEnv(compiled_model)
        c.set_sim_state_fn, c.rollout_fn = env.set_env_state, env.rollout_fn
        if isinstance(c, PFMPC):
            with pytest.raises(NotImplementedError):
                c.get_optimal_value(st)
        else:
            v = c.get_optimal_value(st)
            assert np.isfinite(v) and 0.0 < v < 1e4, type(c).__name__
        env.close()


@pytest.mark.parametrize("change", [dict(gamma=0.9), dict(horizon=6), dict(seed=99), dict(lam=0.05), dict(num_particles=512),
                                    dict(n_iters=1, step_size=0.5)])
def test_set_params_equals_a_fresh_controller(compiled_model, change):
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    states = [synthetic_state(compiled_model, 40 + s) for s in range(3)]
    env_a, env_b = GpuReacherVecEnv(compiled_model), GpuReacherVecEnv(compiled_model)
    a = _mppi(env_a)
    a.optimize(states[0])                                     # some history, then the change
    a.set_params(**change)
    a.reset()
    b = _mppi(env_b, **change)
    if "gamma" in change or "horizon" in change:
        np.testing.assert_array_equal(a.gamma_seq, b.gamma_seq)
    for st in states:
        np.testing.assert_array_equal(a.optimize(st)[0], b.optimize(st)[0])
    np.testing.assert_array_equal(a.mean_action, b.mean_action)
    assert a.mean_action.shape == (b.horizon, 7)
    env_a.close(); env_b.close()


def test_set_params_invalidates_the_graph_and_rejects_unknown_names(compiled_model):
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    env, env_b = GpuReacherVecEnv(compiled_model), GpuReacherVecEnv(compiled_model)
    st = synthetic_state(compiled_model, 1)
    a = _mppi(env, alpha=1, n_iters=1)
    graphed = a.enable_cuda_graph(st)
    a.optimize(st)
    a.set_params(lam=0.05)
    assert a._graph is None
    a.reset()
    b = _mppi(env_b, alpha=1, n_iters=1, lam=0.05)
    np.testing.assert_array_equal(a.optimize(st)[0], b.optimize(st)[0])
    if graphed:                                               # a new capture picks the new parameter up
        assert a.enable_cuda_graph(st)
        np.testing.assert_array_equal(a.optimize(st)[0], b.optimize(st)[0])
    with pytest.raises(ValueError):
        a.set_params(no_such_parameter=1)
    env.close(); env_b.close()
