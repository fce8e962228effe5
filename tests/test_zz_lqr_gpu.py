"""K10: the linear-quadratic toy env (reference mjmpc/envs/basic/lqr.py) on the GPU against golden trajectories
recorded from the unmodified reference class (tests/golden/lqr.npz, gen_golden.py) and against the numpy oracle at
other sizes.  numpy's BLAS may order the tiny dot products differently and the device contracts multiply-adds:
1e-11 relative, not bit-exact."""
import numpy as np
import pytest

from golden_util import load

pytestmark = pytest.mark.gpu


def test_lqr_rollout_matches_reference_golden():
    import torch
    from mjmpc_b200.envs.gpu_lqr import GpuLQRVecEnv
    g = load("lqr")
    env = GpuLQRVecEnv(g["A"], g["B"], g["Q"], g["R"])
    env.set_env_state(dict(state=g["state0"]))
    K, H, d = g["noise"].shape
    out = env.rollout_device(K, H, torch.from_numpy(g["mean"]).cuda(), torch.from_numpy(g["noise"]).cuda(), want_states=True)
    np.testing.assert_allclose(out["costs"].cpu().numpy(), g["costs"], rtol=1e-11)
    np.testing.assert_allclose(out["states"].cpu().numpy(), g["states"], rtol=1e-11, atol=1e-12)
    np.testing.assert_array_equal(out["actions"].cpu().numpy(), g["mean"][None] + g["noise"])
    # the reference closure contract: numpy in, numpy out
    traj = env.rollout_fn(K, H, g["mean"], g["noise"], mode="open_loop")
    np.testing.assert_allclose(traj["costs"], g["costs"], rtol=1e-11)
    with pytest.raises(NotImplementedError):
        env.rollout_fn(K, H, g["mean"], g["noise"], mode="closed_loop_linear")


@pytest.mark.parametrize("n,d,K,H", [(1, 1, 300, 5), (8, 8, 77, 3), (3, 1, 1, 1), (2, 5, 130, 9)])
def test_lqr_rollout_matches_oracle_at_other_sizes(n, d, K, H):
    import torch
    from mjmpc_b200.envs.gpu_lqr import GpuLQRVecEnv
    from oracle import lqr_np
    rng = np.random.RandomState(n * 10 + d)
    A = np.eye(n) * 0.9 + 0.1 * rng.normal(0, 1, (n, n)); B = rng.normal(0, 0.5, (n, d))
    Q = rng.normal(0, 1, (n, n)); Q = Q @ Q.T; R = np.diag(rng.uniform(0.1, 1.0, d))
    x0 = rng.normal(0, 1, n)
    mean = rng.normal(0, 0.3, (H, d)); noise = rng.normal(0, 1, (K, H, d))
    ref = lqr_np.rollout(A, B, Q, R, x0, mean, noise)
    env = GpuLQRVecEnv(A, B, Q, R)
    env.set_env_state(dict(state=x0.reshape(n, 1)))
    out = env.rollout_device(K, H, torch.from_numpy(mean).cuda(), torch.from_numpy(noise).cuda(), want_states=True)
    np.testing.assert_allclose(out["costs"].cpu().numpy(), ref["costs"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(out["states"].cpu().numpy(), ref["states"], rtol=1e-10, atol=1e-12)
    # the mean sequence alone (noise=None), and two batched controllers with their own states
    one = env.rollout_device(1, H, torch.from_numpy(mean).cuda(), None)
    ref1 = lqr_np.rollout(A, B, Q, R, x0, mean, np.zeros((1, H, d)))
    np.testing.assert_allclose(one["costs"].cpu().numpy(), ref1["costs"], rtol=1e-10, atol=1e-12)
    if K % 2 == 0:
        env.set_env_state([dict(state=x0), dict(state=-x0)])
        two = env.rollout_device(K, H, torch.from_numpy(np.stack([mean, mean])).cuda(), torch.from_numpy(noise).cuda())
        refb = lqr_np.rollout(A, B, Q, R, -x0, mean, noise[K // 2:])
        np.testing.assert_allclose(two["costs"].cpu().numpy()[K // 2:], refb["costs"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(two["costs"].cpu().numpy()[:K // 2], ref["costs"][:K // 2], rtol=1e-10, atol=1e-12)


def test_lqr_mppi_drives_the_state_to_the_origin():
    """A controller of this package on the LQR backend (the pairing of the reference's softqmpc LQR sanity scripts,
    mjmpc/control/softqmpc/tests/simple_quadratic_model_lqr_test.py:30-35: A = B = Q = 1, R = 0.1)."""
    from mjmpc_b200.control import MPPI
    from mjmpc_b200.envs.gpu_lqr import GpuLQRVecEnv
    from oracle import lqr_np
    A, B, Q, R = np.array([[1.0]]), np.array([[1.0]]), np.array([[1.0]]), np.array([[0.1]])
    env = GpuLQRVecEnv(A, B, Q, R)
    c = MPPI(d_state=1, d_obs=1, d_action=1, horizon=10, init_cov=1.0, base_action='null', lam=0.5, num_particles=512,
             step_size=1.0, alpha=1, gamma=1.0, n_iters=1, action_lows=env.action_lows, action_highs=env.action_highs,
             filter_coeffs=[1.0, 0.0, 0.0], seed=0)
    c.set_sim_state_fn, c.rollout_fn = env.set_env_state, env.rollout_fn
    x = np.array([[5.0]])
    for _ in range(15):
        u, _ = c.optimize(dict(state=x.copy()))
        x = A.dot(x) + B.dot(u.reshape(1, 1))
    assert abs(x.item()) < 0.5
    with pytest.raises(ValueError):
        GpuLQRVecEnv(np.eye(9), np.ones((9, 1)), np.eye(9), np.eye(1))
    assert lqr_np.rollout(A, B, Q, R, [1.0], np.zeros((1, 1)), np.zeros((1, 1, 1)))["costs"][0, 0] == 1.0
