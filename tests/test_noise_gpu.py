"""K2: Philox noise kernel -- statistics, filter, covariance factor, determinism, shard independence."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_moments_and_filter():
    from mjmpc_b200.utils.control_utils import generate_noise
    K, H, d = 200000, 12, 7
    b = (0.25, 0.8, 0.1)
    cov = np.diag([1.0, 0.5, 2.0, 1.0, 0.1, 3.0, 1.0])
    eps = generate_noise(cov, b, (K, H), 1234, step=5).cpu().numpy()
    assert eps.shape == (K, H, d)
    # steps 0,1 are unfiltered N(0, cov)
    for t in (0, 1):
        assert np.all(np.abs(eps[:, t].mean(0)) < 5 * np.sqrt(np.diag(cov) / K))
        np.testing.assert_allclose(eps[:, t].var(0), np.diag(cov), rtol=0.02)
    # un-filter and recover white noise: z_i = (eps_i - b1 eps_{i-1} - b2 eps_{i-2}) / b0
    z = (eps[:, 2:] - b[1] * eps[:, 1:-1] - b[2] * eps[:, :-2]) / b[0]
    np.testing.assert_allclose(z.var(axis=(0, 1)), np.diag(cov), rtol=0.02)
    zc = z[:, :, 0]
    assert abs(np.corrcoef(zc[:, 0], zc[:, 1])[0, 1]) < 0.01          # white across time
    assert abs(np.corrcoef(z[:, 3, 0], z[:, 3, 1])[0, 1]) < 0.01      # independent dims
    # Gaussian shape: kurtosis and tail mass
    x = eps[:, 0, 0]
    assert abs(np.mean(x ** 4) / np.mean(x ** 2) ** 2 - 3.0) < 0.05
    assert abs(np.mean(np.abs(x) > 1.959964) - 0.05) < 0.002


def test_full_covariance_factor():
    from mjmpc_b200.utils.control_utils import generate_noise
    rng = np.random.default_rng(0)
    A = rng.normal(0, 1, (7, 7))
    cov = A @ A.T / 7 + 0.1 * np.eye(7)
    eps = generate_noise(cov, (1.0, 0.0, 0.0), (300000, 2), 7).cpu().numpy()
    emp = np.cov(eps[:, 0].T)
    np.testing.assert_allclose(emp, cov, atol=0.02)


def test_deterministic_and_reseeded_like_reference():
    """Same (seed, step) -> same samples (all n_iters of one MPC step reuse the noise);
    a different step or seed -> different samples."""
    import torch
    from mjmpc_b200.utils.control_utils import generate_noise
    cov = np.eye(7)
    a = generate_noise(cov, (0.25, 0.8, 0.0), (1000, 16), 3, step=2)
    b = generate_noise(cov, (0.25, 0.8, 0.0), (1000, 16), 3, step=2)
    c = generate_noise(cov, (0.25, 0.8, 0.0), (1000, 16), 3, step=3)
    e = generate_noise(cov, (0.25, 0.8, 0.0), (1000, 16), 4, step=2)
    assert torch.equal(a, b)
    assert not torch.equal(a, c) and not torch.equal(a, e)


def test_sharding_invariance_and_layouts():
    import torch
    from mjmpc_b200.utils.control_utils import generate_noise
    cov = np.eye(7) * 0.7
    K, H = 4096, 8
    full = generate_noise(cov, (0.25, 0.8, 0.0), (K, H), 11, step=1)
    parts = [generate_noise(cov, (0.25, 0.8, 0.0), (K // 4, H), 11, step=1, k_offset=r * K // 4, K_global=K)
             for r in range(4)]
    assert torch.equal(full, torch.cat(parts, 0))
    row = torch.empty(K, H, 7, dtype=torch.float64, device="cuda")
    generate_noise(cov, (0.25, 0.8, 0.0), (K, H), 11, step=1, out=row)
    assert torch.equal(full, row)


def test_zero_control_sequence_particle():
    """olgaussian_mpc.py:110-111: the last particle's noise is -mean, so its actions are zero."""
    import torch
    from mjmpc_b200.utils.control_utils import generate_noise
    K, H = 512, 6
    mean = torch.randn(H, 7, dtype=torch.float64, device="cuda")
    eps = generate_noise(np.eye(7), (1.0, 0.0, 0.0), (K, H), 1, zero_last_mean=mean)
    assert torch.equal(eps[-1], -mean)
    ref = generate_noise(np.eye(7), (1.0, 0.0, 0.0), (K, H), 1)
    assert torch.equal(eps[:-1], ref[:-1])
    # sharded: only the shard that owns particle K_global-1 applies it
    first = generate_noise(np.eye(7), (1.0, 0.0, 0.0), (K // 2, H), 1, k_offset=0, K_global=K, zero_last_mean=mean)
    assert torch.equal(first, ref[:K // 2])


def test_d1_and_small_shapes():
    from mjmpc_b200.utils.control_utils import generate_noise
    eps = generate_noise(np.array([[3.0]]), (0.6, 0.5, 0.0), (4096, 64), 0).cpu().numpy()
    assert eps.shape == (4096, 64, 1)
    assert abs(eps[:, 0, 0].var() - 3.0) < 0.25
    one = generate_noise(np.eye(7), (1.0, 0.0, 0.0), (1, 1), 5)
    assert tuple(one.shape) == (1, 1, 7)
    with pytest.raises(ValueError):
        generate_noise(np.eye(9), (1.0, 0.0, 0.0), (4, 4), 5)
