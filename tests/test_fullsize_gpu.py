"""BASELINE.json's full size (MPPI reacher_7dof, K=65536, H=32) on the GPU.  The CPU oracle cannot roll out
2.1 M particle-steps in seconds, so parity at this size goes through size-independent properties:

* a particle's trajectory does not depend on which launch it is part of -> a random subset of the full
  launch is re-run alone (bit-identical) and checked against the oracle (1e-8, north_star tolerance);
* the update of the full (K,H,7) tensors against the numpy restatement of the reference (1e-10), softmax
  weights summing to one, the elite index set bit-exact, logical shards combining to the unsharded result.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import synthetic_state

pytestmark = pytest.mark.gpu
K, H, D = 65536, 32, 7
RTOL = 1e-10


def _common():
    return dict(d_state=25, d_obs=20, d_action=D, action_lows=-np.ones(D), action_highs=np.ones(D))


@pytest.fixture(scope="module")
def full_rollout(compiled_model):
    """One K=65536, H=32 launch on Philox noise from a synthetic start state; tensors stay on the device."""
    import torch
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    from mjmpc_b200.utils.control_utils import generate_noise
    st = synthetic_state(compiled_model, 11)
    env = GpuReacherVecEnv(compiled_model)
    env.set_env_state(st)
    mean = torch.from_numpy(np.random.default_rng(3).normal(0, 0.2, (H, D))).cuda()
    noise = generate_noise(torch.eye(D, dtype=torch.float64, device="cuda"), [0.25, 0.8, 0.0], (K, H), 123, step=5)
    out = env.rollout_device(K, H, mean, noise)
    torch.cuda.synchronize()
    yield dict(env=env, state=st, mean=mean, noise=noise, costs=out["costs"], actions=out["actions"])
    env.close()


def test_full_launch_subset_is_launch_independent_and_matches_oracle(full_rollout, oracle_model, split_switch):
    import torch
    from oracle import mjstep
    f = full_rollout
    split_switch(0)          # bit equality across launch sizes holds within ONE kernel instantiation
    ids = np.sort(np.random.default_rng(0).choice(K, 1024, replace=False))
    ids[0], ids[-1] = 0, K - 1                                        # first and last particle of the launch
    idt = torch.from_numpy(ids).cuda()
    sub_noise = f["noise"][idt].contiguous()
    sub = f["env"].rollout_device(len(ids), H, f["mean"], sub_noise)           # same kernel instantiation as the full launch
    np.testing.assert_array_equal(sub["costs"].cpu().numpy(), f["costs"][idt].cpu().numpy())
    np.testing.assert_array_equal(sub["actions"].cpu().numpy(), f["actions"][idt].cpu().numpy())
    # the role-split instantiation a small shard takes (sin / cos advanced by angle addition) agrees to rounding
    split_switch(1 << 20)
    sub = f["env"].rollout_device(len(ids), H, f["mean"], sub_noise)
    np.testing.assert_allclose(sub["costs"].cpu().numpy(), f["costs"][idt].cpu().numpy(), rtol=1e-11)
    np.testing.assert_array_equal(sub["actions"].cpu().numpy(), f["actions"][idt].cpu().numpy())
    sub = f["env"].rollout_device(len(ids), H, f["mean"], sub_noise, want_traj=True)
    st = f["state"]
    ref = mjstep.rollout(oracle_model, st["qp"], st["qv"], st["target_pos"], f["mean"].cpu().numpy(),
                         np.ascontiguousarray(sub_noise.cpu().numpy()), want_traj=True, nthreads=8)
    scale = np.abs(ref["qv"]).max(axis=(0, 1))
    assert (np.abs(sub["qv"].cpu().numpy() - ref["qv"]).max(axis=(0, 1)) / scale).max() < 1e-8
    np.testing.assert_allclose(sub["costs"].cpu().numpy(), ref["costs"], rtol=1e-9)
    assert np.isfinite(f["costs"].cpu().numpy()).all()


def test_full_size_mppi_update_matches_numpy(full_rollout):
    import torch
    from mjmpc_b200 import _lib
    from mjmpc_b200.control import MPPI
    from oracle import control_np as O
    f = full_rollout
    c = MPPI(horizon=H, init_cov=1.0, base_action='null', lam=0.2, num_particles=K, step_size=0.9, alpha=0, gamma=0.99,
             n_iters=1, filter_coeffs=[0.25, 0.8, 0.0], seed=1, **_common())
    mean0 = f["mean"].cpu().numpy()
    c.mean_action = mean0
    traj = dict(costs=f["costs"], actions=f["actions"])
    costs, actions = f["costs"].cpu().numpy(), np.ascontiguousarray(f["actions"].cpu().numpy())
    gs = O.gamma_seq(0.99, H)
    want_mean, want_w = O.mppi_update(mean0, np.diag([1.0] * D), costs, actions, gs, 0.2, 0, 0.9)
    want_val = O.mppi_value(mean0, np.diag([1.0] * D), costs, actions, gs, 0.2, 0)
    assert c._calc_val(traj) == pytest.approx(want_val, rel=RTOL)
    stats = c._softmax_update(traj["costs"], traj["actions"], c.lam, control_cost=True, apply=False)
    w = torch.empty(K, dtype=torch.float64, device="cuda")
    _lib.check(_lib.lib().mjb_softmax_weights(_lib.ptr(c._last_total), C.c_int(K), _lib.ptr(stats), C.c_int(0),
                                              C.c_double(c.lam), _lib.ptr(w), _lib.stream_ptr()))
    assert abs(w.sum().item() - 1.0) < 1e-12
    np.testing.assert_allclose(w.cpu().numpy(), want_w, rtol=1e-9, atol=1e-300)
    c._update_distribution(traj)
    np.testing.assert_allclose(c.mean_action, want_mean, rtol=RTOL, atol=1e-13)


def test_full_size_cost_to_go_and_elite_set_bit_exact(full_rollout):
    from mjmpc_b200.control import CEM
    from mjmpc_b200.utils.control_utils import cost_to_go
    from oracle import control_np as O
    f = full_rollout
    costs, actions = f["costs"].cpu().numpy(), np.ascontiguousarray(f["actions"].cpu().numpy())
    gs = O.gamma_seq(0.97, H)
    ctg = O.cost_to_go(costs.copy(), gs)
    np.testing.assert_array_equal(cost_to_go(f["costs"], gs).cpu().numpy(), ctg)
    c = CEM(horizon=H, init_cov=1.0, base_action='null', elite_frac=0.2, num_particles=K, step_size=0.8, gamma=0.97,
            n_iters=1, beta=0.0, cov_type='full', filter_coeffs=[0.25, 0.8, 0.0], seed=1, **_common())
    mean0 = f["mean"].cpu().numpy()
    c.mean_action = mean0
    c._update_distribution(dict(costs=f["costs"], actions=f["actions"]))
    E = c.num_elite
    assert E == int(K * 0.2)
    # ties at the threshold go to the lower index: stable argsort states the same rule
    want_ids = np.sort(np.argsort(ctg[:, 0], kind="stable")[:E])
    np.testing.assert_array_equal(c.elite_ids.cpu().numpy(), want_ids)
    ea = actions[want_ids]
    ed = (ea - mean0[None]).reshape(E * H, D)
    np.testing.assert_allclose(c.mean_action, 0.2 * mean0 + 0.8 * ea.mean(0), rtol=RTOL, atol=1e-13)
    np.testing.assert_allclose(c.cov_action, 0.2 * np.diag([1.0] * D) + 0.8 * np.cov(ed, rowvar=False), rtol=1e-9, atol=1e-13)


def test_full_size_logical_shards_equal_unsharded(full_rollout):
    """8 logical shards of 8192 particles (the per-GPU size of the 8-GPU run): phase-1 partials per slice,
    rank-ordered combine == the unsharded update; the noise of a shard does not depend on the sharding."""
    import torch
    from mjmpc_b200 import _lib
    from mjmpc_b200.control import MPPI
    from mjmpc_b200.utils.control_utils import generate_noise
    f = full_rollout
    L = _lib.lib()
    N, kl = 8, K // 8
    c = MPPI(horizon=H, init_cov=1.0, base_action='null', lam=0.2, num_particles=K, step_size=1.0, alpha=1, gamma=1.0,
             n_iters=1, filter_coeffs=[0.25, 0.8, 0.0], seed=1, **_common())
    c.mean_action = f["mean"].cpu().numpy()
    c._update_distribution(dict(costs=f["costs"], actions=f["actions"]))
    want = c.mean_action
    P = L.mjb_softmax_partial_doubles(H, D, 0, 0)
    parts = torch.empty((N, P), dtype=torch.float64, device="cuda")
    g = np.ones(H)
    for r in range(N):
        costs = f["costs"][r * kl:(r + 1) * kl]
        actions = f["actions"][r * kl:(r + 1) * kl]
        a = _lib.SoftmaxArgs()
        a.K, a.H, a.d = kl, H, D
        a.costs = costs.data_ptr(); a.costs_sk, a.costs_st = costs.stride()
        a.actions = actions.data_ptr(); a.act_sk, a.act_st, a.act_sj = actions.stride()
        a.mean = f["mean"].data_ptr()
        a.gamma_seq, a.lam = g.ctypes.data, 0.2
        total = torch.empty((1, kl), dtype=torch.float64, device="cuda")
        scratch = torch.empty(int(L.mjb_softmax_scratch_doubles(kl, H, D, 0)), dtype=torch.float64, device="cuda")
        a.total, a.scratch, a.partials = total.data_ptr(), scratch.data_ptr(), parts[r].data_ptr()
        _lib.check(L.mjb_softmax_partials(C.byref(a), _lib.stream_ptr()))
        sn = generate_noise(torch.eye(D, dtype=torch.float64, device="cuda"), [0.25, 0.8, 0.0], (kl, H), 123, step=5,
                            k_offset=r * kl, K_global=K)
        assert torch.equal(sn, f["noise"][r * kl:(r + 1) * kl])
    mean = f["mean"].clone()
    cb = _lib.CombineArgs()
    cb.H, cb.d, cb.n_shards, cb.K_global = H, D, N, K
    cb.partials, cb.lam, cb.step_size = parts.data_ptr(), 0.2, 1.0
    cb.mean = mean.data_ptr()
    _lib.check(L.mjb_softmax_combine(C.byref(cb), _lib.stream_ptr()))
    np.testing.assert_allclose(mean.cpu().numpy(), want, rtol=1e-12, atol=1e-14)
