"""K1 parity: CUDA rollout (through the C ABI) vs the CPU oracle on identical injected noise."""
import numpy as np
import pytest

from conftest import reference_noise, synthetic_state

pytestmark = pytest.mark.gpu

# north_star tolerance: state trajectory within 1e-8 relative over the horizon; costs 1e-9.
TRAJ_RTOL = 1e-8
COST_RTOL = 1e-9


def _run_both(compiled_model, oracle_model, state, K, H, seed, mean=None, n_workers=1, scale=1.0):
    import torch
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    from oracle import mjstep
    noise = reference_noise(K, H, 7, seed) * scale
    mean = np.zeros((H, 7)) if mean is None else mean
    env = GpuReacherVecEnv(compiled_model, n_workers=n_workers)
    env.set_env_state(state)
    out = env.rollout_device(K, H, torch.from_numpy(mean).cuda(), torch.from_numpy(noise).cuda(),
                             want_traj=True, want_obs=True, want_ncon=True)
    torch.cuda.synchronize()
    ref = mjstep.rollout(oracle_model, state["qp"], state["qv"], state["target_pos"], mean, noise,
                         want_traj=True, want_obs=True, nthreads=8)
    env.close()
    return {k: v.cpu().numpy() for k, v in out.items()}, ref, noise


def _assert_parity(got, ref):
    scale = np.abs(ref["qv"]).max(axis=(0, 1))
    err = np.abs(got["qv"] - ref["qv"]).max(axis=(0, 1)) / scale
    assert err.max() < TRAJ_RTOL, err
    np.testing.assert_allclose(got["costs"], ref["costs"], rtol=COST_RTOL, atol=0)
    np.testing.assert_array_equal(got["actions"], ref["actions"])       # mean + noise, one add: bit exact
    np.testing.assert_allclose(got["next_observations"], ref["next_observations"], rtol=1e-7, atol=1e-9)
    np.testing.assert_array_equal(got["ncon"], ref["ncon"])


def test_rollout_interior_state(compiled_model, oracle_model):
    st = synthetic_state(compiled_model, 0)
    got, ref, _ = _run_both(compiled_model, oracle_model, st, K=2048, H=32, seed=3)
    _assert_parity(got, ref)


def test_rollout_reset_state_limits_bind(compiled_model, oracle_model):
    """From the env's true reset state qpos=qvel=0 the elbow and wrist-flex limits (upper bound 0)
    bind within the first substeps."""
    st = dict(qp=np.zeros(7), qv=np.zeros(7), qa=np.zeros(7), target_pos=np.array([.1, .1, .1]), timestep=0)
    got, ref, _ = _run_both(compiled_model, oracle_model, st, K=2048, H=32, seed=4)
    assert (ref["ncon"] > 0).mean() > 0.5
    _assert_parity(got, ref)


def test_rollout_table_contact(compiled_model, oracle_model):
    """Shoulder-lift pushed down: the end-effector sphere meets the table plane (contact row)."""
    st = dict(qp=np.array([0.0, 0.45, 0, -0.2, 0, -0.3, 0.0]), qv=np.zeros(7), qa=np.zeros(7),
              target_pos=np.array([.1, .1, .1]), timestep=0)
    H = 32
    mean = np.zeros((H, 7)); mean[:, 1] = 1.0
    got, ref, _ = _run_both(compiled_model, oracle_model, st, K=512, H=H, seed=5, mean=mean, scale=0.3)
    assert ref["next_observations"][:, :, 16].min() < -0.345      # sphere reached the table
    _assert_parity(got, ref)


def test_rollout_shipped_config_shape(compiled_model, oracle_model):
    """configs[0]: the shipped reacher_7dof MPPI sizes K=32, H=16 (reacher_7dof-v0.yml:20-30)."""
    st = synthetic_state(compiled_model, 7)
    got, ref, _ = _run_both(compiled_model, oracle_model, st, K=32, H=16, seed=123)
    _assert_parity(got, ref)


def test_rollout_ragged_sizes(compiled_model, oracle_model):
    """K not a multiple of the block size, H=1, and the mean-only (noise=None) path."""
    import torch
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    from oracle import mjstep
    st = synthetic_state(compiled_model, 9)
    got, ref, _ = _run_both(compiled_model, oracle_model, st, K=77, H=1, seed=1)
    _assert_parity(got, ref)
    env = GpuReacherVecEnv(compiled_model)
    env.set_env_state(st)
    mean = np.random.default_rng(0).normal(0, 0.5, (8, 7))
    out = env.rollout_device(1, 8, torch.from_numpy(mean).cuda(), None)
    ref = mjstep.rollout(oracle_model, st["qp"], st["qv"], st["target_pos"], mean, None)
    np.testing.assert_allclose(out["costs"].cpu().numpy(), ref["costs"], rtol=COST_RTOL)
    env.close()


def test_rollout_layouts_agree(compiled_model):
    """Row-major (K,H,7) noise and the particle-minor (H,7,K) layout give identical bits."""
    import torch
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    st = synthetic_state(compiled_model, 11)
    K, H = 1000, 8
    noise = torch.from_numpy(reference_noise(K, H, 7, 2)).cuda()
    noise_t = noise.permute(1, 2, 0).contiguous().permute(2, 0, 1)
    assert noise_t.stride() == (1, 7 * K, K)
    env = GpuReacherVecEnv(compiled_model)
    env.set_env_state(st)
    mean = torch.zeros(H, 7, dtype=torch.float64, device="cuda")
    a = env.rollout_device(K, H, mean, noise)
    b = env.rollout_device(K, H, mean, noise_t, costs=torch.empty(K, H, dtype=torch.float64, device="cuda"),
                           actions=torch.empty(K, H, 7, dtype=torch.float64, device="cuda"))
    assert torch.equal(a["costs"], b["costs"]) and torch.equal(a["actions"], b["actions"])
    env.close()


def test_rollout_bad_arguments(compiled_model):
    import torch
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    env = GpuReacherVecEnv(compiled_model, n_workers=8)
    env.set_env_state(synthetic_state(compiled_model, 1))
    mean = torch.zeros(4, 7, dtype=torch.float64, device="cuda")
    with pytest.raises(AssertionError):      # subproc_vec_env.py:162
        env.rollout_device(12, 4, mean, torch.zeros(12, 4, 7, dtype=torch.float64, device="cuda"))
    with pytest.raises(NotImplementedError):
        env.rollout(8, 4, np.zeros((4, 7)), np.zeros((8, 4, 7)), mode="closed_loop_nn")
    env.close()


def test_rollout_per_worker_models(compiled_model):
    """Dynamics randomisation: each contiguous particle block runs its own model
    (subproc_vec_env.py:304-312); checked against one oracle model per worker."""
    import torch
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    from oracle import mjstep
    st = synthetic_state(compiled_model, 21)
    K, H, W = 64, 16, 4
    env = GpuReacherVecEnv(compiled_model, n_workers=W)
    params = dict(body_mass={"r_forearm_link": [0.3, 0.1], "r_wrist_roll_link": [0.2, 0.0]},
                  dof_damping={"r_elbow_flex_joint": [0.5, 0.2]},
                  body_inertia={"r_upper_arm_link": [0.3, 0.0]})
    env.randomize_dynamics(params, base_seed=5)
    env.set_env_state(st)
    noise = reference_noise(K, H, 7, 8)
    noise[K // W] = noise[0]            # the first particle of worker 1 gets the controls of the first particle of worker 0
    mean = np.zeros((H, 7))
    out = env.rollout_device(K, H, torch.from_numpy(mean).cuda(), torch.from_numpy(noise).cuda(), want_traj=True)
    oms = [mjstep.OracleModel(m.tree) for m in env._worker_models]
    ref = mjstep.rollout(oms, st["qp"], st["qv"], st["target_pos"], mean, noise, want_traj=True)
    got = out["qv"].cpu().numpy()
    scale = np.abs(ref["qv"]).max(axis=(0, 1))
    assert (np.abs(got - ref["qv"]).max(axis=(0, 1)) / scale).max() < TRAJ_RTOL
    # and the models really differ between workers: same state, same controls, different trajectories
    assert np.abs(ref["qv"][0] - ref["qv"][K // W]).max() > 1e-6
    assert np.abs(got[0] - got[K // W]).max() > 1e-6
    env.close()


@pytest.mark.parametrize("n_workers,extra", [(1, False), (1, True), (4, False)])
def test_fused_noise_equals_materialised_noise(compiled_model, n_workers, extra, split_switch):
    """Fused K2: the rollout kernel drawing its own noise must give exactly the actions and costs of the
    two-kernel path (noise tensor from mjb_generate_noise fed to the rollout) -- same counters, same rounding.
    (Bit equality is a property of ONE instantiation: the small-launch instantiations are switched off here.)"""
    import torch
    split_switch(0)
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    from mjmpc_b200.utils.control_utils import NoiseSpec
    K, H = 1000 if n_workers == 1 else 1024, 12
    env = GpuReacherVecEnv(compiled_model, n_workers=n_workers)
    if n_workers > 1:
        env.randomize_dynamics(dict(body_mass={"r_forearm_link": [0.3, 0.0]}), base_seed=2)
    env.set_env_state(synthetic_state(compiled_model, 4))
    rng = np.random.default_rng(3)
    A = rng.normal(0, 1, (7, 7))
    cov = torch.from_numpy(A @ A.T / 7 + 0.2 * np.eye(7)).cuda()
    mean = torch.from_numpy(rng.normal(0, 0.3, (H, 7))).cuda()
    for zero_last, k_off, Kg in [(False, 0, K), (True, 0, K), (True, 5000, 5000 + K), (False, 77, 100000)]:
        spec = NoiseSpec(cov, [0.25, 0.8, 0.1], (K, H), 1234, step=7, k_offset=k_off, K_global=Kg, zero_last=zero_last,
                         mean=mean)
        noise = spec.materialize()
        a = env.rollout_device(K, H, mean, noise, want_traj=extra)
        b = env.rollout_device(K, H, mean, spec, want_traj=extra)
        assert torch.equal(a["actions"], b["actions"])
        assert torch.equal(a["costs"], b["costs"])
        if extra:
            assert torch.equal(a["qv"], b["qv"])
        if zero_last:
            assert torch.all(b["actions"][-1] == 0.0)
    # a device-resident step counter selects the same stream as the integer step
    step_d = torch.tensor([7], dtype=torch.int64, device="cuda")
    spec_d = NoiseSpec(cov, [0.25, 0.8, 0.1], (K, H), 1234, step=step_d, k_offset=77, K_global=100000)
    c = env.rollout_device(K, H, mean, spec_d)
    assert torch.equal(c["actions"], b["actions"])
    with pytest.raises(ValueError):
        env.rollout_device(K, H + 1, torch.zeros(H + 1, 7, dtype=torch.float64, device="cuda"), spec)
    env.close()
