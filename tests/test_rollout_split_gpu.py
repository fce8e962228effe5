"""K1, role-split instantiation (csrc/rollout_reacher_split.cuh: four warps per 32 particles, selected for small
launches): parity with the CPU oracle on identical injected noise, and agreement with the one-thread-per-particle
kernel, through the C ABI."""
import numpy as np
import pytest

from conftest import reference_noise, synthetic_state

pytestmark = pytest.mark.gpu

TRAJ_RTOL = 1e-8      # north_star: state trajectory within 1e-8 relative over the horizon
COST_RTOL = 1e-9


def _gpu_rollout(compiled_model, state, K, H, noise, mean, n_workers=1, randomize=None):
    import torch
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    env = GpuReacherVecEnv(compiled_model, n_workers=n_workers)
    if randomize is not None:
        env.randomize_dynamics(*randomize)
    env.set_env_state(state)
    out = env.rollout_device(K, H, torch.from_numpy(mean).cuda(), None if noise is None else torch.from_numpy(noise).cuda(),
                             want_traj=True, want_ncon=True)
    torch.cuda.synchronize()
    got = {k: v.cpu().numpy() for k, v in out.items()}
    models = env._worker_models
    env.close()
    return got, models


def _oracle(models, state, mean, noise, H=None):
    from oracle import mjstep
    oms = [mjstep.OracleModel(m.tree) for m in models]
    return mjstep.rollout(oms, state["qp"], state["qv"], state["target_pos"], mean, noise, want_traj=True, nthreads=8)


def _assert_parity(got, ref):
    scale = np.abs(ref["qv"]).max(axis=(0, 1))
    err = np.abs(got["qv"] - ref["qv"]).max(axis=(0, 1)) / scale
    assert err.max() < TRAJ_RTOL, err
    np.testing.assert_allclose(got["costs"], ref["costs"], rtol=COST_RTOL, atol=0)
    np.testing.assert_array_equal(got["actions"], ref["actions"])
    np.testing.assert_array_equal(got["ncon"], ref["ncon"])


CASES = {
    "interior": lambda cm: (synthetic_state(cm, 0), None, 1.0),
    "reset_limits_bind": lambda cm: (dict(qp=np.zeros(7), qv=np.zeros(7), target_pos=np.array([.1, .1, .1])), None, 1.0),
    "table_contact": lambda cm: (dict(qp=np.array([0.0, 0.45, 0, -0.2, 0, -0.3, 0.0]), qv=np.zeros(7),
                                      target_pos=np.array([.1, .1, .1])), 1, 0.3),
}


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("K", [512, 77, 4801])
def test_split_rollout_matches_oracle(compiled_model, split_switch, case, K):
    """K = 4801: more 32-particle groups than SMs -> the 64-particles-per-block instantiation (two warps per role),
    ragged last block."""
    st, push, scale = CASES[case](compiled_model)
    H = {512: 32, 77: 5, 4801: 6}[K]
    mean = np.zeros((H, 7))
    if push is not None:
        mean[:, push] = 1.0
    noise = reference_noise(K, H, 7, 11) * scale
    split_switch(1 << 20)
    got, models = _gpu_rollout(compiled_model, st, K, H, noise, mean)
    ref = _oracle(models, st, mean, noise)
    if case != "interior" and K == 512:
        assert (ref["ncon"] > 0).mean() > 0.3
    _assert_parity(got, ref)


def test_split_rollout_agrees_with_thread_per_particle_kernel(compiled_model, split_switch):
    """Same device functions in the same order: the two instantiations agree to rounding (measured: bit for bit on
    sm_100a; the host build reorders nothing either), far inside the 1e-8 contract."""
    st = dict(qp=np.zeros(7), qv=np.zeros(7), target_pos=np.array([.1, .1, .1]))
    K, H = 256, 16
    mean = np.zeros((H, 7))
    noise = reference_noise(K, H, 7, 5)
    split_switch(1 << 20)
    a, _ = _gpu_rollout(compiled_model, st, K, H, noise, mean)
    split_switch(0)
    b, _ = _gpu_rollout(compiled_model, st, K, H, noise, mean)
    np.testing.assert_allclose(a["qv"], b["qv"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(a["costs"], b["costs"], rtol=1e-11)
    np.testing.assert_array_equal(a["ncon"], b["ncon"])
    np.testing.assert_array_equal(a["actions"], b["actions"])


def test_split_rollout_mean_only_and_single_particle(compiled_model, split_switch):
    """noise=None (the mean sequence alone, K=1: what the plant and `use_zero_control_seq` checks launch)."""
    st = synthetic_state(compiled_model, 4)
    H = 8
    mean = np.random.default_rng(0).normal(0, 0.5, (H, 7))
    split_switch(1 << 20)
    got, models = _gpu_rollout(compiled_model, st, 1, H, None, mean)
    ref = _oracle(models, st, mean, None)
    _assert_parity(got, ref)


def test_split_rollout_per_worker_models(compiled_model, split_switch):
    """Randomised per-worker models (global-memory parameters, contiguous particle blocks per worker as in
    subproc_vec_env.py:161-168), block boundaries inside a 32-particle group."""
    st = synthetic_state(compiled_model, 2)
    K, H, W = 96, 8, 6          # 16 particles per worker: two models per 32-lane group
    rnd = (dict(body_mass={"r_forearm_link": [0.3, 0.0], "r_wrist_roll_link": [0.3, 0.0]},
                dof_damping={"r_elbow_flex_joint": [0.2, 0.1]}), 123)
    mean = np.zeros((H, 7))
    noise = reference_noise(K, H, 7, 9)
    split_switch(1 << 20)
    got, models = _gpu_rollout(compiled_model, st, K, H, noise, mean, n_workers=W, randomize=rnd)
    ref = _oracle(models, st, mean, noise)
    _assert_parity(got, ref)


def test_split_threshold_selects_the_kernel(compiled_model, split_switch):
    """The switch is a pure performance choice: costs agree on both sides of the threshold."""
    import torch
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    st = synthetic_state(compiled_model, 3)
    K, H = 128, 4
    noise = torch.from_numpy(reference_noise(K, H, 7, 2)).cuda()
    mean = torch.zeros(H, 7, dtype=torch.float64, device="cuda")
    env = GpuReacherVecEnv(compiled_model)
    env.set_env_state(st)
    outs = []
    for thr in (K, K - 1):
        split_switch(thr)
        outs.append(env.rollout_device(K, H, mean, noise)["costs"].cpu().numpy().copy())
    env.close()
    np.testing.assert_allclose(outs[0], outs[1], rtol=1e-12)


def test_split_rollout_on_random_hard_states(compiled_model, split_switch):
    """The hard-state fuzz of the thread-per-particle kernel (tests/test_kernel_emu_cpu.py) for the role-split one:
    joints up to 0.2 rad beyond their limits, joint speeds up to ~6 rad/s, torque-saturating noise, the arm driven towards the table, randomised masses / inertias / damping, against the C
    oracle within 1e-8 on the whole trajectory."""
    from hypothesis import HealthCheck, assume, given, settings
    from hypothesis import strategies as hst
    from mjmpc_b200.envs.model import randomized_copy, table_clearance
    lo, hi = compiled_model.tree.jnt_range[:, 0], compiled_model.tree.jnt_range[:, 1]
    split_switch(1 << 20)
    seen = dict(cases=0, constrained=0, fast=0)

    @settings(deadline=None, max_examples=24, derandomize=True, suppress_health_check=list(HealthCheck))
    @given(s=hst.integers(0, 2 ** 31 - 1), beyond=hst.sampled_from([0.0, 0.05, 0.2]), vstd=hst.sampled_from([0.5, 2.0, 6.0]),
           nscale=hst.sampled_from([0.3, 1.0, 3.0]), table=hst.booleans(), rand_model=hst.booleans())
    def run(s, beyond, vstd, nscale, table, rand_model):
        rng = np.random.default_rng(s)
        cm = compiled_model
        if rand_model:
            cm, _, _ = randomized_copy(compiled_model, dict(
                body_mass={"r_forearm_link": [0.3, 0.0], "r_wrist_roll_link": [0.3, 0.0], "r_upper_arm_link": [0.3, 0.0]},
                body_inertia={"r_upper_arm_link": [0.2, 0.0]},
                dof_damping={"r_elbow_flex_joint": [0.3, 0.0], "r_shoulder_lift_joint": [0.3, 0.0]}),
                np.random.RandomState(s % 100000), {})
        qp = rng.uniform(lo - beyond, hi + beyond)
        H = 6
        mean = np.zeros((H, 7))
        if table:
            qp[1] = rng.uniform(0.3, 0.5); qp[3] = rng.uniform(-0.4, 0.0); qp[5] = rng.uniform(-0.5, 0.0)
            mean[:, 1] = 1.0
        assume(table_clearance(cm.tree, qp) > -0.02)
        st = dict(qp=qp, qv=rng.normal(0, vstd, 7), target_pos=rng.uniform([-.3, -.2, -.25], [.3, .2, .25]))
        K = 40
        noise = nscale * reference_noise(K, H, 7, s % 1000)
        got, models = _gpu_rollout(cm, st, K, H, noise, mean)
        ref = _oracle(models, st, mean, noise)
        assert np.isfinite(got["qv"]).all()
        scale = np.abs(ref["qv"]).max(axis=(0, 1))
        assert (np.abs(got["qv"] - ref["qv"]).max(axis=(0, 1)) / scale).max() < TRAJ_RTOL
        np.testing.assert_allclose(got["costs"], ref["costs"], rtol=1e-8)
        np.testing.assert_array_equal(got["ncon"], ref["ncon"])
        seen["cases"] += 1
        seen["constrained"] += int((ref["ncon"] > 0).any())

    run()
    assert seen["cases"] >= 15 and seen["constrained"] >= 8, seen


def test_split_rollout_fast_joint_takes_the_exact_sincos_fallback(compiled_model, split_switch):
    """A joint that turns by more than 0.25 rad in one substep (40 rad/s on the shoulder pan, 30 on the wrist roll)
    is outside the Taylor range of the angle-addition update: the kernel re-evaluates sin / cos exactly.  Two env
    steps from inside the ranges (the centrifugal load drives other joints into their limits on the way)."""
    split_switch(1 << 20)
    qv = np.zeros(7); qv[0] = 40.0; qv[6] = -30.0
    st = dict(qp=np.array([-1.8, 0.3, 0.0, -1.0, 0.0, -0.5, 1.3]), qv=qv, target_pos=np.array([.1, .1, .1]))
    K, H = 33, 2
    mean = np.zeros((H, 7))
    noise = 0.5 * reference_noise(K, H, 7, 3)
    got, models = _gpu_rollout(compiled_model, st, K, H, noise, mean)
    ref = _oracle(models, st, mean, noise)
    assert np.abs(ref["qv"][:, 0, 7]).min() > 25.0
    _assert_parity(got, ref)
