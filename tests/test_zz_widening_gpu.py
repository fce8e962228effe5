"""GPU tests of the pieces added after the last GPU session of round 1 (written without GPU access; they sort
last so that a surprise here cannot hide the rest of the suite behind `-x`)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_continual_reacher_redraws_target_every_50_plant_steps():
    """continual_reacher-v0 (reference reacher_env.py:128-132): the observation of step 50 still refers to the
    old target (get_obs precedes trigger_timed_events, :36-38), the state in the infos to the new one."""
    from mjmpc_b200.envs.gpu_reacher_env import GpuContinualReacherEnv
    env = GpuContinualReacherEnv()
    env.reset(seed=7)
    assert env._max_episode_steps == 250
    targets = [env.target_pos.copy()]
    for t in range(1, 102):
        old = env.target_pos.copy()
        ob, r, done, info = env.step(np.zeros(7))
        np.testing.assert_allclose(ob[17:20], ob[14:17] - old, atol=1e-15)
        if t in (50, 100):
            assert not np.array_equal(env.target_pos, old)
            np.testing.assert_array_equal(info["state"]["target_pos"], env.target_pos)
            targets.append(env.target_pos.copy())
        else:
            np.testing.assert_array_equal(env.target_pos, old)
    # the same generator stream as the reference's target_reset: x, y, z draws in order
    rng = np.random.RandomState(7)
    for tgt in targets:
        want = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.2, 0.2), rng.uniform(-0.25, 0.25)])
        np.testing.assert_array_equal(tgt, want)
    env.real_step = False                      # planning copies never fire timed events (reacher_env.py:13,130)
    env.env_timestep = 149
    old = env.target_pos.copy()
    env.step(np.zeros(7))
    np.testing.assert_array_equal(env.target_pos, old)
    env.close()


def _linear_policy(seed, scale=0.1):
    rng = np.random.default_rng(seed)
    W = scale * rng.normal(0, 1, (21, 7))
    W[14:20] *= 2.0
    return W


@pytest.mark.parametrize("case", ["interior", "reset"])
def test_closed_loop_linear_rollout_matches_oracle(case):
    """mode="closed_loop_linear" (gym_env_wrapper.py:129-136) through the C ABI against the CPU oracle on the same
    noise: actions (= policy output + noise), trajectories and costs within the north-star 1e-8."""
    import torch
    from conftest import reference_noise, synthetic_state
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    from mjmpc_b200.envs.model import compile_model, reacher7dof_spec
    from oracle import mjstep
    cm = compile_model(reacher7dof_spec())
    om = mjstep.OracleModel(cm.tree)
    env = GpuReacherVecEnv(cm)
    K, H = 256, 16
    st = synthetic_state(cm, 5)
    if case == "reset":
        st = dict(st, qp=np.zeros(7), qv=np.zeros(7))
    W = _linear_policy(1)
    noise = np.ascontiguousarray(0.5 * reference_noise(K, H, 7, 9))
    ref = mjstep.rollout(om, st["qp"], st["qv"], st["target_pos"], None, noise, want_traj=True, want_obs=True,
                         nthreads=4, policy_w=W)
    env.set_env_state(st)
    out = env.rollout_device(K, H, torch.from_numpy(W).cuda(), torch.from_numpy(noise).cuda(), want_traj=True,
                             want_obs=True, closed_loop=True)
    scale = np.abs(ref["qv"]).max(axis=(0, 1))
    assert (np.abs(out["qv"].cpu().numpy() - ref["qv"]).max(axis=(0, 1)) / scale).max() < 1e-8
    np.testing.assert_allclose(out["actions"].cpu().numpy(), ref["actions"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(out["costs"].cpu().numpy(), ref["costs"], rtol=1e-8)
    np.testing.assert_allclose(out["next_observations"].cpu().numpy(), ref["next_observations"], rtol=1e-8, atol=1e-9)
    # a bias-only policy is the open-loop rollout with a constant mean: bit-identical through the same kernel
    Wb = np.zeros((21, 7)); Wb[20] = [0.3, -0.2, 0.1, 0.0, 0.2, -0.1, 0.05]
    a = env.rollout_device(K, H, torch.from_numpy(Wb).cuda(), torch.from_numpy(noise).cuda(), closed_loop=True)
    b = env.rollout_device(K, H, torch.from_numpy(np.tile(Wb[20], (H, 1))).cuda(), torch.from_numpy(noise).cuda(),
                           want_obs=True)                 # observations: the thread-per-particle trajectory instantiation
    assert torch.equal(a["costs"], b["costs"]) and torch.equal(a["actions"], b["actions"])
    env.close()


def test_closed_loop_linear_through_the_reference_rollout_fn():
    """The closure contract (examples/example_mpc.py:112-133) with numpy in / numpy out, as CLGaussianMPC calls it
    (clgaussian_mpc.py:104-106): observations[0, 0] is the observation at the set state."""
    from conftest import reference_noise, synthetic_state
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    from mjmpc_b200.envs.model import compile_model, forward_kinematics, reacher7dof_spec
    cm = compile_model(reacher7dof_spec())
    env = GpuReacherVecEnv(cm)
    st = synthetic_state(cm, 8)
    env.set_env_state(st)
    K, H = 64, 8
    W = _linear_policy(2)
    noise = 0.3 * reference_noise(K, H, 7, 4)
    traj = env.rollout_fn(K, H, W, noise, mode="closed_loop_linear")
    assert traj["observations"].shape == (K, H, 20) and traj["actions"].shape == (K, H, 7)
    hand = forward_kinematics(cm.tree, st["qp"])["hand"]
    obs0 = np.concatenate([st["qp"], st["qv"], hand, hand - st["target_pos"]])
    np.testing.assert_allclose(traj["observations"][0, 0], obs0, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(traj["actions"][:, 0], (W.T @ np.append(obs0, 1.0))[None] + noise[:, 0], rtol=1e-10, atol=1e-12)
    np.testing.assert_array_equal(traj["observations"][:, 1:], traj["next_observations"][:, :-1])
    with pytest.raises(NotImplementedError):
        env.rollout_fn(K, H, W, noise, mode="closed_loop_nn")
    env.close()


def test_batched_instances_take_state_rows():
    """Sweeps of many instances: an (n_ctrl, 17) array (host or device) of [qpos | qvel | target] rows is accepted
    in place of the list of state dicts and gives the same step, bit for bit."""
    import torch
    from conftest import synthetic_state
    from mjmpc_b200.control import MPPI
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    from mjmpc_b200.envs.model import compile_model, reacher7dof_spec
    cm = compile_model(reacher7dof_spec())
    B, K, H = 6, 32, 8
    dicts = [synthetic_state(cm, 40 + b) for b in range(B)]
    rows = np.stack([np.concatenate([s["qp"], s["qv"], s["target_pos"]]) for s in dicts])
    outs = []
    for form in ("dicts", "rows", "device"):
        env = GpuReacherVecEnv(cm, n_workers=B)
        c = MPPI(d_state=25, d_obs=20, d_action=7, action_lows=-np.ones(7), action_highs=np.ones(7), horizon=H, init_cov=1.0,
                 base_action='null', lam=0.2, num_particles=K, step_size=1.0, alpha=1, gamma=1.0, n_iters=1,
                 filter_coeffs=[0.25, 0.8, 0.0], seed=3, batch_size=B)
        c.set_sim_state_fn, c.rollout_fn = env.set_env_state, env.rollout_fn
        st = {"dicts": dicts, "rows": rows, "device": torch.from_numpy(rows).cuda()}[form]
        acts = [c.optimize(st)[0] for _ in range(2)]
        outs.append((np.stack(acts), c.mean_action))
        env.close()
    for o in outs[1:]:
        np.testing.assert_array_equal(o[0], outs[0][0])
        np.testing.assert_array_equal(o[1], outs[0][1])
    with pytest.raises(ValueError):
        GpuReacherVecEnv(cm).set_env_state(np.zeros((2, 5)))


def test_signed_zero_costs_tie_like_numpy():
    """A reward of exactly 0.0 becomes a cost of -0.0 under `costs = -rewards`; numpy compares -0.0 == +0.0, so
    argmin / the elite set break that tie by index, not by sign bit."""
    import ctypes as C
    import torch
    from mjmpc_b200 import _lib
    K = 3000
    z = np.abs(np.random.RandomState(0).normal(2, 1, K)) + 1.0
    z[40] = 0.0; z[90] = -0.0; z[7] = 0.0; z[2999] = -0.0
    zt = torch.from_numpy(z).cuda()
    idx = torch.zeros(1, dtype=torch.int64, device="cuda"); val = torch.zeros(1, dtype=torch.float64, device="cuda")
    L = _lib.lib()
    _lib.check(L.mjb_argmin(_lib.ptr(zt), C.c_longlong(K), _lib.ptr(idx), _lib.ptr(val), _lib.stream_ptr()))
    assert int(idx.item()) == 7 == int(np.argmin(z))
    flags = torch.zeros(K, dtype=torch.uint8, device="cuda"); sel = torch.zeros(3, dtype=torch.int64, device="cuda")
    scr = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    _lib.check(L.mjb_select_elites(_lib.ptr(zt), C.c_longlong(K), C.c_longlong(3), _lib.ptr(flags), _lib.ptr(sel), _lib.ptr(scr),
                                   _lib.stream_ptr()))
    np.testing.assert_array_equal(np.flatnonzero(flags.cpu().numpy()), [7, 40, 90])
    np.testing.assert_array_equal(np.sort(sel.cpu().numpy()), np.sort(np.argsort(z, kind="stable")[:3]))


@pytest.mark.parametrize("d", [2, 3, 4, 5, 6, 8])
def test_noise_kernel_every_action_dimension(d):
    """K2 is instantiated for d_action = 1..8 (1 and 7 are covered in test_noise_gpu.py): full covariance and the
    AR filter for the other sizes."""
    from mjmpc_b200.utils.control_utils import generate_noise
    rng = np.random.default_rng(d)
    A = rng.normal(0, 1, (d, d))
    cov = A @ A.T / d + 0.2 * np.eye(d)
    b = (0.25, 0.8, 0.1)
    eps = generate_noise(cov, b, (80000, 4), 3, step=1).cpu().numpy()
    assert eps.shape == (80000, 4, d) and np.isfinite(eps).all()
    np.testing.assert_allclose(np.cov(eps[:, 0].T), cov, atol=0.05)
    z = (eps[:, 2:] - b[1] * eps[:, 1:-1] - b[2] * eps[:, :-2]) / b[0]
    np.testing.assert_allclose(np.cov(z[:, 0].T), cov, atol=0.05)


@pytest.mark.parametrize("name", ["mppi", "cem", "dmd", "random_shooting"])
def test_unmodified_reference_controller_on_this_rollout_backend(name, monkeypatch):
    """SURVEY 8(f-1), from the other side: the reference's own controller class (imported read-only from
    /root/reference, so this runs where that tree exists -- the container, under the host emulation -- and is
    skipped on the GPU box) drives this package's rollout backend through the injected rollout_fn /
    set_sim_state_fn with numpy arrays, exactly as examples/example_mpc.py:154-155 wires it.  Fed the same noise,
    this package's controller must produce the same actions over several hot-started MPC steps."""
    import os
    import sys
    import torch
    from conftest import ROOT, synthetic_state
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import refload
    if not refload.available():
        pytest.skip("reference tree not present")
    R = refload.load()
    import mjmpc_b200.control as ctl
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    from mjmpc_b200.envs.model import compile_model, reacher7dof_spec
    monkeypatch.setenv("MJB_FUSED_STEP", "0")          # the step-by-step path calls sample_noise(), patched below
    cm = compile_model(reacher7dof_spec())
    K, H = 128, 10
    common = dict(d_state=25, d_obs=20, d_action=7, horizon=H, init_cov=0.8, base_action='null', num_particles=K,
                  step_size=0.9, gamma=0.98, n_iters=1, action_lows=-np.ones(7), action_highs=np.ones(7),
                  filter_coeffs=[0.25, 0.8, 0.0], seed=17)
    extra = dict(mppi=dict(lam=0.3, alpha=0), cem=dict(elite_frac=0.25, beta=0.1, cov_type='diagonal'),
                 dmd=dict(lam=0.3, beta=0.1, update_cov=True, cov_type='full'), random_shooting=dict())[name]
    ref_cls = dict(mppi=R.mppi.MPPI, cem=R.cem.CEM, dmd=R.dmd.DMDMPC, random_shooting=R.rs.RandomShooting)[name]
    own_cls = dict(mppi=ctl.MPPI, cem=ctl.CEM, dmd=ctl.DMDMPC, random_shooting=ctl.RandomShooting)[name]
    states = [synthetic_state(cm, 60 + s) for s in range(4)]

    env_r = GpuReacherVecEnv(cm)
    ref = ref_cls(**common, **extra)
    ref.set_sim_state_fn = env_r.set_env_state
    ref.rollout_fn = env_r.rollout_fn
    want = np.stack([ref.optimize(dict(s))[0] for s in states])

    env_o = GpuReacherVecEnv(cm)
    own = own_cls(**common, **extra)
    own.set_sim_state_fn = env_o.set_env_state
    own.rollout_fn = env_o.rollout_fn
    own.sample_noise = lambda: torch.from_numpy(R.control_utils.generate_noise(
        own.cov_action, own.filter_coeffs, (K, H), own.seed_val + own.num_steps)).cuda()
    got = np.stack([own.optimize(dict(s))[0] for s in states])
    np.testing.assert_allclose(got, want, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(own.mean_action, ref.mean_action, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(own.cov_action, ref.cov_action, rtol=1e-8, atol=1e-10)
    env_r.close(); env_o.close()


def test_pendulum_plant_follows_the_reference_env():
    """GpuPendulumEnv (the plant of SimplePendulum-v0) against the golden trajectory of the reference's PendulumEnv:
    same states, rewards and observations step by step; reset draws like pendulum.py:52-56."""
    from golden_util import load
    from mjmpc_b200.envs.gpu_pendulum import GpuPendulumEnv
    g = load("pendulum")
    env = GpuPendulumEnv(seed=3)
    ob = env.reset()
    rng = np.random.RandomState(3)
    high = np.array([np.pi, 1])
    np.testing.assert_array_equal(env.get_env_state()["state"], rng.uniform(low=-high, high=high))
    np.testing.assert_allclose(ob, [np.cos(env.state[0]), np.sin(env.state[0]), env.state[1]], rtol=0, atol=0)
    env.set_env_state(dict(state=g["state0"]))
    for t in range(g["mean"].shape[0]):
        ob, rew, done, info = env.step(g["mean"][t] + g["noise"][0, t])
        np.testing.assert_allclose(env.state, g["states"][0, t], rtol=1e-13, atol=1e-14)
        assert rew == pytest.approx(-g["costs"][0, t], rel=1e-13, abs=1e-14) and done is False
    assert env.evaluate_success([]) == 0.0 and env._max_episode_steps == 200
    env.close()


def test_example_driver_runs_the_pendulum_config():
    import os
    import subprocess
    import sys
    from conftest import ROOT
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "run_mpc.py"), "--config",
                        os.path.join(ROOT, "examples", "configs", "simple_pendulum-v0.yml"), "--controller", "cem",
                        "--n_episodes", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "Avg. reward" in r.stdout


def test_clgaussian_mpc_subclass_and_rollout_cl():
    """the reference's CLGaussianMPC surface (clgaussian_mpc.py:10-145) on the in-kernel linear
    policy, through a minimal subclass that supplies the missing update rule (a cost-weighted average of per-particle
    bias perturbations -- just enough to exercise optimize()), and GymEnvWrapper.rollout_cl's 7-tuple
    (gym_env_wrapper.py:255-325) for a linear policy, both against the CPU oracle's closed-loop restatement."""
    import torch
    from conftest import synthetic_state
    from mjmpc_b200.control import CLGaussianMPC
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    from mjmpc_b200.envs.model import compile_model, reacher7dof_spec
    from oracle import mjstep
    cm = compile_model(reacher7dof_spec())
    om = mjstep.OracleModel(cm.tree)
    env = GpuReacherVecEnv(cm)
    K, H = 128, 8
    W0 = _linear_policy(2)

    class BiasCL(CLGaussianMPC):
        def _update_distribution(self, traj):
            costs = traj["costs"].sum(dim=1)
            w = torch.softmax(-(costs - costs.min()) / 0.5, dim=0)
            delta = (traj["actions"][:, 0] * w[:, None]).sum(0) - traj["actions"][:, 0].mean(0)
            self._weights[-1] += 0.1 * delta                     # nudge the bias row only

    c = BiasCL(d_state=25, d_obs=20, d_action=7, action_lows=-np.ones(7), action_highs=np.ones(7), horizon=H, init_cov=0.3,
               init_mean=W0, num_particles=K, gamma=0.99, n_iters=1, filter_coeffs=[0.25, 0.8, 0.0], seed=3)
    c.set_sim_state_fn, c.rollout_fn = env.set_env_state, env.rollout_fn
    st = synthetic_state(cm, 6)
    traj = c.generate_rollouts(st)
    noise = c.sample_noise().cpu().numpy()
    ref = mjstep.rollout(om, st["qp"], st["qv"], st["target_pos"], None, np.ascontiguousarray(noise), want_obs=True,
                         nthreads=4, policy_w=W0)
    np.testing.assert_allclose(traj["costs"].cpu().numpy(), ref["costs"], rtol=1e-8)
    np.testing.assert_allclose(traj["actions"].cpu().numpy(), ref["actions"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(traj["observations"][:, 1:].cpu().numpy(), ref["next_observations"][:, :-1], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(c.curr_obs[:14], np.concatenate([st["qp"], st["qv"]]))
    action, value = c.optimize(st)
    want = c.mean_weights.T @ np.append(c.curr_obs, 1.0)             # clgaussian_mpc.py:63
    np.testing.assert_allclose(action, want, rtol=1e-12)
    assert c.num_steps == 1 and not np.array_equal(c.mean_weights[-1], W0[-1]) and np.array_equal(c.mean_weights[:-1], W0[:-1])
    with pytest.raises(NotImplementedError):
        c.get_optimal_value(st)                                     # _calc_val raises, as in the reference
    c.reset()
    np.testing.assert_array_equal(c.mean_weights, W0)
    with pytest.raises(TypeError):
        CLGaussianMPC(d_state=25, d_obs=20, d_action=7, action_lows=None, action_highs=None, horizon=H, init_cov=0.3,
                      init_mean=W0, num_particles=K, gamma=0.99, n_iters=1, filter_coeffs=[1, 0, 0])   # abstract, as in the reference

    # rollout_cl: the reference's closed-loop rollout signature for a linear policy
    class LinearPolicy:
        weights = W0
    env.set_env_state(st)
    obs, act, act_infos, rew, done, next_obs, info = env.rollout_cl(LinearPolicy(), K, H, mode='mean', noise=noise)
    np.testing.assert_allclose(rew, -ref["costs"], rtol=1e-8)
    np.testing.assert_allclose(act, ref["actions"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(next_obs, ref["next_observations"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(obs[:, 1:], next_obs[:, :-1])
    assert obs.shape == (K, H, 20) and done.shape == (K, H) and "total_time" in info
    with pytest.raises(NotImplementedError):
        env.rollout_cl(object(), K, H)
    env.close()
