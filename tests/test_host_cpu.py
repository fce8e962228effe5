"""CPU-only checks (-m "not gpu"): the model compiler, the oracle's physics, the product's device math
compiled for the host against the oracle, and that the C-ABI library loads and exports every symbol the
header declares."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, reference_noise, synthetic_state


# ---- model compiler -----------------------------------------------------------------------------------
def test_model_matches_survey_data_sheet(compiled_model):
    """SURVEY Appendix A (derived independently at survey time from sawyer.xml)."""
    t = compiled_model.tree
    np.testing.assert_allclose(t.mass, [24.3117, 10.4720, 0.2848, 5.4287, 1.3572, 0.2848, 2.8091, 0.0168, 2.1447], atol=6e-5)
    np.testing.assert_allclose(t.ipos[0], [0, 0.00299, -0.08429], atol=1e-5)
    np.testing.assert_allclose(t.ipos[3], [0.2, 0, 0], atol=1e-12)
    np.testing.assert_allclose(t.ipos[6], [0.1455, 0, 0], atol=1e-12)
    np.testing.assert_allclose(np.diag(t.inertia[1]), [0.126711, 0.048171, 0.126711], atol=1e-6)
    np.testing.assert_allclose(np.diag(t.inertia[3]), [0.009446, 0.110027, 0.110027], atol=1e-6)
    np.testing.assert_allclose(np.diag(t.inertia[8]), [0.005490] * 3, atol=1e-6)
    assert compiled_model.chain.axes == (2, 1, 0, 1, 0, 1, 0)
    np.testing.assert_allclose(t.gear, [20, 10, 10, 10, 10, 10, 10])
    np.testing.assert_allclose(t.damping, [2, 2, .8, .8, .8, .8, .8])


@pytest.mark.skipif(not os.path.exists("/root/reference/mjmpc/envs/assets/xml/sawyer.xml"),
                    reason="reference tree not mounted (GPU box)")
def test_builtin_spec_equals_reference_mjcf(compiled_model):
    """The built-in reacher spec is pinned to the reference's MJCF: parsing sawyer.xml gives the same block."""
    from mjmpc_b200.envs.mjcf import load_mjcf
    from mjmpc_b200.envs.model import compile_model
    ref = compile_model(load_mjcf("/root/reference/mjmpc/envs/assets/xml/sawyer.xml"))
    np.testing.assert_array_equal(ref.chain.params, compiled_model.chain.params)
    assert ref.chain.axes == compiled_model.chain.axes
    np.testing.assert_array_equal(ref.tree.dof_invweight0, compiled_model.tree.dof_invweight0)


def test_param_layout_matches_header():
    """The Python CH_* offsets mirror csrc/chain_model.h."""
    from mjmpc_b200.envs import model as M
    from mjmpc_b200 import _lib
    src = r'''
    #include <stdio.h>
    #include "mjmpc_b200/csrc/chain_model.h"
    #include "include/mjmpc_b200.h"
    int main(){printf("%d %d %d %d %d %d %d %d %d %d %d %d %d %d %d %d %d\n", CH_OFF, CH_MASS, CH_COM, CH_INERTIA,
      CH_ARMATURE, CH_DAMPING, CH_GEAR, CH_CTRL_LO, CH_CTRL_HI, CH_RANGE_LO, CH_RANGE_HI, CH_INVW0, CH_SCALARS,
      CH_NPARAM, MJB_MODEL_NPARAM, MJB_STATE_DIM, MJB_OBS_DIM); return 0;}'''
    exe = os.path.join(ROOT, "tests", "hostcheck", "layout_probe")
    subprocess.run(["gcc", "-x", "c", "-", "-I", ROOT, "-o", exe], input=src.encode(), cwd=ROOT, check=True)
    vals = [int(x) for x in subprocess.check_output([exe]).split()]
    want = [M.CH_OFF, M.CH_MASS, M.CH_COM, M.CH_INERTIA, M.CH_ARMATURE, M.CH_DAMPING, M.CH_GEAR, M.CH_CTRL_LO,
            M.CH_CTRL_HI, M.CH_RANGE_LO, M.CH_RANGE_HI, M.CH_INVW0, M.CH_SCALARS, M.CH_NPARAM, _lib.MODEL_NPARAM,
            _lib.STATE_DIM, _lib.OBS_DIM]
    assert vals == want


def test_randomized_copy_follows_reference_rule(compiled_model):
    """gym_env_wrapper.py:408-410: U(biased*(1-noise), biased*(1+noise)), defaults cached, repeated calls
    re-randomise around the original."""
    from mjmpc_b200.envs.model import randomized_copy
    rng = np.random.RandomState(3)
    pd = dict(body_mass={"r_forearm_link": [0.3, 0.1]}, dof_damping={"r_elbow_flex_joint": [0.0, -0.5]})
    m1, defaults, rnd = randomized_copy(compiled_model, pd, rng)
    base = compiled_model.tree.mass[6]
    assert defaults["body_mass"]["r_forearm_link"] == base
    assert 1.1 * base * 0.7 <= rnd["body_mass"]["r_forearm_link"] <= 1.1 * base * 1.3
    assert rnd["dof_damping"]["r_elbow_flex_joint"] == pytest.approx(0.4)
    m2, defaults2, rnd2 = randomized_copy(m1, pd, rng, defaults)
    assert defaults2["body_mass"]["r_forearm_link"] == base               # not overwritten
    assert 1.1 * base * 0.7 <= rnd2["body_mass"]["r_forearm_link"] <= 1.1 * base * 1.3
    np.testing.assert_array_equal(m1.tree.dof_invweight0, compiled_model.tree.dof_invweight0)
    with pytest.raises(ValueError):
        randomized_copy(compiled_model, dict(bogus={"x": [0.1, 0.1]}), rng)


# ---- oracle physics (no MuJoCo available: two independent derivations + invariants) ----------------------
def test_oracle_matches_lagrangian_derivation(compiled_model, oracle_model):
    from oracle import lagrange
    rng = np.random.default_rng(1)
    for _ in range(4):
        q = rng.uniform(-2, 2, 7); v = rng.normal(0, 3, 7)
        M, b, _ = oracle_model.mass_bias(q, v)
        M2, c2 = lagrange.mass_bias(compiled_model.tree, q, v)
        np.testing.assert_allclose(M, M2, atol=1e-13)
        np.testing.assert_allclose(b, c2, atol=1e-12 * max(1.0, np.abs(c2).max()))
        assert np.linalg.eigvalsh(M).min() > 0
        np.testing.assert_allclose(M, M.T, atol=0)


def test_oracle_energy_conservation(compiled_model):
    """damping=0, ctrl=0, gravity=0, limits off: kinetic energy is conserved to O(h)."""
    import copy
    from oracle import mjstep
    tree = copy.deepcopy(compiled_model.tree)
    tree.damping[:] = 0.0
    tree.jnt_limited[:] = 0
    tree.con_radius = -1.0
    om = mjstep.OracleModel(tree)
    q = np.array([0.1, 0.2, -0.3, -0.5, 0.4, -0.2, 0.3]); v = np.array([0.5, -0.4, 0.8, 0.6, -0.7, 0.9, -0.5])
    M, _, _ = om.mass_bias(q, v)
    e0 = 0.5 * v @ M @ v
    for _ in range(50):
        q, v, _, n = om.substep(q, v, np.zeros(7))
        assert n == 0
    M, _, _ = om.mass_bias(q, v)
    assert abs(0.5 * v @ M @ v - e0) / e0 < 0.02


def test_oracle_limits_and_contact_hold(compiled_model, oracle_model):
    from oracle import mjstep
    K, H = 256, 32
    noise = reference_noise(K, H, 7, 1)
    out = mjstep.rollout(oracle_model, np.zeros(7), np.zeros(7), np.array([.1, .1, .1]), np.zeros((H, 7)), noise,
                         want_traj=True, nthreads=4)
    q = out["qv"][:, :, :7]
    lo, hi = compiled_model.tree.jnt_range[:, 0], compiled_model.tree.jnt_range[:, 1]
    assert (q - hi).max() < 0.3 and (lo - q).max() < 0.3                   # soft limits: bounded violation
    assert (out["ncon"] > 0).mean() > 0.5
    # pushing down onto the table: the soft contact stops the sphere near the plane
    mean = np.zeros((H, 7)); mean[:, 1] = 1.0
    out = mjstep.rollout(oracle_model, np.array([0.0, 0.45, 0, -0.2, 0, -0.3, 0.0]), np.zeros(7),
                         np.array([.1, .1, .1]), mean, 0.3 * noise, want_obs=True)
    hz = out["next_observations"][:, :, 16]
    assert hz.min() < -0.345 and hz.min() > -0.40


# ---- the product's device math, compiled for the host, against the oracle ---------------------------------
@pytest.fixture(scope="module")
def hostcheck():
    so = os.path.join(ROOT, "tests", "hostcheck", "libhostcheck.so")
    src = os.path.join(ROOT, "tests", "hostcheck", "hostcheck.cpp")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src, "-lm"])
    return C.CDLL(so)


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.mark.parametrize("case", ["interior", "reset", "table"])
def test_device_math_on_host_matches_oracle(case, compiled_model, oracle_model, hostcheck):
    from oracle import mjstep
    P = compiled_model.chain.params
    assert hostcheck.hostcheck_fits_sawyer(_p(P)) == 1
    K, H = 192, 32
    noise = reference_noise(K, H, 7, 5)
    mean = np.zeros((H, 7))
    if case == "interior":
        st = synthetic_state(compiled_model, 3)
        q0, v0 = st["qp"], st["qv"]
    elif case == "reset":
        q0, v0 = np.zeros(7), np.zeros(7)
    else:
        q0, v0 = np.array([0.0, 0.45, 0, -0.2, 0, -0.3, 0.0]), np.zeros(7)
        mean[:, 1] = 1.0
        noise = 0.3 * noise
    tgt = np.array([0.1, 0.1, 0.1])
    ref = mjstep.rollout(oracle_model, q0, v0, tgt, mean, noise, want_traj=True, nthreads=4)
    for dense in (0, 1):
        costs = np.zeros((K, H)); qv = np.zeros((K, H, 14))
        hostcheck.hostcheck_rollout(_p(P), dense, _p(q0), _p(v0), _p(tgt), K, H, _p(mean), _p(noise), _p(costs), _p(qv))
        scale = np.abs(ref["qv"]).max(axis=(0, 1))
        assert (np.abs(qv - ref["qv"]).max(axis=(0, 1)) / scale).max() < 1e-9
        np.testing.assert_allclose(costs, ref["costs"], rtol=1e-10)


# ---- C ABI ------------------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    from mjmpc_b200 import _lib, build
    build.build()
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "mjmpc_b200.h")).read()
    declared = set(re.findall(r"\b(mjb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert declared == set(_lib.EXPORTS)
    assert L.mjb_version() >= 100
    assert L.mjb_softmax_partial_doubles(32, 7, 0, 0) == 1 + 32 * 8
    assert L.mjb_softmax_partial_doubles(32, 7, 1, 0) == 32 + 32 * 8
    assert L.mjb_softmax_partial_doubles(16, 7, 0, 2) == 1 + 16 * (8 + 28)


def test_product_never_imports_the_oracle():
    for dp, _, files in os.walk(os.path.join(ROOT, "mjmpc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f
                # ... nor the host emulation the tests run it on
                assert "emu_device" not in txt and "MJB_TEST_EMU" not in txt and "b200_emu" not in txt, f


def test_sincos_joint_accuracy(hostcheck):
    """The compact sincos used for the (bounded) joint angles: < 2 ulp-ish over far more than the joint ranges."""
    xs = np.concatenate([np.linspace(-12, 12, 20001), np.random.default_rng(0).uniform(-300, 300, 5000),
                         [0.0, np.pi / 4, np.pi / 2, -np.pi / 2, np.pi, 1e-9, -1e-300]])
    s, c = C.c_double(), C.c_double()
    hostcheck.hostcheck_sincos.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    err = 0.0
    for x in xs:
        hostcheck.hostcheck_sincos(float(x), C.byref(s), C.byref(c))
        err = max(err, abs(s.value - np.sin(x)), abs(c.value - np.cos(x)))
    assert err < 4e-16


# ---- driver / config glue --------------------------------------------------------------------------------
def test_example_config_loads_like_the_reference_file():
    import yaml
    mine = yaml.safe_load(open(os.path.join(ROOT, "examples", "configs", "reacher_7dof-v0.yml")))
    for name in ("mppi", "cem", "dmd", "pfmpc", "random_shooting"):
        assert name in mine and mine[name]["horizon"] == 16 and mine[name]["num_cpu"] * mine[name]["particles_per_cpu"] == 32
    ref_path = "/root/reference/examples/configs/reacher_7dof-v0.yml"
    if os.path.exists(ref_path):
        ref = yaml.safe_load(open(ref_path))
        for name in ("mppi", "cem", "dmd", "pfmpc"):       # the reference's random_shooting block has a malformed init_cov
            assert mine[name] == ref[name], name
        for key in ("env_name", "n_episodes", "max_ep_length", "seed", "base_action"):
            assert mine[key] == ref[key]


def test_load_policy_params_follows_the_reference_driver():
    """examples/example_mpc.py:71-79,135-136: env dims injected, num_particles = num_cpu * particles_per_cpu,
    both popped; plus the top-level base_action fallback."""
    import sys
    import types
    import yaml
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    from run_mpc import load_policy_params
    exp = yaml.safe_load(open(os.path.join(ROOT, "examples", "configs", "reacher_7dof-v0.yml")))
    env = types.SimpleNamespace(d_obs=20, d_state=25, d_action=7, action_lows=-np.ones(7), action_highs=np.ones(7))
    p, num_cpu = load_policy_params(exp, "mppi", env)
    assert num_cpu == 8 and p["num_particles"] == 32 and "num_cpu" not in p and "particles_per_cpu" not in p
    assert p["base_action"] == "null" and p["d_action"] == 7 and p["lam"] == 0.2
    assert "num_cpu" in exp["mppi"]              # the loaded YAML itself is left untouched


def test_mjcf_reader_rejects_unsupported_models(tmp_path):
    from mjmpc_b200.envs.mjcf import load_mjcf
    bad = tmp_path / "bad.xml"
    bad.write_text('<mujoco><compiler inertiafromgeom="true" angle="radian" coordinate="local"/>'
                   '<option timestep="0.01" gravity="0 0 -9.81"/><worldbody/></mujoco>')
    with pytest.raises(ValueError):
        load_mjcf(str(bad))
    bad.write_text('<mujoco><compiler inertiafromgeom="false" angle="radian"/><option/><worldbody/></mujoco>')
    with pytest.raises(ValueError):
        load_mjcf(str(bad))


def test_controllers_refuse_to_run_without_cuda():
    """No CPU fallback: on a machine without a GPU the product path raises instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mjmpc_b200 import _lib
    from mjmpc_b200.control import MPPI
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    with pytest.raises(_lib.MjbError):
        GpuReacherVecEnv()
    with pytest.raises(_lib.MjbError):
        MPPI(d_state=25, d_obs=20, d_action=7, horizon=4, init_cov=1.0, base_action='null', lam=0.2, num_particles=8,
             step_size=1.0, alpha=1, gamma=1.0, n_iters=1, action_lows=-np.ones(7), action_highs=np.ones(7))


def test_ctypes_structs_match_header_layout():
    """sizeof and the offset of the last field of every argument struct, C compiler vs ctypes mirror."""
    import ctypes
    from mjmpc_b200 import _lib
    pairs = [("mjb_rollout_args", _lib.RolloutArgs, "closed_loop"), ("mjb_pendulum_args", _lib.PendulumArgs, "states_out"),
             ("mjb_noise_args", _lib.NoiseArgs, "out_sj"), ("mjb_softmax_args", _lib.SoftmaxArgs, "q_st"),
             ("mjb_combine_args", _lib.CombineArgs, "stats"), ("mjb_elite_args", _lib.EliteArgs, "partial"),
             ("mjb_elite_combine_args", _lib.EliteCombineArgs, "cov"), ("mjb_mppi_batched_args", _lib.MppiBatchedArgs, "value"),
             ("mjb_pf_batched_args", _lib.PfBatchedArgs, "mean"), ("mjb_mpc_step_args", _lib.MpcStepArgs, "cov_shift_beta"),
             ("mjb_lqr_args", _lib.LqrArgs, "states_out")]
    body = "".join('printf("%%zu %%zu\\n", sizeof(%s), offsetof(%s, %s));' % (c, c, last) for c, _, last in pairs)
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "include/mjmpc_b200.h"\nint main(){%s return 0;}' % body
    exe = os.path.join(ROOT, "tests", "hostcheck", "layout_probe")
    subprocess.run(["gcc", "-x", "c", "-", "-I", ROOT, "-o", exe], input=src.encode(), cwd=ROOT, check=True)
    lines = subprocess.check_output([exe]).decode().split("\n")
    for (cname, cls, last), line in zip(pairs, lines):
        size, off = [int(x) for x in line.split()]
        assert ctypes.sizeof(cls) == size, cname
        assert getattr(cls, last).offset == off, cname


def test_c_abi_argument_validation_maps_to_reference_exceptions():
    """Every entry point validates shapes before touching CUDA, returns MJB_EINVAL with a message, and the
    ctypes shim raises the exception type the reference raises for the same mistake (ValueError / the
    `K % num_cpu` assertion text of subproc_vec_env.py:162).  No compute call is made: the checks fail first."""
    import ctypes
    from mjmpc_b200 import _lib
    L = _lib.lib()
    D = 8            # any non-null address: never dereferenced, the shape checks fail first

    def expect(rc, text):
        assert rc == _lib.MJB_EINVAL
        msg = L.mjb_last_error().decode()
        assert text in msg, msg
        with pytest.raises(ValueError):
            _lib.check(rc)

    a = _lib.RolloutArgs()
    expect(L.mjb_rollout_reacher(None, ctypes.byref(a), None), "null handle")
    n = _lib.NoiseArgs()
    expect(L.mjb_generate_noise(ctypes.byref(n), None), "null pointer")
    n.K, n.H, n.d, n.cov, n.out = 4, 4, 9, D, D
    expect(L.mjb_generate_noise(ctypes.byref(n), None), "not in 1..8")
    p = _lib.PfBatchedArgs()
    p.n_ctrl, p.K, p.H, p.d = 1, 5000, 4, 7
    p.costs = p.samples = p.gamma_seq = p.r = p.out = p.mean = D
    expect(L.mjb_pf_update_batched(ctypes.byref(p), None), "<= 4096")
    m = _lib.MppiBatchedArgs()
    m.n_ctrl, m.K, m.H, m.d, m.lam = 1, 8, 4, 7, 0.0
    m.costs = m.actions = m.mean = m.gamma_seq = D
    expect(L.mjb_mppi_update_batched(ctypes.byref(m), None), "lam must be positive")
    g = np.ones(200)
    rc = L.mjb_cost_to_go(ctypes.c_void_p(D), 1, 1, g.ctypes.data_as(ctypes.c_void_p), 4, 200, ctypes.c_void_p(D), 1, 1, None)
    expect(rc, "exceeds the supported maximum")
    e = _lib.EliteArgs()
    assert L.mjb_elite_moments1(ctypes.byref(e), None) == _lib.MJB_EINVAL
    assert L.mjb_select_elites(None, 8, 2, None, None, None, None) == _lib.MJB_EINVAL
    assert L.mjb_shift_mean(None, 4, 7, 0, None, None) == _lib.MJB_EINVAL
    assert _lib.check(_lib.MJB_OK) is None
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.MJB_ENOTIMPL)


def test_continual_reacher_timed_events_rule():
    """reacher_env.py:128-132 on the host class (no device needed for the rule itself)."""
    from mjmpc_b200.envs.gpu_reacher_env import GpuContinualReacherEnv, GpuReacherEnv
    env = object.__new__(GpuContinualReacherEnv)
    env.np_random = np.random.RandomState(3)
    env.real_step, env.target_pos = True, np.array([0.1, 0.1, 0.1])
    fired = []
    for t in range(0, 151):
        env.env_timestep = t
        before = env.target_pos.copy()
        env.trigger_timed_events()
        if not np.array_equal(before, env.target_pos):
            fired.append(t)
    assert fired == [50, 100, 150]
    rng = np.random.RandomState(3)
    for _ in range(3):
        want = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.2, 0.2), rng.uniform(-0.25, 0.25)])
    np.testing.assert_array_equal(env.target_pos, want)
    env.real_step = False
    env.env_timestep = 200
    env.trigger_timed_events()
    np.testing.assert_array_equal(env.target_pos, want)
    base = object.__new__(GpuReacherEnv)
    base.env_timestep, base.real_step, base.target_pos = 50, True, want.copy()
    base.trigger_timed_events()
    np.testing.assert_array_equal(base.target_pos, want)


def test_rank_one_repair_settles_misjudged_rows_on_the_host(compiled_model, hostcheck):
    """The solver's first active-set guess is wrong for a few percent of the constrained substeps, almost always
    in one row; the rank-one repair (chain_dynamics.cuh) must settle those without a further factor/solve trip:
    nearly every repair confirmed, a third trip the exception even at warp level (max over 32 lanes)."""
    P = compiled_model.chain.params
    K, H = 512, 24
    st = synthetic_state(compiled_model, 1)
    noise = reference_noise(K, H, 7, 11)
    mean = np.zeros((H, 7))
    trips = np.zeros(K * H * 2, dtype=np.int32)
    stats = (C.c_longlong * 6)()
    hostcheck.hostcheck_stats(stats, 1)
    hostcheck.hostcheck_record_trips(trips.ctypes.data_as(C.POINTER(C.c_int)))
    costs = np.zeros((K, H)); qv = np.zeros((K, H, 14))
    hostcheck.hostcheck_rollout(_p(P), 0, _p(st["qp"]), _p(st["qv"]), _p(st["target_pos"]), K, H, _p(mean), _p(noise),
                                _p(costs), _p(qv))
    hostcheck.hostcheck_record_trips(None)
    hostcheck.hostcheck_stats(stats, 0)
    substeps, with_rows, passes, slow, tried, confirmed = list(stats)
    assert substeps == K * H * 2 and with_rows > 0.2 * substeps          # the workload does exercise the limits
    assert tried > 0.01 * with_rows and confirmed >= 0.99 * tried
    assert slow == 0
    t = trips.reshape(K, H * 2)
    assert t.min() == 1 and t.max() <= 4
    warp = t.reshape(K // 32, 32, H * 2).max(axis=1)
    assert (warp >= 3).mean() < 0.10, (warp >= 3).mean()
    assert np.isfinite(qv).all()


# ---- closed-loop linear rollouts (gym_env_wrapper.py:129-136) --------------------------------------------------
def _linear_policy(seed, scale=0.1):
    """Random feedback gains, small enough that the closed loop does not amplify rounding differences between
    two formulations of the dynamics by more than a few orders of magnitude over the horizon."""
    rng = np.random.default_rng(seed)
    W = scale * rng.normal(0, 1, (21, 7))
    W[14:20] *= 2.0             # the hand / hand-target rows matter for reaching
    return W


def test_oracle_closed_loop_follows_the_reference_loop(compiled_model, oracle_model):
    """ora_rollout_cl against the reference's loop written out in Python on top of the oracle's own single
    steps: curr_obs = get_obs() at the set state (fresh kinematics), u = mean.T @ [obs; 1] + noise,
    curr_obs = next_obs (stale hand position of the step's last forward pass)."""
    from oracle import mjstep
    K, H = 6, 10
    st = synthetic_state(compiled_model, 2)
    W = _linear_policy(0)
    noise = 0.5 * reference_noise(K, H, 7, 3)
    out = mjstep.rollout(oracle_model, st["qp"], st["qv"], st["target_pos"], None, noise, want_obs=True, policy_w=W)
    tgt = st["target_pos"]
    for k in range(K):
        q, v = st["qp"].copy(), st["qv"].copy()
        hand = oracle_model.mass_bias(q, v)[2]
        obs = np.concatenate([q, v, hand, hand - tgt])
        for t in range(H):
            u = W.T @ np.append(obs, 1.0) + noise[k, t]
            # BLAS and the C loop sum W'obs in different orders; the feedback loop amplifies the last-bit difference
            np.testing.assert_allclose(out["actions"][k, t], u, rtol=1e-9, atol=1e-10)
            for s in range(2):
                hand = oracle_model.mass_bias(q, v)[2]          # kinematics before this substep's integration
                q, v, _, _ = oracle_model.substep(q, v, u)
            obs = np.concatenate([q, v, hand, hand - tgt])
            np.testing.assert_allclose(out["next_observations"][k, t], obs, rtol=1e-9, atol=1e-10)
            d = hand - tgt
            np.testing.assert_allclose(out["costs"][k, t], np.abs(d).sum() + 5 * np.linalg.norm(d), rtol=1e-9)
    # a policy with only the bias row is an open-loop rollout with a constant mean
    Wb = np.zeros((21, 7)); Wb[20] = [0.3, -0.2, 0.1, 0.0, 0.2, -0.1, 0.05]
    a = mjstep.rollout(oracle_model, st["qp"], st["qv"], tgt, None, noise, policy_w=Wb)
    b = mjstep.rollout(oracle_model, st["qp"], st["qv"], tgt, np.tile(Wb[20], (H, 1)), noise)
    np.testing.assert_array_equal(a["costs"], b["costs"])


@pytest.mark.parametrize("case", ["interior", "reset"])
def test_closed_loop_device_math_on_host_matches_oracle(case, compiled_model, oracle_model, hostcheck):
    from oracle import mjstep
    P = compiled_model.chain.params
    K, H = 96, 24
    st = synthetic_state(compiled_model, 5)
    q0, v0 = (st["qp"], st["qv"]) if case == "interior" else (np.zeros(7), np.zeros(7))
    tgt = st["target_pos"]
    W = _linear_policy(1)
    noise = np.ascontiguousarray(0.5 * reference_noise(K, H, 7, 9))
    ref = mjstep.rollout(oracle_model, q0, v0, tgt, None, noise, want_traj=True, nthreads=4, policy_w=W)
    costs = np.zeros((K, H)); qv = np.zeros((K, H, 14)); act = np.zeros((K, H, 7))
    hostcheck.hostcheck_rollout_cl(_p(P), _p(np.ascontiguousarray(q0)), _p(np.ascontiguousarray(v0)), _p(tgt), K, H,
                                   _p(np.ascontiguousarray(W)), _p(noise), _p(costs), _p(qv), _p(act))
    # north-star tolerance: 1e-8 relative over the horizon (the feedback loop amplifies the ~1e-13 per-step
    # difference between the link-frame and the world-frame formulation)
    scale = np.abs(ref["qv"]).max(axis=(0, 1))
    assert (np.abs(qv - ref["qv"]).max(axis=(0, 1)) / scale).max() < 1e-8
    np.testing.assert_allclose(costs, ref["costs"], rtol=1e-8)
    np.testing.assert_allclose(act, ref["actions"], rtol=1e-8, atol=1e-9)
