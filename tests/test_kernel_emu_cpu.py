"""The actual kernel SOURCES of K2 (noise.cu) and K1 (rollout_reacher.cu) executed on the host, one simulated
thread after another (tests/hostcheck/kernel_emu.cpp), through the same argument structs the C ABI takes.
These kernels have no inter-thread communication, so this runs the real code paths -- strides, the noise
prefetch, fused noise, trajectory / observation outputs, the closed-loop policy, per-worker models -- without a
GPU; only the special-function-unit approximations are replaced by libm.  The GPU suite repeats the same
checks on the device."""
import ctypes as C
import os
import sys
import subprocess

import numpy as np
import pytest

from conftest import ROOT, reference_noise, synthetic_state
from mjmpc_b200 import _lib


@pytest.fixture(scope="module")
def emu():
    d = os.path.join(ROOT, "tests", "hostcheck")
    so = os.path.join(d, "libkernel_emu.so")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I", cuda_inc, "-o", so,
                           os.path.join(d, "kernel_emu.cpp"), "-lm"])
    L = C.CDLL(so)
    # K2 is a cooperating-thread kernel: it runs on the block emulator build of the whole C ABI
    sys.path.insert(0, d)
    import gen_lib_emu
    L.block_emu = C.CDLL(gen_lib_emu.build())
    return L


def _vp(a):
    return a.ctypes.data


class _Guard:
    """Every array handed to the kernel sources sits between bands of NaN sentinels: an out-of-bounds write
    destroys a sentinel (check()), an out-of-bounds read pulls a NaN into a result the tests compare."""
    bufs = []

    @classmethod
    def full(cls, shape, fill, dtype=np.float64, pad=128):
        n = int(np.prod(shape))
        sentinel = np.nan if np.dtype(dtype).kind == "f" else np.iinfo(dtype).min + 7
        flat = np.full(n + 2 * pad, sentinel, dtype=dtype)
        view = flat[pad:pad + n].reshape(shape)
        view[...] = fill
        cls.bufs.append((flat, pad, n, sentinel))
        return view

    @classmethod
    def copy(cls, a):
        a = np.ascontiguousarray(a, np.float64)
        return cls.full(a.shape, a)

    @classmethod
    def check(cls):
        for flat, pad, n, sentinel in cls.bufs:
            band = np.concatenate([flat[:pad], flat[pad + n:]])
            assert (np.isnan(band).all() if flat.dtype.kind == "f" else (band == sentinel).all()), "out-of-bounds write"
        cls.bufs.clear()


def _noise(emu, cov, b, K, H, seed, step=0, k_offset=0, K_global=None, zero_last_mean=None, layout="row"):
    cov = np.ascontiguousarray(cov, np.float64)
    d = cov.shape[0]
    a = _lib.NoiseArgs()
    a.K, a.H, a.d = K, H, d
    a.k_offset, a.K_global = k_offset, K if K_global is None else K_global
    a.seed, a.offset = seed, step
    a.cov, a.beta0, a.beta1, a.beta2 = _vp(cov), b[0], b[1], b[2]
    if zero_last_mean is not None:
        zl = np.ascontiguousarray(zero_last_mean, np.float64)
        a.zero_last, a.neg_mean = 1, _vp(zl)
    if layout == "row":
        out = _Guard.full((K, H, d), np.nan)
        a.out_sk, a.out_st, a.out_sj = H * d, d, 1
        view = out
    else:                                   # the product's particle-minor layout (H, d, K) viewed as (K, H, d)
        out = _Guard.full((H, d, K), np.nan)
        a.out_sk, a.out_st, a.out_sj = 1, d * K, K
        view = out.transpose(2, 0, 1)
    a.out = _vp(out)
    assert emu.block_emu.mjb_generate_noise(C.byref(a), None) == 0
    _Guard.check()
    return view


def test_noise_kernel_moments_filter_and_shape(emu):
    K, H, d = 120000, 8, 7
    b = (0.25, 0.8, 0.1)
    cov = np.diag([1.0, 0.5, 2.0, 1.0, 0.1, 3.0, 1.0])
    eps = _noise(emu, cov, b, K, H, 1234, step=5)
    assert np.isfinite(eps).all()
    for t in (0, 1):
        assert np.all(np.abs(eps[:, t].mean(0)) < 5 * np.sqrt(np.diag(cov) / K))
        np.testing.assert_allclose(eps[:, t].var(0), np.diag(cov), rtol=0.03)
    z = (eps[:, 2:] - b[1] * eps[:, 1:-1] - b[2] * eps[:, :-2]) / b[0]
    np.testing.assert_allclose(z.var(axis=(0, 1)), np.diag(cov), rtol=0.02)
    assert abs(np.corrcoef(z[:, 0, 0], z[:, 1, 0])[0, 1]) < 0.015
    assert abs(np.corrcoef(z[:, 3, 0], z[:, 3, 1])[0, 1]) < 0.015
    x = eps[:, 0, 0]
    assert abs(np.mean(x ** 4) / np.mean(x ** 2) ** 2 - 3.0) < 0.06
    assert abs(np.mean(np.abs(x) > 1.959964) - 0.05) < 0.003
    assert np.abs(eps[:, 0, 0]).max() < 6.8                       # 32-bit radius uniform


def test_noise_kernel_full_covariance_determinism_shards_layouts(emu):
    rng = np.random.default_rng(0)
    A = rng.normal(0, 1, (7, 7))
    cov = A @ A.T / 7 + 0.1 * np.eye(7)
    eps = _noise(emu, cov, (1.0, 0.0, 0.0), 100000, 2, 7)
    np.testing.assert_allclose(np.cov(eps[:, 0].T), cov, atol=0.04)
    K, H = 1024, 6
    bf = (0.25, 0.8, 0.0)
    full = _noise(emu, np.eye(7) * 0.7, bf, K, H, 11, step=1)
    np.testing.assert_array_equal(full, _noise(emu, np.eye(7) * 0.7, bf, K, H, 11, step=1))
    assert not np.array_equal(full, _noise(emu, np.eye(7) * 0.7, bf, K, H, 11, step=2))
    assert not np.array_equal(full, _noise(emu, np.eye(7) * 0.7, bf, K, H, 12, step=1))
    parts = [_noise(emu, np.eye(7) * 0.7, bf, K // 4, H, 11, step=1, k_offset=r * K // 4, K_global=K) for r in range(4)]
    np.testing.assert_array_equal(full, np.concatenate(parts, 0))
    np.testing.assert_array_equal(full, _noise(emu, np.eye(7) * 0.7, bf, K, H, 11, step=1, layout="particle_minor"))
    mean = rng.normal(0, 1, (H, 7))
    zl = _noise(emu, np.eye(7) * 0.7, bf, K, H, 11, step=1, zero_last_mean=mean)
    np.testing.assert_array_equal(zl[-1], -mean)
    np.testing.assert_array_equal(zl[:-1], full[:-1])
    one = _noise(emu, np.array([[3.0]]), (0.6, 0.5, 0.0), 4096, 16, 0)
    assert one.shape == (4096, 16, 1) and abs(one[:, 0, 0].var() - 3.0) < 0.3


def _rollout(emu, P, st, K, H, mean, noise, n_inst=1, traj=False, obs=False, ncon=False, closed=False, fused=None,
             particle_minor=False):
    a = _lib.RolloutArgs()
    a.K, a.H, a.particles_per_ctrl, a.particles_per_model = K, H, K, K // n_inst
    state = _Guard.copy(np.concatenate([st["qp"], st["qv"], st["target_pos"]]))
    mean = _Guard.copy(mean)
    a.state, a.mean = _vp(state), _vp(mean)
    keep = [state, mean]
    if noise is not None:
        noise = _Guard.copy(noise)
        a.noise = _vp(noise)
        a.noise_sk, a.noise_st, a.noise_sj = H * 7, 7, 1
        keep.append(noise)
    if particle_minor:
        costs = _Guard.full((H, K), np.nan); actions = _Guard.full((H, 7, K), np.nan)
        a.costs_sk, a.costs_st = 1, K
        a.act_sk, a.act_st, a.act_sj = 1, 7 * K, K
        out = dict(costs=costs.T, actions=actions.transpose(2, 0, 1))
    else:
        costs = _Guard.full((K, H), np.nan); actions = _Guard.full((K, H, 7), np.nan)
        a.costs_sk, a.costs_st = H, 1
        a.act_sk, a.act_st, a.act_sj = H * 7, 7, 1
        out = dict(costs=costs, actions=actions)
    a.costs, a.actions = _vp(costs), _vp(actions)
    if traj:
        out["qv"] = _Guard.full((K, H, 14), np.nan); a.qv_traj = _vp(out["qv"])
    if obs:
        out["next_observations"] = _Guard.full((K, H, 20), np.nan); a.next_obs = _vp(out["next_observations"])
    if ncon:
        out["ncon"] = _Guard.full((K,), 0, np.int32); a.ncon = _vp(out["ncon"])
    a.closed_loop = 1 if closed else 0
    if fused is not None:
        cov, seed, step, b = fused
        cov = np.ascontiguousarray(cov, np.float64); keep.append(cov)
        a.noise_cov, a.noise_seed, a.noise_offset = _vp(cov), seed, step
        a.noise_beta0, a.noise_beta1, a.noise_beta2 = b
        a.noise_k_offset, a.noise_K_global = 0, K
    params = np.ascontiguousarray(np.tile(P, (n_inst, 1)))
    assert emu.emu_rollout_reacher(C.c_void_p(_vp(params)), n_inst, C.byref(a)) == 0
    _Guard.check()
    for k in ("costs", "actions", "qv", "next_observations"):
        assert k not in out or np.isfinite(out[k]).all(), k            # every output element written, no sentinel read
    return out


@pytest.mark.parametrize("case", ["interior", "reset", "table"])
def test_rollout_kernel_source_matches_oracle(case, emu, compiled_model, oracle_model):
    from oracle import mjstep
    P = compiled_model.chain.params
    K, H = 96, 16
    noise = reference_noise(K, H, 7, 5)
    mean = np.zeros((H, 7))
    st = synthetic_state(compiled_model, 3)
    if case == "reset":
        st = dict(st, qp=np.zeros(7), qv=np.zeros(7))
    elif case == "table":
        st = dict(st, qp=np.array([0.0, 0.45, 0, -0.2, 0, -0.3, 0.0]), qv=np.zeros(7))
        mean[:, 1] = 1.0
        noise = 0.3 * noise
    ref = mjstep.rollout(oracle_model, st["qp"], st["qv"], st["target_pos"], mean, noise, want_traj=True, want_obs=True,
                         nthreads=4)
    out = _rollout(emu, P, st, K, H, mean, noise, traj=True, obs=True, ncon=True)
    scale = np.abs(ref["qv"]).max(axis=(0, 1))
    assert (np.abs(out["qv"] - ref["qv"]).max(axis=(0, 1)) / scale).max() < 1e-9
    np.testing.assert_allclose(out["costs"], ref["costs"], rtol=1e-10)
    np.testing.assert_array_equal(out["actions"], mean[None] + noise)
    np.testing.assert_allclose(out["next_observations"], ref["next_observations"], rtol=1e-9, atol=1e-11)
    np.testing.assert_array_equal(out["ncon"] > 0, ref["ncon"] > 0)
    # the production instantiation (no extra outputs), the particle-minor layout the controllers use, and two
    # per-worker copies of the same model read from "global memory": bit-identical costs
    plain = _rollout(emu, P, st, K, H, mean, noise)
    np.testing.assert_array_equal(plain["costs"], out["costs"])
    pm = _rollout(emu, P, st, K, H, mean, noise, particle_minor=True)
    np.testing.assert_array_equal(pm["costs"], out["costs"])
    np.testing.assert_array_equal(pm["actions"], out["actions"])
    two = _rollout(emu, P, st, K, H, mean, noise, n_inst=2)
    np.testing.assert_array_equal(two["costs"], out["costs"])
    # K not a multiple of the block, H = 1, no noise (the mean sequence alone)
    rag = _rollout(emu, P, st, 5, 1, mean[:1], noise[:5, :1])
    np.testing.assert_array_equal(rag["costs"][:, 0], out["costs"][:5, 0])
    mo = _rollout(emu, P, st, 1, H, mean, None)
    mref = mjstep.rollout(oracle_model, st["qp"], st["qv"], st["target_pos"], mean, None)
    np.testing.assert_allclose(mo["costs"], mref["costs"], rtol=1e-10)


def test_rollout_kernel_source_closed_loop_matches_oracle(emu, compiled_model, oracle_model):
    from oracle import mjstep
    P = compiled_model.chain.params
    K, H = 64, 20
    rng = np.random.default_rng(1)
    W = 0.1 * rng.normal(0, 1, (21, 7)); W[14:20] *= 2.0
    noise = 0.5 * reference_noise(K, H, 7, 9)
    for st in (synthetic_state(compiled_model, 5), dict(synthetic_state(compiled_model, 5), qp=np.zeros(7), qv=np.zeros(7))):
        ref = mjstep.rollout(oracle_model, st["qp"], st["qv"], st["target_pos"], None, noise, want_traj=True, want_obs=True,
                             nthreads=4, policy_w=W)
        out = _rollout(emu, P, st, K, H, W, noise, traj=True, obs=True, closed=True)
        scale = np.abs(ref["qv"]).max(axis=(0, 1))
        assert (np.abs(out["qv"] - ref["qv"]).max(axis=(0, 1)) / scale).max() < 1e-8
        np.testing.assert_allclose(out["actions"], ref["actions"], rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(out["costs"], ref["costs"], rtol=1e-8)
        np.testing.assert_allclose(out["next_observations"], ref["next_observations"], rtol=1e-8, atol=1e-9)
    Wb = np.zeros((21, 7)); Wb[20] = [0.3, -0.2, 0.1, 0.0, 0.2, -0.1, 0.05]
    a = _rollout(emu, P, st, K, H, Wb, noise, closed=True)
    b = _rollout(emu, P, st, K, H, np.tile(Wb[20], (H, 1)), noise, ncon=True)
    np.testing.assert_array_equal(a["costs"], b["costs"])
    np.testing.assert_array_equal(a["actions"], b["actions"])


def test_fused_noise_equals_noise_kernel_then_rollout(emu, compiled_model):
    """The in-rollout generation draws exactly the samples of the noise kernel (same header, same counters)."""
    P = compiled_model.chain.params
    K, H = 70, 9
    st = synthetic_state(compiled_model, 7)
    cov = np.diag([1.0, 0.5, 2.0, 1.0, 0.3, 1.5, 1.0])
    b = (0.25, 0.8, 0.05)
    mean = np.random.default_rng(2).normal(0, 0.2, (H, 7))
    eps = _noise(emu, cov, b, K, H, 99, step=4)
    two = _rollout(emu, P, st, K, H, mean, eps)
    fused = _rollout(emu, P, st, K, H, mean, None, fused=(cov, 99, 4, b))
    np.testing.assert_array_equal(fused["actions"], two["actions"])
    np.testing.assert_array_equal(fused["costs"], two["costs"])


def test_rollout_kernel_source_on_random_hard_states(emu, compiled_model):
    """Randomised start states the bench never visits -- joints up to 0.2 rad beyond their limits, joint speeds up
    to ~6 rad/s, torque-saturating noise, the arm driven towards the table, per-case randomised masses / inertias /
    damping -- through the kernel source against the C oracle: every constraint path (several limit rows at once,
    wrong first guesses, rank-one repairs, the contact row, the out-of-line robust solver) within the north-star
    1e-8 on the whole trajectory."""
    from hypothesis import HealthCheck, assume, given, settings
    from hypothesis import strategies as hst
    from mjmpc_b200.envs.model import randomized_copy, table_clearance
    from oracle import mjstep
    lo, hi = compiled_model.tree.jnt_range[:, 0], compiled_model.tree.jnt_range[:, 1]
    seen = dict(cases=0, constrained=0, contact=0)

    @settings(deadline=None, max_examples=40, derandomize=True, suppress_health_check=list(HealthCheck))
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
        mean = np.zeros((10, 7))
        if table:
            qp[1] = rng.uniform(0.3, 0.5); qp[3] = rng.uniform(-0.4, 0.0); qp[5] = rng.uniform(-0.5, 0.0)
            mean[:, 1] = 1.0
        # the sphere may touch or slightly enter the table; a start state DEEP inside it (lift joint beyond its limit,
        # pointing down: -0.4 m was drawn once) is not a state of the system and explodes in any simulator
        assume(table_clearance(cm.tree, qp) > -0.02)
        st = dict(qp=qp, qv=rng.normal(0, vstd, 7), target_pos=rng.uniform([-.3, -.2, -.25], [.3, .2, .25]))
        K, H = 32, 10
        noise = nscale * reference_noise(K, H, 7, s % 1000)
        ref = mjstep.rollout(mjstep.OracleModel(cm.tree), st["qp"], st["qv"], st["target_pos"], mean, noise, want_traj=True,
                             nthreads=2)
        out = _rollout(emu, cm.chain.params, st, K, H, mean, noise, traj=True, ncon=True)
        assert np.isfinite(out["qv"]).all()
        scale = np.abs(ref["qv"]).max(axis=(0, 1))
        assert (np.abs(out["qv"] - ref["qv"]).max(axis=(0, 1)) / scale).max() < 1e-8
        np.testing.assert_allclose(out["costs"], ref["costs"], rtol=1e-8)
        seen["cases"] += 1
        seen["constrained"] += int((ref["ncon"] > 0).any())
        seen["contact"] += int(table)

    run()
    assert seen["constrained"] >= seen["cases"] // 2 and seen["contact"] >= 5, seen
