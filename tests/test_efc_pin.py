"""Pins the soft-constraint half of the dynamics (joint limits, sphere-plane contact) against tests/golden/efc_pin.npz:
outputs of oracle/efc_ref.py, the second restatement of MuJoCo 2.0's mj_step for the reference's sawyer.xml that
shares no code and no pre-computed constant with oracle/mjstep.c, mjmpc_b200/envs/model.py or the CUDA kernel (own
MJCF reader, Jacobian-sum mass matrix, complex-step Coriolis terms, its own dof_invweight0 / body_invweight0 / K / B,
active-set ENUMERATION instead of Newton).  A formula misread the same way in model.py's constants and in
mjstep.c's rows fails here.

CPU tests: the C oracle and the compiled model constants vs the vectors, efc_ref itself vs the vectors where the
reference tree is present, and 1-DOF closed forms.  GPU test: the CUDA rollout vs efc_ref's rollouts (1e-8)."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
XML = "/root/reference/mjmpc/envs/assets/xml/sawyer.xml"


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(HERE, "golden", "efc_pin.npz"))


def _rel(a, b):
    b = np.asarray(b, float)
    return float(np.abs(np.asarray(a, float) - b).max() / max(1e-300, np.abs(b).max())) if b.size else 0.0


def test_compiled_model_constants_match_the_independent_compile(g, compiled_model):
    t = compiled_model.tree
    np.testing.assert_allclose(t.mass, g["model_mass"], rtol=1e-13)
    np.testing.assert_allclose(t.ipos, g["model_com"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(t.inertia.reshape(-1, 3, 3), g["model_inertia"], rtol=1e-12, atol=1e-16)
    np.testing.assert_allclose(t.dof_invweight0, g["model_dof_invweight0"], rtol=1e-11)
    np.testing.assert_allclose(t.con_invweight, float(g["model_con_invweight"]), rtol=1e-11)
    from mjmpc_b200.envs.model import solref_to_kb
    K, B = solref_to_kb(t.solref, t.solimp, t.timestep)
    np.testing.assert_allclose([K, B], [float(g["model_solK"]), float(g["model_solB"])], rtol=1e-14)
    np.testing.assert_array_equal(t.jnt_range, g["model_range"])
    np.testing.assert_array_equal(t.gear, g["model_gear"])


def test_c_oracle_rows_and_step_match_the_independent_restatement(g, oracle_model):
    assert int((g["nefc"] > 0).sum()) >= 40 and int(g["has_contact"].sum()) >= 20 and int(g["nefc"].max()) >= 4
    worst = dict(aref=0.0, D=0.0, J=0.0, force=0.0, qfrc=0.0, q=0.0, v=0.0)
    for i in range(len(g["q"])):
        o = oracle_model.substep_efc(g["q"][i], g["v"][i], g["u"][i])
        n = int(g["nefc"][i])
        assert len(o["aref"]) == n, "state %d: %d rows vs %d" % (i, len(o["aref"]), n)
        worst["q"] = max(worst["q"], _rel(o["q"], g["q2"][i]))
        worst["v"] = max(worst["v"], _rel(o["v"], g["v2"][i]))
        if n:
            worst["J"] = max(worst["J"], float(np.abs(o["J"] - g["J"][i, :n]).max()))
            worst["aref"] = max(worst["aref"], _rel(o["aref"], g["aref"][i, :n]))
            worst["D"] = max(worst["D"], _rel(o["D"], g["D"][i, :n]))
            if np.abs(g["force"][i, :n]).max() > 0:
                worst["force"] = max(worst["force"], _rel(o["force"], g["force"][i, :n]))
                worst["qfrc"] = max(worst["qfrc"], _rel(o["qfrc_constraint"], g["qfrc_constraint"][i]))
    assert worst["J"] < 1e-12 and worst["aref"] < 1e-11 and worst["D"] < 1e-12, worst
    assert worst["force"] < 1e-9 and worst["qfrc"] < 1e-9, worst          # D ~ 1e3..1e5 amplifies rounding of J a - aref
    assert worst["q"] < 1e-12 and worst["v"] < 1e-11, worst


def test_c_oracle_rollouts_match_the_independent_restatement(g, oracle_model):
    from oracle import mjstep
    for s in range(3):
        ref = mjstep.rollout(oracle_model, g["ro_q0"][s], g["ro_v0"][s], g["ro_target"], g["ro_mean"][s],
                             np.ascontiguousarray(g["ro_noise"][s]), want_traj=True)
        scale = np.abs(g["ro_qv"][s]).max(axis=(0, 1))
        assert (np.abs(ref["qv"] - g["ro_qv"][s]).max(axis=(0, 1)) / scale).max() < 1e-10
        np.testing.assert_allclose(ref["costs"], g["ro_costs"][s], rtol=1e-11)


@pytest.mark.skipif(not os.path.exists(XML), reason="reference tree absent (GPU box)")
def test_vectors_are_what_efc_ref_computes_from_the_reference_xml(g):
    from oracle import efc_ref
    m = efc_ref.read_model(XML)
    np.testing.assert_allclose(m["dof_invweight0"], g["model_dof_invweight0"], rtol=1e-13)
    for i in range(0, len(g["q"]), 7):
        q2, v2, info = efc_ref.step(m, g["q"][i], g["v"][i], g["u"][i])
        np.testing.assert_allclose(q2, g["q2"][i], rtol=1e-13)
        np.testing.assert_allclose(v2, g["v2"][i], rtol=1e-12)
        assert info["nefc"] == int(g["nefc"][i])


def test_one_dof_limit_closed_form():
    """A single hinge beyond its upper limit: MuJoCo's formulas by hand -- impedance (power-2 ramp), R, aref, the
    one-row minimiser a = (f + D aref J) / (M + D), the force and the implicit-damping Euler step -- against the C
    oracle on a one-body model."""
    import ctypes as C
    from oracle import mjstep
    L = mjstep.lib()
    I, arm, damp, gear, h = 0.03, 0.004, 0.8, 10.0, 0.01
    lo, hi = -1.0, 0.5
    solimp = np.array([0.9, 0.95, 0.001, 0.5, 2.0])
    tc, dr = 0.02, 1.0
    Kk, B = 1.0 / (0.95 ** 2 * tc ** 2 * dr ** 2), 2.0 / (0.95 * tc)
    M = I + arm
    f64 = lambda a: np.ascontiguousarray(a, np.float64)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    P = mjstep._p
    keep = [i32([-1]), f64([0, 0, 0]), f64([1.0]), f64([0, 0, 0]), f64(np.diag([I, I, I])), i32([0]), f64([0, 0, 1]),
            f64([lo, hi]), i32([1]), f64([arm]), f64([damp]), f64([gear]), f64([-1, 1]), f64([1.0 / M]), f64(solimp),
            f64([0, 0, 0]), f64([0, 0, 0])]
    hnd = L.ora_model_create(C.c_int(1), C.c_int(1), P(keep[0], C.c_int), P(keep[1]), P(keep[2]), P(keep[3]), P(keep[4]),
                             P(keep[5], C.c_int), P(keep[6]), P(keep[7]), P(keep[8], C.c_int), P(keep[9]), P(keep[10]),
                             P(keep[11]), P(keep[12]), P(keep[13]), C.c_double(h), C.c_int(1), C.c_double(Kk), C.c_double(B),
                             P(keep[14]), C.c_int(0), P(keep[15]), C.c_int(0), P(keep[16]), C.c_double(-1.0), C.c_double(0.0),
                             C.c_double(0.0), C.c_double(0.0))
    assert hnd
    for pen, vel, u in [(0.0004, 0.7, 0.3), (0.0009, -0.2, -1.0), (0.02, 2.0, 1.0), (0.0002, -3.0, 0.0)]:
        q, v = np.array([hi + pen]), np.array([vel])
        dist = hi - q[0]                                      # < 0, row Jacobian -1
        x = min(abs(dist) / solimp[2], 1.0)
        y = x * x / 0.5 if x <= 0.5 else 1.0 - (1.0 - x) ** 2 / 0.5
        imp = 0.9 + y * 0.05 if x < 1.0 else 0.95
        D = 1.0 / max(1e-15, (1.0 - imp) * (1.0 / M) / imp)
        aref = -B * (-vel) - Kk * imp * dist
        f = gear * np.clip(u, -1, 1) - damp * vel             # single hinge about a principal axis: no bias force
        a0 = f / M
        active = (-a0 - aref) < 0.0
        a = (f - D * aref) / (M + D) if active else a0
        force = -D * (-a - aref) if active else 0.0
        qacc = (f - force) / (M + h * damp)                   # J' force = -force
        v2 = vel + h * qacc
        q2 = q[0] + h * v2
        qq, vv, qa = q.copy(), v.copy(), np.zeros(1)
        n = L.ora_substep(C.c_void_p(hnd), P(qq), P(vv), P(f64([u])), P(qa))
        assert n == 1
        np.testing.assert_allclose([qq[0], vv[0], qa[0]], [q2, v2, qacc], rtol=1e-12)
    L.ora_model_destroy(C.c_void_p(hnd))


@pytest.mark.gpu
def test_cuda_rollout_matches_the_independent_restatement(g, compiled_model, split_switch):
    """The CUDA kernels (thread-per-particle and role-split) against efc_ref's rollouts: interior start, the env's
    reset state (limits bind at once), the sphere 1 cm above the table moving down (contact row)."""
    import torch
    from mjmpc_b200.envs.gpu_vec_env import GpuReacherVecEnv
    env = GpuReacherVecEnv(compiled_model)
    for thr in (0, 1 << 20):
        split_switch(thr)
        for s in range(3):
            K, H = g["ro_noise"][s].shape[:2]
            env.set_env_state(dict(qp=g["ro_q0"][s], qv=g["ro_v0"][s], target_pos=g["ro_target"]))
            out = env.rollout_device(K, H, torch.from_numpy(g["ro_mean"][s]).cuda(),
                                     torch.from_numpy(np.ascontiguousarray(g["ro_noise"][s])).cuda(), want_traj=True)
            scale = np.abs(g["ro_qv"][s]).max(axis=(0, 1))
            err = (np.abs(out["qv"].cpu().numpy() - g["ro_qv"][s]).max(axis=(0, 1)) / scale).max()
            assert err < 1e-8, (thr, s, err)
            np.testing.assert_allclose(out["costs"].cpu().numpy(), g["ro_costs"][s], rtol=1e-9)
    env.close()
