"""SURVEY §8 f-3 on the CPU: the MJCF-subset compiler, the C oracle of the tree kernel (oracle/tree_step.c) against
the committed vectors of the independent restatement (oracle/tree_ref.py -> tests/golden/tree_pin.npz), closed forms
for the fluid model and energy, and the rejection of models outside the subset."""
import math
import os

import numpy as np
import pytest

from mjmpc_b200.envs import mjcf_tree as T
from oracle.tree_step import TreeOracle

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, "fixtures")
REF_XML = "/root/reference/mjmpc/envs/assets/xml"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF_XML), reason="reference tree not present")


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(HERE, "golden", "tree_pin.npz"))


def _model(name):
    if name == "swimmer":
        return T.compile_mjcf_string(T.swimmer_mjcf(), allow_contacts="ignore")
    if name == "swimmer_contact":
        return T.compile_mjcf_string(T.swimmer_mjcf(), allow_contacts="model")
    if name == "walker":
        return T.compile_mjcf(os.path.join(FIX, "planar_walker.xml"), allow_contacts="model")
    return T.compile_mjcf(os.path.join(FIX, name + ".xml"))


def _rel(a, b):
    b = np.asarray(b, float)
    return float(np.abs(np.asarray(a, float) - b).max() / (1.0 + np.abs(b).max()))


@pytest.mark.parametrize("name", ["swimmer", "tree3d", "tree3d_weld", "swimmer_contact", "walker"])
def test_c_oracle_and_compiled_constants_match_the_independent_restatement(g, name):
    m = _model(name)
    np.testing.assert_allclose(m.dof_invweight0, g[name + "_invweight0"], rtol=1e-10)
    if name != "tree3d_weld":                       # the compiler merges the welded body, tree_ref keeps it
        np.testing.assert_allclose(m.body_mass, g[name + "_mass"], rtol=1e-13)
        np.testing.assert_allclose(np.sort(m.body_inertia, axis=1), np.sort(g[name + "_inertia"], axis=1), rtol=1e-12)
    o = TreeOracle(m, T.solref_to_kb)
    n = len(g[name + "_q"])
    assert int((g[name + "_nefc"] > 0).sum()) >= 20
    if name in ("swimmer_contact", "walker"):            # four pyramid rows per contact on top of the limit rows
        assert int((g[name + "_nefc"] >= 4).sum()) >= 14 and int(g[name + "_nefc"].max()) >= 10
    worst = {}
    for i in range(n):
        r = o.substep(g[name + "_q"][i], g[name + "_v"][i], g[name + "_u"][i])
        assert r["nefc"] == int(g[name + "_nefc"][i])
        for k in ("M", "bias", "passive", "actuation", "constraint", "qacc"):
            worst[k] = max(worst.get(k, 0.0), _rel(r[k], g[name + "_" + k][i]))
        worst["q"] = max(worst.get("q", 0.0), _rel(r["q"], g[name + "_q2"][i]))
        worst["v"] = max(worst.get("v", 0.0), _rel(r["v"], g[name + "_v2"][i]))
    assert max(worst.values()) < 1e-10, worst


def test_swimmer_rollout_and_reward_match_the_independent_restatement(g):
    m = _model("swimmer")
    o = TreeOracle(m, T.solref_to_kb)
    acts = g["swimmer_roll_actions"]
    r = o.rollout(g["swimmer_roll_state0"], acts, np.zeros((1,) + acts.shape), frame_skip=4)
    assert _rel(r["states"][0], g["swimmer_roll_states"]) < 1e-10
    assert _rel(-r["costs"][0], g["swimmer_roll_rewards"]) < 1e-9          # swimmer.py:10-19


@needs_ref
def test_generated_swimmer_equals_the_reference_file():
    """The parameter list in swimmer_mjcf() IS mjmpc/envs/assets/xml/swimmer.xml as far as the physics goes."""
    a = _model("swimmer")
    b = T.compile_mjcf(os.path.join(REF_XML, "swimmer.xml"), allow_contacts="ignore")
    for f in ("body_parent", "body_pos", "body_mat", "body_mass", "body_ipos", "body_imat", "body_inertia", "jnt_type",
              "jnt_body", "jnt_pos", "jnt_axis", "jnt_limited", "jnt_damping", "jnt_armature", "jnt_stiffness", "act_dof",
              "act_gear", "act_ctrlrange", "gravity", "dof_invweight0"):
        np.testing.assert_array_equal(getattr(a, f), getattr(b, f), err_msg=f)
    np.testing.assert_array_equal(a.jnt_range[a.jnt_limited], b.jnt_range[b.jnt_limited])
    assert (a.timestep, a.density, a.viscosity) == (b.timestep, b.density, b.viscosity) == (0.005, 1000.0, 0.000894)
    with pytest.raises(T.UnsupportedMjcf, match="contacts"):
        T.compile_mjcf(os.path.join(REF_XML, "swimmer.xml"))             # non-adjacent capsules may touch: opt-in only
    assert b.ignored and "contacts dropped" in b.ignored[0]


@needs_ref
def test_generated_half_cheetah_equals_the_reference_file():
    """half_cheetah.xml: contacts must be asked for (the default refuses them), and the parameter list in
    half_cheetah_mjcf() IS the reference file as far as the physics goes -- contact candidates included."""
    with pytest.raises(T.UnsupportedMjcf, match="contacts"):
        T.compile_mjcf(os.path.join(REF_XML, "half_cheetah.xml"))
    a = T.compile_mjcf_string(T.half_cheetah_mjcf(), allow_contacts="model")
    b = T.compile_mjcf(os.path.join(REF_XML, "half_cheetah.xml"), allow_contacts="model")
    assert a.nv == 9 and a.nu == 6 and abs(a.body_mass.sum() - 14.0) < 1e-12                # settotalmass
    assert list(a.jnt_type[:3]) == [T.SLIDE, T.SLIDE, T.HINGE] and a.jnt_stiffness[3] == 240.0
    for f in ("body_parent", "body_pos", "body_mat", "body_mass", "body_ipos", "body_imat", "body_inertia", "jnt_type",
              "jnt_body", "jnt_pos", "jnt_axis", "jnt_limited", "jnt_range", "jnt_damping", "jnt_armature", "jnt_stiffness",
              "jnt_solimp", "jnt_solref", "act_dof", "act_gear", "act_ctrlrange", "gravity", "dof_invweight0"):
        np.testing.assert_array_equal(getattr(a, f), getattr(b, f), err_msg=f)
    assert len(a.contacts) == len(b.contacts) == 8                                          # floor x 8 capsules
    for ca, cb in zip(a.contacts, b.contacts):
        assert ca["kind"] == cb["kind"] == "plane" and ca["mu"] == cb["mu"] == 0.4
        for k in ("a0", "a1", "b0", "b1", "rb", "solimp", "solref", "invweight", "body2"):
            np.testing.assert_array_equal(ca[k], cb[k], err_msg=k)
    assert T.pack_planar(a) is not None and T.pack_planar_contacts(a)[0].shape == (8, 3)


@pytest.mark.parametrize("bad,why", [
    ('<joint type="ball"/>', "joint type"), ('<joint type="hinge" frictionloss="0.1"/>', "frictionloss"),
    ('<freejoint/>', "free joint"), ('<joint type="hinge" ref="0.3"/>', "joint ref")])
def test_unsupported_elements_raise(bad, why):
    xml = ('<mujoco><compiler angle="radian"/><worldbody><body>%s<geom type="sphere" size="0.1" contype="0" conaffinity="0"/>'
           '</body></worldbody></mujoco>' % bad)
    with pytest.raises(T.UnsupportedMjcf, match=why):
        T.compile_mjcf_string(xml)
    for opt, why2 in (('integrator="RK4"', "integrator"), ('wind="1 0 0"', "wind")):
        with pytest.raises(T.UnsupportedMjcf, match=why2):
            T.compile_mjcf_string('<mujoco><option %s/><worldbody><body><joint/><geom size="0.1" contype="0" conaffinity="0"/>'
                                  '</body></worldbody></mujoco>' % opt)


def test_compiler_geometry_rules_closed_forms():
    """Mass properties of every geom type, degree angles, nested default classes, class / childclass, <inertial>."""
    xml = ('<mujoco><compiler inertiafromgeom="auto"/>'                                   # angle: degrees by default
           '<default><geom contype="0" conaffinity="0" density="500"/><joint damping="0.2"/>'
           '<default class="a"><joint range="-90 90" limited="true"/><default class="b"><joint armature="0.3"/></default></default>'
           '</default><worldbody>'
           '<body name="box" childclass="a"><joint name="j0" axis="0 0 1"/><geom type="box" size="0.1 0.2 0.3"/>'
           '<body name="cyl" pos="1 0 0" euler="0 0 90"><joint name="j1" class="b" type="slide" axis="1 0 0" range="-1 2"/>'
           '<geom type="cylinder" size="0.1 0.25" axisangle="1 0 0 90"/>'
           '<body name="given" pos="0 1 0"><joint name="j2" class="main" axis="0 1 0"/>'
           '<inertial pos="0.1 0 0" mass="2.5" diaginertia="0.3 0.2 0.1"/><geom type="sphere" size="0.5"/></body></body></body>'
           '</worldbody></mujoco>')
    m = T.compile_mjcf_string(xml)
    assert m.nv == 3 and list(m.jnt_type) == [T.HINGE, T.SLIDE, T.HINGE]
    # box 0.2 x 0.4 x 0.6 at density 500
    mb = 500 * 0.2 * 0.4 * 0.6
    np.testing.assert_allclose(m.body_mass[0], mb, rtol=1e-14)
    np.testing.assert_allclose(m.body_inertia[0], mb / 12 * np.array([0.4 ** 2 + 0.6 ** 2, 0.2 ** 2 + 0.6 ** 2, 0.2 ** 2 + 0.4 ** 2]), rtol=1e-13)
    # cylinder r 0.1, length 0.5, its axis turned from z to -y by the 90 degree axisangle
    mc = 500 * math.pi * 0.01 * 0.5
    np.testing.assert_allclose(m.body_mass[1], mc, rtol=1e-14)
    np.testing.assert_allclose(m.body_inertia[1], [mc * (3 * 0.01 + 0.25) / 12] * 2 + [0.5 * mc * 0.01], rtol=1e-13)
    np.testing.assert_allclose(np.abs(m.body_imat[1][:, 2]), [0, 1, 0], atol=1e-15)
    np.testing.assert_allclose(m.body_mat[1], [[0, -1, 0], [1, 0, 0], [0, 0, 1]], atol=1e-15)        # euler 0 0 90 degrees
    # <inertial> wins over the geom under inertiafromgeom="auto"
    np.testing.assert_allclose([m.body_mass[2], *m.body_ipos[2], *m.body_inertia[2]], [2.5, 0.1, 0, 0, 0.3, 0.2, 0.1])
    # defaults: class a (range in degrees for the hinge, limited), class b inherits a and adds armature; slide ranges are lengths
    np.testing.assert_allclose(m.jnt_range[0], [-math.pi / 2, math.pi / 2], rtol=1e-15)
    np.testing.assert_allclose(m.jnt_range[1], [-1, 2])
    assert list(m.jnt_limited) == [True, True, False] and list(m.jnt_armature) == [0.0, 0.3, 0.0]
    assert list(m.jnt_damping) == [0.2, 0.2, 0.2]


def test_solimp_is_clamped_like_mujoco_and_both_restatements_agree(tmp_path):
    """half_cheetah.xml's joints ask for solimplimit="0 .8 .03": MuJoCo clamps d0 to 1e-4 (a zero impedance would make
    the regulariser infinite).  The C oracle and the independent restatement agree on such a model."""
    from oracle import tree_ref
    xml = ('<mujoco><compiler angle="radian" inertiafromgeom="true"/><option timestep="0.01" gravity="0 0 -9.81"/>'
           '<default><joint limited="true" solimplimit="0 .8 .03" solreflimit=".02 1" armature="0.1" damping="0.01" stiffness="8"/>'
           '<geom contype="0" conaffinity="0"/></default><worldbody><body pos="0 0 1">'
           '<joint name="a" axis="0 1 0" range="-.5 .7"/><geom type="capsule" fromto="0 0 0 0.3 0 -0.2" size="0.05"/>'
           '<body pos="0.3 0 -0.2"><joint name="b" axis="0 1 0" range="-.4 .4"/><geom type="capsule" fromto="0 0 0 -0.1 0 -0.3" size="0.04"/>'
           '</body></body></worldbody><actuator><motor joint="a" gear="30"/><motor joint="b" gear="10"/></actuator></mujoco>')
    f = tmp_path / "m.xml"
    f.write_text(xml)
    m = T.compile_mjcf(str(f))
    np.testing.assert_allclose(m.jnt_solimp[0], [1e-4, 0.8, 0.03, 0.5, 2.0])
    o, rm = TreeOracle(m, T.solref_to_kb), tree_ref.read_model(str(f))
    rng = np.random.default_rng(0)
    hit = 0
    for _ in range(20):
        q, v, u = rng.uniform(-.9, .9, 2), rng.normal(0, 2, 2), rng.normal(0, 1, 2)
        a = o.substep(q, v, u)
        q2, v2, info = tree_ref.step(rm, q, v, u)
        assert a["nefc"] == info["nefc"]
        hit += a["nefc"] > 0
        assert _rel(a["q"], q2) < 1e-11 and _rel(a["v"], v2) < 1e-10 and _rel(a["constraint"], info["constraint"]) < 1e-10
    assert hit >= 8


@pytest.mark.parametrize("seed", range(8))
def test_random_trees_both_restatements_agree(seed):
    """Random 3-D trees (tests/helpers/random_tree.py: topology, axes, anchors, orientations, springs, dampers, limits with
    random solref / solimp, motors, fluid or not): the compiler + C oracle against the independent restatement."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "helpers"))
    from random_tree import random_tree_xml
    from oracle import tree_ref
    xml = random_tree_xml(seed)
    m = T.compile_mjcf_string(xml)
    o, rm = TreeOracle(m, T.solref_to_kb), tree_ref.read_model(None, text=xml)
    np.testing.assert_allclose(m.dof_invweight0, rm["invweight0"], rtol=1e-9)
    rng = np.random.default_rng(100 + seed)
    rows = 0
    for _ in range(4):
        q, v, u = rng.uniform(-1, 1, m.nv), rng.normal(0, 2, m.nv), rng.normal(0, 1, m.nu)
        a = o.substep(q, v, u)
        q2, v2, info = tree_ref.step(rm, q, v, u)
        assert a["nefc"] == info["nefc"]
        rows += a["nefc"]
        for k in ("M", "bias", "passive", "actuation", "constraint", "qacc"):
            assert _rel(a[k], info[k]) < 1e-10, k
        assert _rel(a["q"], q2) < 1e-11 and _rel(a["v"], v2) < 1e-10


def test_capsule_contacts_in_a_tilted_plane_are_refused():
    """mju_makeFrame takes the first tangent of a capsule-capsule contact from the world y / z axis: in a tilted plane of
    motion the friction pyramid is turned out of the plane, which the planar kernel's three merged rows do not represent."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "helpers"))
    from random_tree import random_contact_mechanism_xml, random_tree_xml
    ok = T.compile_mjcf_string(random_contact_mechanism_xml(1), allow_contacts="model")       # a coordinate plane
    assert T.pack_planar_contacts(ok)[0].shape[0] == len(ok.contacts) >= 5
    found = False
    for seed in range(6):
        tilted = random_tree_xml(seed, planar=True).replace('contype="0" conaffinity="0"', 'contype="1" conaffinity="1"')
        m = T.compile_mjcf_string(tilted, allow_contacts="model")
        if any(c["kind"] == "capsule" for c in m.contacts):
            found = True
            with pytest.raises(T.UnsupportedMjcf, match="coordinate plane|heights"):
                T.pack_planar_contacts(m)
    assert found


def test_welded_body_in_a_fluid_is_rejected():
    xml = open(os.path.join(FIX, "tree3d_weld.xml")).read().replace('density="0" viscosity="0"', 'density="10" viscosity="0"')
    with pytest.raises(T.UnsupportedMjcf, match="welded"):
        T.compile_mjcf_string(xml)


def _slider(rho, mu, r=0.05, half=0.2):
    return ('<mujoco><compiler angle="radian" inertiafromgeom="true"/><option timestep="0.001" gravity="0 0 0" density="%g" '
            'viscosity="%g"/><worldbody><body><joint type="slide" axis="1 0 0"/><joint type="hinge" axis="0 0 1"/>'
            '<geom type="capsule" size="%g %g" contype="0" conaffinity="0"/></body></worldbody></mujoco>' % (rho, mu, r, half))


def test_fluid_forces_closed_form():
    """One capsule (axis = z) sliding along x and spinning about z: mj_passive's inertia-box model by hand."""
    rho, mu, r, half = 700.0, 0.3, 0.05, 0.2
    m = T.compile_mjcf_string(_slider(rho, mu, r, half))
    o = TreeOracle(m, T.solref_to_kb)
    hh = 2 * half
    mc, ms = 1000 * math.pi * r * r * hh, 1000 * 4 / 3 * math.pi * r ** 3
    mass = mc + ms
    It = mc * (3 * r * r + hh * hh) / 12 + ms * (0.4 * r * r + 0.25 * hh * hh + 0.375 * r * hh)
    Ia = 0.5 * mc * r * r + 0.4 * ms * r * r
    bx = by = math.sqrt((It + Ia - It) / mass * 6)
    bz = math.sqrt((2 * It - Ia) / mass * 6)
    d = (bx + by + bz) / 3
    vx, wz = 1.7, -2.3
    fx = -3 * math.pi * d * mu * vx - 0.5 * rho * by * bz * abs(vx) * vx
    tz = -math.pi * d ** 3 * mu * wz - rho * bz * (bx ** 4 + by ** 4) * abs(wz) * wz / 64
    res = o.substep(np.zeros(2), np.array([vx, wz]), np.zeros(0))
    np.testing.assert_allclose(res["passive"], [fx, tz], rtol=1e-12)
    np.testing.assert_allclose(res["qacc"], [fx / mass, tz / Ia], rtol=1e-12)


def test_pendulum_energy_is_conserved_to_first_order():
    xml = ('<mujoco><compiler angle="radian" inertiafromgeom="true"/><option timestep="0.0005"/><worldbody><body pos="0 0 1">'
           '<joint type="hinge" axis="0 1 0"/><geom type="capsule" fromto="0 0 0 0.4 0 0" size="0.03" contype="0" '
           'conaffinity="0"/><body pos="0.4 0 0"><joint type="hinge" axis="0 1 0"/><geom type="capsule" fromto="0 0 0 0.3 0 0" '
           'size="0.02" contype="0" conaffinity="0"/></body></body></worldbody></mujoco>')
    m = T.compile_mjcf_string(xml)
    o = TreeOracle(m, T.solref_to_kb)

    def energy(q, v):
        xpos, xmat, _, _ = T.kinematics(m, q)
        pe = sum(-m.body_mass[b] * (m.gravity @ (xpos[b] + xmat[b] @ m.body_ipos[b])) for b in range(m.nb))
        return 0.5 * v @ T.mass_matrix(m, q) @ v + pe
    q, v = np.array([0.3, -0.5]), np.array([0.0, 0.0])
    e0 = energy(q, v)
    for _ in range(2000):                                  # 1 s of a double pendulum falling under gravity
        r = o.substep(q, v, np.zeros(0))
        q, v = r["q"], r["v"]
    assert np.abs(v).max() > 1.0                           # it really moved
    assert abs(energy(q, v) - e0) < 2e-3 * abs(e0 - energy(np.array([math.pi / 2, 0.0]), np.zeros(2)))


def test_linear_momentum_drift_is_first_order_in_the_timestep():
    """No fluid, planar base on two slides: x / y momentum (the slides' generalised momenta) is conserved by the
    continuous dynamics whatever the joint torques and limit forces do; semi-implicit Euler drifts O(h)."""
    drift = []
    for h, n in ((0.005, 100), (0.0005, 1000)):
        m = T.compile_mjcf_string(T.swimmer_mjcf(density=0.0, viscosity=0.0, timestep=h), allow_contacts="ignore")
        o = TreeOracle(m, T.solref_to_kb)
        rng = np.random.default_rng(0)
        q, v = rng.uniform(-.3, .3, 7), rng.normal(0, 1, 7)
        u = np.array([1, -1, .5, -.7])
        p0 = (T.mass_matrix(m, q) @ v)[:2].copy()
        for t in range(n):
            r = o.substep(q, v, u * np.sin(40 * h * t))
            q, v = r["q"], r["v"]
        drift.append(np.abs((T.mass_matrix(m, q) @ v)[:2] - p0).max())
    assert drift[0] > 1.0 and 7.0 < drift[0] / drift[1] < 13.0, drift


def test_self_clearance_marks_the_contact_free_subset():
    """swimmer.xml's non-adjacent capsules only touch when three consecutive joints fold beyond ~1.4 rad to the same
    side; the compiler drops those contacts and self_clearance() says when a pose would have needed them."""
    m = _model("swimmer")
    assert len(m.shapes) == 5
    q = np.zeros(7)
    assert abs(T.self_clearance(m, q) - (0.3 - 0.07 - 0.06)) < 1e-12         # torso and link 2, end to end
    q[3:6] = 1.5
    assert T.self_clearance(m, q) < 0.0
    q[3:6] = 1.3
    assert T.self_clearance(m, q) > 0.1
    q[3:7] = [1.5, 1.5, -1.5, -1.5]
    assert T.self_clearance(m, q) > 0.1
    assert T.self_clearance(_model("tree3d"), np.zeros(7)) == np.inf          # contype = conaffinity = 0


def test_kernel_layout_matches_the_python_packer():
    """csrc/tree_model.h and mjcf_tree.py agree (host-only call into the library, no GPU work)."""
    import ctypes as C
    from mjmpc_b200 import _lib
    try:
        L = _lib.lib()
    except _lib.MjbError:
        pytest.skip("extension not built")
    out = (C.c_int * 60)()
    L.mjb_tree_layout(out)
    assert list(out)[:4] == [T.LK_RFIX, T.LK_OFF, T.LK_AXIS, T.LK_MASS]
    assert out[21] == T.LK_STRIDE and out[27] == T.LI_STRIDE and out[32] == T.G_STRIDE and out[33] == T.MAX_LINKS
    assert list(out)[34:45] == [T.PK_OFF, T.PK_DIR, T.PK_MASS, T.PK_COM, T.PK_INN, T.PK_CLIN, T.PK_KV1, T.PK_KV2, T.PK_E, T.PK_AK, T.PK_STRIDE]
    assert list(out)[45:] == [T.CT_A, T.CT_HA, T.CT_RA, T.CT_B, T.CT_HB, T.CT_RB, T.CT_MU, T.CT_K, T.CT_BB, T.CT_SOLIMP, T.CT_INVW,
                              T.CT_BOUND, T.CT_STRIDE, T.CTI_STRIDE, T.MAX_CAND]


def test_pack_links_shapes_and_topology():
    m = _model("tree3d")
    P, I, G = T.pack_links(m)
    assert P.shape == (7, T.LK_STRIDE) and I.shape == (7, T.LI_STRIDE) and G.shape == (T.G_STRIDE,)
    assert list(I[:, T.LI_PARENT]) == [-1, 0, 1, 2, 3, 1, 5]            # two branches off the base's last link
    assert list(I[:, T.LI_BODY] & 1) == [0, 1, 0, 1, 1, 1, 1]               # massless links of the multi-joint bodies
    sw = T.pack_links(_model("swimmer"))[1]
    assert list(sw[:, T.LI_PARENT]) == [-1, 0, 1, 2, 3, 4, 5] and list(sw[:, T.LI_ACT]) == [-1, -1, -1, 0, 1, 2, 3]
