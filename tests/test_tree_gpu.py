"""K11: the runtime-parameterised tree rollout kernel against the CPU oracle (oracle/tree_step.c),
through the C ABI (mjb_tree_model_create / mjb_rollout_tree) and through the reference-facing env surface.
Tolerance: 1e-8 relative on costs / states over whole rollouts (FP64 on both sides, different formulations; the
measured agreement is ~1e-12), the bar SURVEY §8(c) sets for the MuJoCo dynamics."""
import os

import numpy as np
import pytest
import torch

from mjmpc_b200.envs import mjcf_tree as T
from mjmpc_b200.envs.gpu_tree_env import GpuSwimmerEnv, GpuTreeVecEnv
from oracle.tree_step import TreeOracle

pytestmark = pytest.mark.gpu
FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures")
TOL = 1e-8


def _rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / (1.0 + np.abs(np.asarray(b)).max()))


def _compare(env, oracle, state, mean, noise, nthreads=8):
    K, H, nu = noise.shape
    nv = env.nv
    env.set_env_state({"qpos": state[:nv], "qvel": state[nv:]})
    out = env.rollout_device(K, H, torch.as_tensor(mean, device=env.device), torch.as_tensor(noise, device=env.device),
                             want_states=True, want_obs=True, want_nefc=True)
    ref = oracle.rollout(state, mean, noise, env.frame_skip, env.fwd_dof, env.w_fwd, env.w_ctrl, nthreads=nthreads)
    costs, states = out["costs"].cpu().numpy(), out["states"].cpu().numpy()
    assert np.array_equal(out["actions"].cpu().numpy(), ref["actions"])          # mean + noise, recorded unclipped
    assert _rel(states, ref["states"]) < TOL
    assert _rel(costs, ref["costs"]) < TOL
    obs = out["next_observations"].cpu().numpy()
    s = env.obs_qpos_start
    assert np.array_equal(obs[..., :nv - s], states[..., s:nv]) and np.array_equal(obs[..., nv - s:], states[..., nv:])
    return int(out["nefc"].cpu().numpy().sum()), ref["nefc"], _rel(states, ref["states"])


@pytest.fixture(scope="module")
def swimmer():
    env = GpuTreeVecEnv.swimmer(contacts=False)          # the contact-free dynamics; contacts: the tests further down
    yield env, TreeOracle(env.model, T.solref_to_kb)
    env.close()


def test_swimmer_rollout_matches_oracle(swimmer):
    env, oracle = swimmer
    rng = np.random.default_rng(0)
    K, H = 192, 12
    state = np.concatenate([rng.uniform(-.1, .1, 7), rng.uniform(-.1, .1, 7)])
    mean = rng.normal(0, 0.3, (H, 4))
    noise = rng.normal(0, 0.6, (K, H, 4))
    noise[0] = 0.0
    nefc, nefc_ref, err = _compare(env, oracle, state, mean, noise)
    assert nefc == nefc_ref
    assert err < 1e-10


def test_swimmer_limits_binding(swimmer):
    """Controls pinned at the clamp drive the joints into their soft limits: most substeps carry limit rows."""
    env, oracle = swimmer
    rng = np.random.default_rng(1)
    K, H = 64, 40
    state = np.zeros(14)
    state[3:7] = [1.3, -1.4, 1.45, -1.2]
    mean = np.tile(np.array([2.0, -2.0, 2.0, -2.0]), (H, 1))          # beyond ctrlrange: clamped to +-1 inside
    noise = rng.normal(0, 0.2, (K, H, 4))
    nefc, nefc_ref, _ = _compare(env, oracle, state, mean, noise)
    assert nefc == nefc_ref and nefc > K * H            # limits really bind


def test_swimmer_ragged_and_single_particle(swimmer):
    env, oracle = swimmer
    rng = np.random.default_rng(2)
    for K, H in ((1, 1), (77, 3), (130, 2)):
        state = rng.uniform(-.3, .3, 14)
        _compare(env, oracle, state, rng.normal(0, 0.5, (H, 4)), rng.normal(0, 0.5, (K, H, 4)), nthreads=2)


@pytest.mark.parametrize("fixture", ["tree3d.xml", "tree3d_weld.xml"])
def test_branched_3d_tree_matches_oracle(fixture):
    """The run-time-topology instantiation: branches, skew axes, anchors off the body origin, rotated frames, slides
    under hinges, gravity, springs, dampers (implicit Euler), armature, per-joint solref / solimp, a fluid or a weld."""
    model = T.compile_mjcf(os.path.join(FIX, fixture))
    env = GpuTreeVecEnv(model, frame_skip=3, fwd_dof=1, w_fwd=0.7, w_ctrl=0.05, obs_qpos_start=1)
    oracle = TreeOracle(model, T.solref_to_kb)
    rng = np.random.default_rng(3)
    K, H = 96, 10
    state = np.concatenate([rng.uniform(-.4, .4, 7), rng.normal(0, 1.0, 7)])
    nefc, nefc_ref, _ = _compare(env, oracle, state, rng.normal(0, 0.5, (H, 4)), rng.normal(0, 1.0, (K, H, 4)))
    assert nefc == nefc_ref and nefc > 0
    env.close()


def _planar_switch(on):
    from mjmpc_b200 import _lib
    return _lib.lib().mjb_tree_use_planar(int(on))


def test_planar_and_general_instantiations_agree(swimmer):
    """swimmer.xml is a planar mechanism: the planar instantiation (3-vectors, registers) and the general 3-D one are
    two formulations of the same step; both are held to the oracle above, here to each other."""
    env, oracle = swimmer
    assert env.dmodel.planar
    rng = np.random.default_rng(11)
    K, H = 128, 16
    state = np.concatenate([rng.uniform(-.2, .2, 7), rng.uniform(-.5, .5, 7)])
    state[3:7] = [1.2, -1.3, 0.4, 1.45]                                     # some joints near / into their limits
    env.set_env_state({"qpos": state[:7], "qvel": state[7:]})
    mean = torch.as_tensor(rng.normal(0, 0.6, (H, 4)), device=env.device)
    noise = torch.as_tensor(rng.normal(0, 0.8, (K, H, 4)), device=env.device)
    outs = []
    for on in (1, 0):
        old = _planar_switch(on)
        try:
            o = env.rollout_device(K, H, mean, noise, want_states=True, want_nefc=True)
            outs.append({k: v.cpu().numpy().copy() for k, v in o.items()})
        finally:
            _planar_switch(old)
    assert _rel(outs[0]["states"], outs[1]["states"]) < 1e-9 and _rel(outs[0]["costs"], outs[1]["costs"]) < 1e-9
    assert np.array_equal(outs[0]["nefc"], outs[1]["nefc"]) and outs[0]["nefc"].sum() > 0
    ref = oracle.rollout(state, mean.cpu().numpy(), noise.cpu().numpy(), 4)
    for o in outs:
        assert _rel(o["states"], ref["states"]) < TOL


@pytest.mark.parametrize("n_axis", ["0 1 0", "0.6 0 0.8"])
def test_planar_branched_mechanism_matches_oracle(n_axis):
    """A half-cheetah-shaped planar tree (two legs off a torso on slide / slide / hinge, plane normal n, gravity in the
    plane, springs, dampers, armature, limits, a fluid; no contacts): 9 dofs, branched -- the <9, tree> planar
    instantiation with run-time parents, against the oracle and against the general instantiation."""
    n = np.array([float(x) for x in n_axis.split()])
    ex = np.cross(n, [0.0, 0.0, 1.0]) if abs(n[2]) < 0.9 else np.array([1.0, 0.0, 0.0])
    ex = ex / np.linalg.norm(ex)
    ez = np.cross(ex, n)
    v = lambda a, b: "%.12g %.12g %.12g" % tuple(a * ex + b * ez)          # a point / direction of the plane

    def leg(name, x, sgn):
        return ('<body name="%sthigh" pos="%s"><joint name="%sthigh" axis="%s" range="-.6 .9" stiffness="24" damping="0.6"/>'
                '<geom type="capsule" fromto="0 0 0 %s" size="0.046"/>'
                '<body name="%sshin" pos="%s"><joint name="%sshin" axis="%s" range="-.8 .8" stiffness="18" damping="0.45"/>'
                '<geom type="capsule" fromto="0 0 0 %s" size="0.04"/>'
                '<body name="%sfoot" pos="%s"><joint name="%sfoot" axis="%s" range="-.4 .7" stiffness="12" damping="0.3"/>'
                '<geom type="capsule" fromto="0 0 0 %s" size="0.035"/></body></body></body>'
                % (name, v(x, 0), name, n_axis, v(sgn * .1, -.25), name, v(sgn * .1, -.25), name, n_axis, v(-sgn * .14, -.2),
                   name, v(-sgn * .14, -.2), name, n_axis, v(sgn * .12, -.05)))
    xml = ('<mujoco model="planar9"><compiler angle="radian" inertiafromgeom="true"/>'
           '<default><joint armature="0.1" limited="true"/><geom contype="0" conaffinity="0"/>'
           '<motor ctrllimited="true" ctrlrange="-1 1"/></default>'
           '<option timestep="0.01" gravity="%s" density="30" viscosity="0.01"/><worldbody><body name="torso" pos="%s">'
           '<joint name="rx" type="slide" axis="%s" limited="false" armature="0"/>'
           '<joint name="rz" type="slide" axis="%s" limited="false" armature="0"/>'
           '<joint name="ry" type="hinge" axis="%s" limited="false" armature="0"/>'
           '<geom type="capsule" fromto="%s %s" size="0.046"/>%s%s</body></worldbody><actuator>%s</actuator></mujoco>'
           % (v(0.3, -9.81), v(0, .7), v(1, 0), v(0, 1), n_axis, v(-.5, 0), v(.5, 0), leg("b", -.5, 1.0), leg("f", .5, -1.0),
              "".join('<motor joint="%s" gear="%d"/>' % (j, g) for j, g in
                      (("bthigh", 12), ("bshin", 9), ("bfoot", 6), ("fthigh", 12), ("fshin", 6), ("ffoot", 3)))))
    model = T.compile_mjcf_string(xml)
    assert model.nv == 9 and T.pack_planar(model) is not None
    env = GpuTreeVecEnv(model, frame_skip=5, fwd_dof=0, w_fwd=1.0, w_ctrl=0.1, obs_qpos_start=1)      # half_cheetah.py:7-25
    assert env.dmodel.planar
    oracle = TreeOracle(model, T.solref_to_kb)
    rng = np.random.default_rng(5)
    K, H = 96, 8
    state = np.concatenate([rng.uniform(-.3, .3, 9), rng.normal(0, 1.0, 9)])
    mean, noise = rng.normal(0, 0.5, (H, 6)), rng.normal(0, 1.0, (K, H, 6))
    nefc, nefc_ref, _ = _compare(env, oracle, state, mean, noise)
    assert nefc == nefc_ref and nefc > 0
    old = _planar_switch(0)
    try:
        _compare(env, oracle, state, mean, noise)
    finally:
        _planar_switch(old)
    env.close()


def _walker_xml():
    return open(os.path.join(FIX, "planar_walker.xml")).read()


def test_planar_walker_on_a_floor_matches_oracle():
    """Ground contact with friction cones (the half-cheetah's kind): plane against the end spheres of capsules, several
    contacts at once, with limits, springs and implicit damping; dropped from 0.7 m, pushed by random torques."""
    model = T.compile_mjcf_string(_walker_xml(), allow_contacts="model")
    assert model.nv == 9 and len(model.contacts) == 7
    env = GpuTreeVecEnv(model, frame_skip=5, fwd_dof=0, w_fwd=1.0, w_ctrl=0.1, obs_qpos_start=1)      # half_cheetah.py:7-25
    oracle = TreeOracle(model, T.solref_to_kb)
    rng = np.random.default_rng(31)
    total = 0
    for z0 in (0.0, -0.12, -0.2):
        K, H = 64, 8
        state = np.concatenate([rng.uniform(-.1, .1, 9), rng.normal(0, 0.5, 9)])
        state[1] = z0                                     # torso height offset: feet above / at / into the floor
        nefc, nefc_ref, _ = _compare(env, oracle, state, rng.normal(0, 0.3, (H, 6)), rng.normal(0, 0.7, (K, H, 6)))
        assert nefc == nefc_ref
        total += nefc
    assert total > 64 * 8 * 5 * 8                          # the feet are on the floor most of the time
    env.close()


def test_half_cheetah_rollouts_match_oracle():
    """HalfCheetah-v0 (half_cheetah.py:7-25) on the reference's own model: the cheetah standing on its feet (the file's
    initial pose penetrates the floor by a few mm), dropped, and thrown sideways; reward = forward speed - 0.1 |a|^2."""
    env = GpuTreeVecEnv.half_cheetah()
    assert env.nv == 9 and env.d_obs == 17 and env.frame_skip == 5 and env.dmodel.n_contacts == 8
    oracle = TreeOracle(env.model, T.solref_to_kb)
    rng = np.random.default_rng(41)
    total = 0
    for case in range(3):
        K, H = 64, 8
        state = np.concatenate([rng.uniform(-.1, .1, 9), 0.1 * rng.normal(0, 1.0, 9)])          # half_cheetah.py:27-31
        if case == 1:
            state[1] += 0.3                                   # in the air first
        if case == 2:
            state[2] = 1.2; state[9] = 2.0                    # pitched forward and moving: head and torso hit the floor
        nefc, nefc_ref, _ = _compare(env, oracle, state, rng.normal(0, 0.3, (H, 6)), rng.normal(0, 0.7, (K, H, 6)))
        assert nefc == nefc_ref
        total += nefc
    assert total > 10000                                   # contact rows in most substeps
    env.close()


@pytest.mark.parametrize("name", ["cheetah", "swimmer"])
def test_contact_fuzz_one_state_per_particle(name):
    """Violent random states, one per particle (each its own controller row), two env steps: bodies deep in the floor,
    limbs folded through each other, joints beyond their limits, high speeds -- the active-set iteration (and its
    fall-back) against the oracle's Newton, row counts included."""
    env = GpuTreeVecEnv.half_cheetah() if name == "cheetah" else GpuTreeVecEnv.swimmer()
    oracle = TreeOracle(env.model, T.solref_to_kb)
    nv, nu, n, H = env.nv, env.d_action, 192, 2
    rng = np.random.default_rng(123)
    states = np.zeros((n, 2 * nv))
    for i in range(n):
        q, v = rng.uniform(-.4, .4, nv), rng.normal(0, 2.0, nv)
        if name == "cheetah":
            q[1], q[2], q[3:] = rng.uniform(-.45, 0.1), rng.uniform(-1.5, 1.5), rng.uniform(-1.2, 1.2, 6)
        else:
            q[3:] = rng.choice([-1, 1]) * rng.uniform(1.0, 1.65, 4) * rng.choice([1, 1, 1, -1], 4)
        states[i] = np.concatenate([q, v])
    env.set_env_state([dict(qpos=s[:nv], qvel=s[nv:]) for s in states])
    mean = rng.normal(0, .5, (n, H, nu))
    out = env.rollout_device(n, H, torch.as_tensor(mean, device=env.device), None, want_states=True, want_nefc=True)
    st, nefc = out["states"].cpu().numpy(), out["nefc"].cpu().numpy()
    rows = 0
    for i in range(n):
        ref = oracle.rollout(states[i], mean[i], np.zeros((1, H, nu)), env.frame_skip, env.fwd_dof, env.w_fwd, env.w_ctrl, nthreads=1)
        assert _rel(st[i], ref["states"][0]) < TOL and int(nefc[i]) == ref["nefc"]
        rows += ref["nefc"]
    assert rows > (4 if name == "cheetah" else 1.5) * n * H * env.frame_skip       # rows per substep on average
    env.close()


def test_mppi_makes_the_cheetah_run():
    from mjmpc_b200.control import MPPI
    from mjmpc_b200.envs.gpu_tree_env import GpuHalfCheetahEnv
    plant = GpuHalfCheetahEnv(seed=0)
    sim = GpuTreeVecEnv.half_cheetah()
    H, K = 16, 192
    ctrl = MPPI(d_state=18, d_obs=17, d_action=6, action_lows=sim.action_lows, action_highs=sim.action_highs, horizon=H,
                init_cov=0.3, base_action="null", num_particles=K, lam=0.05, step_size=1.0, alpha=0, gamma=1.0, n_iters=1,
                set_sim_state_fn=sim.set_env_state, rollout_fn=sim.rollout_fn, sample_mode="mean", batch_size=1, seed=0,
                filter_coeffs=[1.0, 0.0, 0.0])
    plant.reset(seed=0)
    x0 = plant.qpos[0]
    for _ in range(40):
        a, _ = ctrl.optimize(plant.get_env_state())
        obs, r, done, info = plant.step(a)
    assert obs.shape == (17,) and np.isfinite(plant.qpos).all()
    assert plant.qpos[0] - x0 > 0.3, plant.qpos[0] - x0          # 2 s of control: it has moved forward
    assert plant.qpos[1] > -0.45                                  # and has not fallen through the floor
    plant.close(); sim.close()


@pytest.mark.parametrize("n_links", [3, 8, 10])
def test_other_dof_counts_run_the_general_instantiation(n_links):
    """Swimmers of 3, 8 and 10 links (5, 10, 12 dofs; 12 = the kernel's limit): the planar instantiation of a size other
    than the reference models' (5) and the run-time-size general instantiation (10, 12; also 5 with the planar switch
    off) against the oracle."""
    radii = tuple(0.07 - 0.004 * i for i in range(n_links))
    model = T.compile_mjcf_string(T.swimmer_mjcf(radii=radii), allow_contacts="ignore")
    assert model.nv == n_links + 2 and model.nu == n_links - 1
    env = GpuTreeVecEnv(model, frame_skip=4, fwd_dof=0, w_fwd=1.0, w_ctrl=1e-4, obs_qpos_start=2)
    oracle = TreeOracle(model, T.solref_to_kb)
    rng = np.random.default_rng(n_links)
    K, H = 70, 5
    state = np.concatenate([rng.uniform(-.3, .3, model.nv), rng.normal(0, 1.0, model.nv)])
    state[3] = 1.55                                          # one joint beyond its limit from the start
    mean, noise = rng.normal(0, 0.5, (H, model.nu)), rng.normal(0, 0.8, (K, H, model.nu))
    nefc, nefc_ref, _ = _compare(env, oracle, state, mean, noise)
    assert nefc == nefc_ref and nefc > 0
    old = _planar_switch(0)
    try:
        _compare(env, oracle, state, mean, noise)
    finally:
        _planar_switch(old)
    env.close()
    if n_links == 10:
        too_big = T.compile_mjcf_string(T.swimmer_mjcf(radii=radii + (0.03,)), allow_contacts="ignore")
        with pytest.raises(T.UnsupportedMjcf, match="dofs"):
            GpuTreeVecEnv(too_big, frame_skip=4)


@pytest.mark.parametrize("seed", range(8))
def test_random_trees_match_oracle(seed):
    """Random 3-D trees (tests/helpers/random_tree.py) through the compiler and the general instantiation: rollouts vs the
    oracle; limits are hit on the way (nefc compared)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "helpers"))
    from random_tree import random_tree_xml
    model = T.compile_mjcf_string(random_tree_xml(seed))
    env = GpuTreeVecEnv(model, frame_skip=2, fwd_dof=0, w_fwd=1.0, w_ctrl=0.01, obs_qpos_start=0)
    oracle = TreeOracle(model, T.solref_to_kb)
    rng = np.random.default_rng(200 + seed)
    K, H = 48, 6
    state = np.concatenate([rng.uniform(-.6, .6, model.nv), rng.normal(0, 1.0, model.nv)])
    nefc, nefc_ref, _ = _compare(env, oracle, state, rng.normal(0, 0.5, (H, model.nu)), rng.normal(0, 1.0, (K, H, model.nu)))
    assert nefc == nefc_ref
    env.close()


@pytest.mark.parametrize("seed", range(8))
def test_random_planar_mechanisms_both_instantiations(seed):
    """Random planar mechanisms in a tilted plane whose bodies, anchors and centres of mass sit OFF the plane, with
    arbitrary inertia tensors, gravity with a component along the normal, fluid or not: the planar reduction is exact --
    planar and general instantiation against the oracle."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "helpers"))
    from random_tree import random_tree_xml
    model = T.compile_mjcf_string(random_tree_xml(seed, planar=True))
    assert model.nv <= 9 and T.pack_planar(model) is not None
    env = GpuTreeVecEnv(model, frame_skip=2, fwd_dof=0, w_fwd=1.0, w_ctrl=0.01, obs_qpos_start=0)
    assert env.dmodel.planar
    oracle = TreeOracle(model, T.solref_to_kb)
    rng = np.random.default_rng(300 + seed)
    K, H = 48, 6
    state = np.concatenate([rng.uniform(-.6, .6, model.nv), rng.normal(0, 1.0, model.nv)])
    mean, noise = rng.normal(0, 0.5, (H, model.nu)), rng.normal(0, 1.0, (K, H, model.nu))
    nefc, nefc_ref, _ = _compare(env, oracle, state, mean, noise)
    assert nefc == nefc_ref
    old = _planar_switch(0)
    try:
        _compare(env, oracle, state, mean, noise)
    finally:
        _planar_switch(old)
    env.close()


@pytest.mark.parametrize("seed", range(6))
def test_random_planar_mechanisms_with_contacts(seed):
    """Random 9-dof planar trees in the x-y, y-z and z-x planes whose capsules collide with each other (non-adjacent bodies)
    and with a floor, random friction / solref / solimp: geometry unlike the swimmer's or the cheetah's through the same
    contact code.  (In a TILTED plane MuJoCo's mju_makeFrame turns the friction pyramid of a capsule-capsule contact out of
    the plane of motion; the compiler refuses that case -- checked at the end.)"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "helpers"))
    from random_tree import random_contact_mechanism_xml
    model = T.compile_mjcf_string(random_contact_mechanism_xml(seed), allow_contacts="model")
    assert model.nv == 9 and len(model.contacts) >= 5
    env = GpuTreeVecEnv(model, frame_skip=3, fwd_dof=0, w_fwd=1.0, w_ctrl=0.01, obs_qpos_start=0)
    assert env.dmodel.n_contacts == len(model.contacts)
    oracle = TreeOracle(model, T.solref_to_kb)
    rng = np.random.default_rng(400 + seed)
    K, H = 48, 6
    state = np.concatenate([rng.uniform(-.5, .5, 9), rng.normal(0, 1.0, 9)])
    nefc, nefc_ref, _ = _compare(env, oracle, state, rng.normal(0, 0.5, (H, model.nu)), rng.normal(0, 1.0, (K, H, model.nu)))
    assert nefc == nefc_ref and nefc > 0
    env.close()
    if seed == 0:
        from random_tree import random_tree_xml
        tilted = random_tree_xml(3, planar=True).replace('contype="0" conaffinity="0"', 'contype="1" conaffinity="1"')
        m2 = T.compile_mjcf_string(tilted, allow_contacts="model")
        if any(c["kind"] == "capsule" for c in m2.contacts):
            with pytest.raises(T.UnsupportedMjcf):
                T.pack_planar_contacts(m2)


def test_non_planar_models_take_the_general_instantiation():
    model = T.compile_mjcf(os.path.join(FIX, "tree3d.xml"))
    assert T.pack_planar(model) is None


@pytest.mark.parametrize("per_worker", [4, 96])
def test_per_worker_randomised_models(per_worker):
    """randomize_dynamics (subproc_vec_env.py:304-312, gym_env_wrapper.py:367-416): every worker's particles run that
    worker's perturbed model -- against one oracle model per worker; the reference's rule and seeding on the host."""
    n_workers, H = 5, 6
    env = GpuTreeVecEnv.swimmer(n_workers=n_workers)
    spec = dict(body_mass={"torso": [0.3, 0.1], "link3": [0.2, 0.0]}, body_inertia={"link1": [0.2, -0.1]},
                dof_damping={"j2": [0.0, 0.0]}, geom_friction={"x": [0.0, 0.0]})
    d, r = env.randomize_dynamics(spec, base_seed=7)
    rng0 = np.random.RandomState(7 + 2 * 12345)
    m0 = float(env.model.body_mass[0]) * 1.1
    assert abs(r[2]["body_mass"]["torso"] - rng0.uniform(m0 - 0.3 * m0, m0 + 0.3 * m0)) < 1e-15
    assert d[0]["body_mass"]["torso"] == env.model.body_mass[0] and len({x["body_mass"]["torso"] for x in r}) == n_workers
    K = n_workers * per_worker
    rng = np.random.default_rng(8)
    state = rng.uniform(-.2, .2, 14)
    env.set_env_state({"qpos": state[:7], "qvel": state[7:]})
    mean, noise = rng.normal(0, 0.4, (H, 4)), rng.normal(0, 0.6, (K, H, 4))
    out = env.rollout_device(K, H, torch.as_tensor(mean, device=env.device), torch.as_tensor(noise, device=env.device),
                             want_states=True)
    states = out["states"].cpu().numpy()
    for w in range(n_workers):
        sl = slice(w * per_worker, (w + 1) * per_worker)
        ref = TreeOracle(env._worker_models[w], T.solref_to_kb).rollout(state, mean, noise[sl], 4)
        assert _rel(states[sl], ref["states"]) < TOL
    assert _rel(states[:per_worker], states[per_worker:2 * per_worker]) > 1e-6           # the models really differ
    with pytest.raises(AssertionError, match="divisible by number of cpus"):
        env.rollout_device(K + 1, H, torch.zeros(H, 4, dtype=torch.float64, device=env.device), None)
    with pytest.raises(ValueError):
        env.randomize_dynamics(dict(dof_frictionloss={"j1": [0.1, 0.0]}), base_seed=1)
    env.close()


def test_swimmer_self_contact_matches_oracle():
    """swimmer.xml's non-adjacent capsules collide (MuJoCo's default): rollouts started from folded poses, where the
    capsule-capsule contacts with their pyramidal friction rows act together with the joint limits."""
    env = GpuTreeVecEnv.swimmer()                            # contacts on by default
    assert env.dmodel.n_contacts == 6 and env.dmodel.planar
    oracle = TreeOracle(env.model, T.solref_to_kb)
    rng = np.random.default_rng(21)
    total = 0
    for trial in range(3):
        K, H = 64, 6
        state = np.concatenate([rng.uniform(-.2, .2, 7), rng.normal(0, 1.0, 7)])
        sgn = 1.0 if trial != 1 else -1.0
        state[3:7] = sgn * np.array([1.45, 1.5, 1.42, 0.3 if trial == 2 else 1.4])
        assert T.self_clearance(env.model, state[:7]) < 0
        nefc, nefc_ref, _ = _compare(env, oracle, state, rng.normal(0, 0.4, (H, 4)), rng.normal(0, 0.6, (K, H, 4)))
        assert nefc == nefc_ref
        total += nefc
    assert total > 3 * 64 * 4 * 4                            # contacts (4 rows each) in most substeps
    # away from contact the two models coincide
    free = GpuTreeVecEnv.swimmer(contacts=False)
    state = np.concatenate([rng.uniform(-.3, .3, 7), rng.normal(0, 0.5, 7)])
    mean, noise = rng.normal(0, 0.2, (4, 4)), rng.normal(0, 0.2, (32, 4, 4))
    outs = []
    for e in (env, free):
        e.set_env_state({"qpos": state[:7], "qvel": state[7:]})
        outs.append(e.rollout_device(32, 4, torch.as_tensor(mean, device=e.device), torch.as_tensor(noise, device=e.device),
                                     want_states=True)["states"].cpu().numpy())
    assert np.array_equal(outs[0], outs[1])
    env.close(); free.close()


def test_batched_controllers_share_one_launch(swimmer):
    env, oracle = swimmer
    rng = np.random.default_rng(4)
    n_ctrl, per, H = 3, 32, 5
    states = rng.uniform(-.2, .2, (n_ctrl, 14))
    env.set_env_state([{"qpos": s[:7], "qvel": s[7:]} for s in states])
    mean = rng.normal(0, 0.3, (n_ctrl, H, 4))
    noise = rng.normal(0, 0.5, (n_ctrl * per, H, 4))
    out = env.rollout_device(n_ctrl * per, H, torch.as_tensor(mean, device=env.device),
                             torch.as_tensor(noise, device=env.device))
    for c in range(n_ctrl):
        ref = oracle.rollout(states[c], mean[c], noise[c * per:(c + 1) * per], 4)
        assert _rel(out["costs"][c * per:(c + 1) * per].cpu().numpy(), ref["costs"]) < TOL


def test_reference_rollout_signature(swimmer):
    """GymEnvWrapper.rollout's return value (gym_env_wrapper.py:80-156): obs[:, 0] is the observation of the set state."""
    env, oracle = swimmer
    rng = np.random.default_rng(5)
    state = rng.uniform(-.1, .1, 14)
    env.set_env_state({"qpos": state[:7], "qvel": state[7:]})
    K, H = 16, 4
    mean, noise = rng.normal(0, 0.3, (H, 4)), rng.normal(0, 0.3, (K, H, 4))
    obs, rew, act, done, info, nobs = env.rollout(K, H, mean, noise)
    ref = oracle.rollout(state, mean, noise, 4)
    assert obs.shape == (K, H, 12) and rew.shape == (K, H) and act.shape == (K, H, 4) and not done.any()
    assert np.allclose(obs[:, 0], np.concatenate([state[2:7], state[7:]]))
    assert np.array_equal(obs[:, 1:], nobs[:, :-1])
    assert _rel(rew, -ref["costs"]) < TOL
    with pytest.raises(NotImplementedError):
        env.rollout(K, H, mean, noise, mode="closed_loop_linear")


def test_mppi_swims_forward():
    """MPPI on the plant (swimmer.py reward = forward velocity): the planner makes the swimmer advance along +x."""
    from mjmpc_b200.control import MPPI
    plant = GpuSwimmerEnv(seed=0, contacts=False)
    sim = GpuTreeVecEnv.swimmer(contacts=False)
    H, K = 20, 256
    ctrl = MPPI(d_state=14, d_obs=12, d_action=4, action_lows=sim.action_lows, action_highs=sim.action_highs, horizon=H,
                init_cov=0.5, base_action="null", num_particles=K, lam=0.05, step_size=1.0, alpha=0, gamma=1.0, n_iters=1,
                set_sim_state_fn=sim.set_env_state, rollout_fn=sim.rollout_fn, sample_mode="mean", batch_size=1, seed=0,
                filter_coeffs=[1.0, 0.0, 0.0])
    plant.reset(seed=0)
    x0 = plant.qpos[0]
    total, clear = 0.0, []
    for _ in range(60):
        a, _ = ctrl.optimize(plant.get_env_state())
        _, r, _, info = plant.step(a)
        total += r
        clear.append(plant.self_clearance())
    assert plant.qpos[0] - x0 > 0.15, (plant.qpos[0] - x0, total)
    assert np.all(np.abs(plant.qpos[3:]) < 1.7)
    # KNOWN LIMIT of the contact-free subset: this gait curls the swimmer up until non-adjacent capsules
    # overlap (about a third of the steps), where MuJoCo would add capsule-capsule contacts; the plant reports it
    clear = np.array(clear)
    assert np.isfinite(clear).all() and clear.max() > 0.1 and clear.min() > -0.13
    plant.close(); sim.close()


def test_cuda_graph_step_equals_eager_on_the_swimmer():
    """The captured MPC step (noise -> tree rollout -> fused update) replays bit-identically to eager launches."""
    from mjmpc_b200.control import MPPI
    rng = np.random.default_rng(7)
    states = [dict(qpos=rng.uniform(-.1, .1, 7), qvel=rng.uniform(-.1, .1, 7)) for _ in range(4)]
    acts = []
    for graph in (False, True):
        sim = GpuTreeVecEnv.swimmer()
        c = MPPI(d_state=14, d_obs=12, d_action=4, action_lows=sim.action_lows, action_highs=sim.action_highs, horizon=12,
                 init_cov=0.4, base_action="null", num_particles=512, lam=0.1, step_size=1.0, alpha=1, gamma=1.0, n_iters=1,
                 set_sim_state_fn=sim.set_env_state, rollout_fn=sim.rollout_fn, seed=3, filter_coeffs=[0.25, 0.8, 0.0])
        if graph:
            c.enable_cuda_graph(states[0])
        acts.append(np.stack([c.optimize(s)[0] for s in states]))
        sim.close()
    np.testing.assert_array_equal(acts[0], acts[1])


def test_bad_arguments_are_rejected(swimmer):
    env, _ = swimmer
    with pytest.raises(ValueError):
        env.rollout_device(8, 2, torch.zeros(2, 4, dtype=torch.float64, device=env.device),
                           torch.zeros(8, 2, 3, dtype=torch.float64, device=env.device))
    env.set_env_state([{"qpos": np.zeros(7), "qvel": np.zeros(7)}] * 3)
    with pytest.raises(ValueError):
        env.rollout_device(8, 2, torch.zeros(3, 2, 4, dtype=torch.float64, device=env.device), None)   # 8 % 3 != 0
    env.set_env_state({"qpos": np.zeros(7), "qvel": np.zeros(7)})
