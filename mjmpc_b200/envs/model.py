"""Host-side model compiler for hinge-joint articulated models (the reacher_7dof arm).

The reference gets its dynamics model from MuJoCo compiling
``mjmpc/envs/assets/xml/sawyer.xml`` (reference: sawyer.xml:1-110; env ctor
``mjmpc/envs/basic/reacher_env.py:21``).  MuJoCo is not part of the reference
tree, so this module restates the few compiler rules that model needs:

* ``inertiafromgeom``: every geom contributes mass at density 1000 kg/m^3
  (sphere and capsule volume / inertia formulas, parallel-axis composition
  about the body COM);
* joint defaults (armature, damping, limited), motor gear / ctrlrange;
* model constants MuJoCo derives at qpos0: ``dof_invweight0`` and
  ``body_invweight0`` (regularisers of the soft-constraint solver).

Two products come out of :func:`compile_model`:

``TreeModel``   un-merged body tree, generic axes -- consumed by the CPU oracle
                (``oracle/mjstep.c``), which follows MuJoCo's own formulation.
``ChainModel``  welded bodies folded into their parents, axis-aligned serial
                chain -- consumed by the CUDA rollout kernel
                (``csrc/rollout_reacher.cu``) as one flat double array.

Nothing here runs on the hot path: it executes once per model (and once per
randomised instance, see ``randomize`` in ``gpu_vec_env.py``).
"""
from __future__ import annotations

import copy
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

AXES = {"x": (1.0, 0.0, 0.0), "y": (0.0, 1.0, 0.0), "z": (0.0, 0.0, 1.0)}


# --------------------------------------------------------------------------
# Specification (what an MJCF file says)
# --------------------------------------------------------------------------
@dataclass
class GeomSpec:
    kind: str                      # "sphere" | "capsule" | "plane"
    size: float                    # radius
    pos: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    fromto: Optional[Tuple[float, ...]] = None   # capsule end points
    contact: bool = False          # contype/conaffinity == 1
    name: str = ""
    density: float = 1000.0


@dataclass
class JointSpec:
    name: str
    axis: Tuple[float, float, float]
    range: Tuple[float, float]
    damping: float
    armature: float
    limited: bool = True
    frictionloss: float = 0.0


@dataclass
class BodySpec:
    name: str
    pos: Tuple[float, float, float]
    parent: int                    # index into ModelSpec.bodies, -1 = world
    joint: Optional[JointSpec] = None
    geoms: List[GeomSpec] = field(default_factory=list)
    sites: Dict[str, Tuple[float, float, float]] = field(default_factory=dict)


@dataclass
class ActuatorSpec:
    joint: str
    gear: float
    ctrlrange: Tuple[float, float]
    ctrllimited: bool = True


@dataclass
class ModelSpec:
    bodies: List[BodySpec]
    actuators: List[ActuatorSpec]
    timestep: float = 0.01
    gravity: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    frame_skip: int = 2
    geom_margin: float = 0.002
    world_geoms: List[GeomSpec] = field(default_factory=list)
    world_sites: Dict[str, Tuple[float, float, float]] = field(default_factory=dict)
    # MuJoCo 2.0 built-in solver defaults (not present in the xml)
    solref: Tuple[float, float] = (0.02, 1.0)
    solimp: Tuple[float, float, float, float, float] = (0.9, 0.95, 0.001, 0.5, 2.0)


def reacher7dof_spec() -> ModelSpec:
    """The 7-DOF Sawyer-like reacher of the reference (sawyer.xml:15-59 body
    tree, :101-109 actuators, :3 options, :5-6 defaults)."""
    arm, dmp = 0.004, 0.8

    def hinge(name, ax, lo, hi, damping=dmp):
        return JointSpec(name, AXES[ax], (lo, hi), damping, arm)

    def cap(name, a, b, r):
        return GeomSpec("capsule", r, fromto=tuple(a) + tuple(b), name=name)

    def sph(name, p, r, contact=False):
        return GeomSpec("sphere", r, pos=tuple(p), name=name, contact=contact)

    B = []
    B.append(BodySpec("r_shoulder_pan_link", (0.0, -0.6, 0.0), -1,
                      hinge("r_shoulder_pan_joint", "z", -2.2854, 1.714602, 2.0),
                      [sph("e1", (-0.06, 0.05, 0.2), 0.05), sph("e2", (0.06, 0.05, 0.2), 0.05),
                       sph("e1p", (-0.06, 0.09, 0.2), 0.03), sph("e2p", (0.06, 0.09, 0.2), 0.03),
                       cap("sp", (0, 0, -0.4), (0, 0, 0.2), 0.1)]))
    B.append(BodySpec("r_shoulder_lift_link", (0.1, 0.0, 0.0), 0,
                      hinge("r_shoulder_lift_joint", "y", -0.5236, 1.3963, 2.0),
                      [cap("sl", (0, -0.1, 0), (0, 0.1, 0), 0.1)]))
    B.append(BodySpec("r_upper_arm_roll_link", (0.0, 0.0, 0.0), 1,
                      hinge("r_upper_arm_roll_joint", "x", -1.5, 1.7),
                      [cap("uar", (-0.1, 0, 0), (0.1, 0, 0), 0.02)]))
    B.append(BodySpec("r_upper_arm_link", (0.0, 0.0, 0.0), 2, None,
                      [cap("ua", (0, 0, 0), (0.4, 0, 0), 0.06)]))
    B.append(BodySpec("r_elbow_flex_link", (0.4, 0.0, 0.0), 3,
                      hinge("r_elbow_flex_joint", "y", -2.3213, 0.0),
                      [cap("ef", (0, -0.02, 0), (0, 0.02, 0), 0.06)]))
    B.append(BodySpec("r_forearm_roll_link", (0.0, 0.0, 0.0), 4,
                      hinge("r_forearm_roll_joint", "x", -1.5, 1.5),
                      [cap("fr", (-0.1, 0, 0), (0.1, 0, 0), 0.02)]))
    B.append(BodySpec("r_forearm_link", (0.0, 0.0, 0.0), 5, None,
                      [cap("fa", (0, 0, 0), (0.291, 0, 0), 0.05)]))
    B.append(BodySpec("r_wrist_flex_link", (0.321, 0.0, 0.0), 6,
                      hinge("r_wrist_flex_joint", "y", -1.094, 0.0),
                      [cap("wf", (0, -0.02, 0), (0, 0.02, 0), 0.01)]))
    B.append(BodySpec("r_wrist_roll_link", (0.0, 0.0, 0.0), 7,
                      hinge("r_wrist_roll_joint", "x", -1.5, 1.5),
                      [sph("ee", (0.03, 0, 0), 0.08, contact=True)],
                      sites={"finger": (0.0, 0.0, 0.0)}))
    gears = [20.0] + [10.0] * 6
    acts = [ActuatorSpec(b.joint.name, g, (-1.0, 1.0))
            for b, g in zip([b for b in B if b.joint], gears)]
    table = GeomSpec("plane", 0.0, pos=(0.0, 0.5, -0.425), contact=True, name="table")
    return ModelSpec(B, acts, timestep=0.01, gravity=(0.0, 0.0, 0.0), frame_skip=2,
                     geom_margin=0.002, world_geoms=[table],
                     world_sites={"target": (0.1, 0.1, 0.1)})


# --------------------------------------------------------------------------
# geom -> mass / COM / inertia  (MuJoCo compiler rules, density * volume)
# --------------------------------------------------------------------------
def geom_inertial(g: GeomSpec):
    """Return (mass, centre(3), inertia(3x3) about the centre, body axes)."""
    r = g.size
    if g.kind == "sphere":
        m = g.density * 4.0 / 3.0 * math.pi * r ** 3
        return m, np.array(g.pos, float), np.eye(3) * (0.4 * m * r * r)
    if g.kind == "capsule":
        a = np.array(g.fromto[:3], float)
        b = np.array(g.fromto[3:], float)
        h = float(np.linalg.norm(b - a))            # cylinder length
        u = (b - a) / h
        vol = math.pi * r * r * h + 4.0 / 3.0 * math.pi * r ** 3
        m = g.density * vol
        ms = m * 4.0 * r / (4.0 * r + 3.0 * h)      # the two hemispheres
        mc = m - ms
        it = mc * (3 * r * r + h * h) / 12.0 + 0.4 * ms * r * r + ms * h * (3 * r + 2 * h) / 8.0
        ia = mc * r * r / 2.0 + 0.4 * ms * r * r
        I = it * np.eye(3) + (ia - it) * np.outer(u, u)
        return m, 0.5 * (a + b), I
    raise ValueError("geom kind %r carries no mass" % g.kind)


def body_inertial(geoms: Sequence[GeomSpec]):
    parts = [geom_inertial(g) for g in geoms if g.kind != "plane"]
    m = sum(p[0] for p in parts)
    com = sum(p[0] * p[1] for p in parts) / m
    I = np.zeros((3, 3))
    for mg, c, Ig in parts:
        d = c - com
        I += Ig + mg * (d @ d * np.eye(3) - np.outer(d, d))
    return m, com, I


# --------------------------------------------------------------------------
# compiled products
# --------------------------------------------------------------------------
@dataclass
class TreeModel:
    """Un-merged body tree in MuJoCo's own terms (world body excluded)."""
    parent: np.ndarray          # (nb,) int32, -1 = world
    pos: np.ndarray             # (nb,3) body frame offset in parent frame
    mass: np.ndarray            # (nb,)
    ipos: np.ndarray            # (nb,3) COM in body frame
    inertia: np.ndarray         # (nb,3,3) about COM, body axes
    jnt_body: np.ndarray        # (nv,) int32 body that owns dof j
    jnt_axis: np.ndarray        # (nv,3)
    jnt_range: np.ndarray       # (nv,2)
    jnt_limited: np.ndarray     # (nv,) int32
    armature: np.ndarray        # (nv,)
    damping: np.ndarray         # (nv,)
    gear: np.ndarray            # (nv,)
    ctrlrange: np.ndarray       # (nv,2)
    dof_invweight0: np.ndarray  # (nv,)
    timestep: float
    frame_skip: int
    solref: np.ndarray          # (2,)
    solimp: np.ndarray          # (5,)
    hand_body: int              # body carrying the "finger" site
    hand_pos: np.ndarray        # (3,) site offset in that body
    # one sphere-vs-plane contact pair (or radius < 0 when absent)
    con_body: int
    con_pos: np.ndarray         # (3,) sphere centre in con_body frame
    con_radius: float
    con_plane_z: float
    con_margin: float
    con_invweight: float        # body_invweight0 (translational) sum of the pair

    @property
    def nb(self):
        return len(self.mass)

    @property
    def nv(self):
        return len(self.armature)


# Flat layout of ChainModel.params (doubles).  Mirrored by csrc/chain_model.h.
CH_NJ = 7
CH_OFF = 0                      # 7*3 link offsets o_i in parent link frame
CH_MASS = CH_OFF + 21           # 7
CH_COM = CH_MASS + 7            # 7*3 COM in link frame
CH_INERTIA = CH_COM + 21        # 7*6 (xx,yy,zz,xy,xz,yz) about COM, link axes
CH_ARMATURE = CH_INERTIA + 42   # 7
CH_DAMPING = CH_ARMATURE + 7    # 7
CH_GEAR = CH_DAMPING + 7        # 7
CH_CTRL_LO = CH_GEAR + 7        # 7
CH_CTRL_HI = CH_CTRL_LO + 7     # 7
CH_RANGE_LO = CH_CTRL_HI + 7    # 7
CH_RANGE_HI = CH_RANGE_LO + 7   # 7
CH_INVW0 = CH_RANGE_HI + 7      # 7 dof_invweight0
CH_SCALARS = CH_INVW0 + 7       # see below
#   +0 timestep, +1 solref K, +2 solref B, +3 solimp d0, +4 solimp dwidth,
#   +5 solimp width, +6 solimp midpoint, +7 solimp power,
#   +8 hand offset x, +9 y, +10 z (in last link frame)
#   +11 con sphere x, +12 y, +13 z (last link frame), +14 radius (<0: none)
#   +15 plane z, +16 margin, +17 contact invweight
#   +18 limited mask (bit j set = joint j limited), +19 frame_skip
CH_NPARAM = CH_SCALARS + 20     # = 167


@dataclass
class ChainModel:
    axes: Tuple[int, ...]       # per joint 0/1/2 = x/y/z (compile-time pattern of the kernel)
    params: np.ndarray          # (CH_NPARAM,) float64

    def copy(self):
        return ChainModel(self.axes, self.params.copy())


@dataclass
class CompiledModel:
    spec: ModelSpec
    tree: TreeModel
    chain: ChainModel
    body_names: List[str]
    joint_names: List[str]
    # chain link index that absorbed each body (welded bodies map to their parent's link)
    body_to_link: List[int]


def _rot_axis(axis, q):
    """Rotation matrix about a unit axis (Rodrigues)."""
    a = np.asarray(axis, float)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + math.sin(q) * K + (1 - math.cos(q)) * (K @ K)


def forward_kinematics(tree: "TreeModel", q):
    """World positions / rotations of every body, the hand site and the contact sphere centre
    (host-side helper for observations at the set state and for validating start states)."""
    nb = tree.nb
    dof = {int(b): j for j, b in enumerate(tree.jnt_body)}
    R, p = [None] * nb, [None] * nb
    for b in range(nb):
        pa = int(tree.parent[b])
        Rp = np.eye(3) if pa < 0 else R[pa]
        pp = np.zeros(3) if pa < 0 else p[pa]
        p[b] = pp + Rp @ tree.pos[b]
        R[b] = Rp @ _rot_axis(tree.jnt_axis[dof[b]], q[dof[b]]) if b in dof else Rp
    hand = p[tree.hand_body] + R[tree.hand_body] @ tree.hand_pos
    sphere = p[tree.con_body] + R[tree.con_body] @ tree.con_pos if tree.con_radius > 0 else None
    return dict(pos=p, rot=R, hand=hand, sphere=sphere)


def table_clearance(tree: "TreeModel", q):
    """Distance between the end-effector sphere and the table plane (inf without a contact pair)."""
    if tree.con_radius <= 0:
        return float("inf")
    return float(forward_kinematics(tree, q)["sphere"][2] - tree.con_plane_z - tree.con_radius)


def tree_mass_matrix(tree: "TreeModel", q: np.ndarray) -> np.ndarray:
    """Dense joint-space inertia (with armature) from geometric Jacobians.
    Host-side helper for the qpos0 constants; NOT used by the rollout."""
    nb, nv = tree.nb, tree.nv
    R = [None] * nb
    p = [None] * nb
    dof_of_body = {int(b): j for j, b in enumerate(tree.jnt_body)}
    axis_w = np.zeros((nv, 3))
    anchor = np.zeros((nv, 3))
    chain_dofs = [None] * nb
    for b in range(nb):
        pa = int(tree.parent[b])
        Rp = np.eye(3) if pa < 0 else R[pa]
        pp = np.zeros(3) if pa < 0 else p[pa]
        p[b] = pp + Rp @ tree.pos[b]
        R[b] = Rp
        chain_dofs[b] = [] if pa < 0 else list(chain_dofs[pa])
        if b in dof_of_body:
            j = dof_of_body[b]
            R[b] = Rp @ _rot_axis(tree.jnt_axis[j], q[j])
            axis_w[j] = Rp @ tree.jnt_axis[j]
            anchor[j] = p[b]
            chain_dofs[b].append(j)
    M = np.diag(tree.armature.astype(float))
    for b in range(nb):
        c = p[b] + R[b] @ tree.ipos[b]
        Iw = R[b] @ tree.inertia[b] @ R[b].T
        Jv = np.zeros((3, nv))
        Jw = np.zeros((3, nv))
        for j in chain_dofs[b]:
            Jw[:, j] = axis_w[j]
            Jv[:, j] = np.cross(axis_w[j], c - anchor[j])
        M += tree.mass[b] * Jv.T @ Jv + Jw.T @ Iw @ Jw
    return M


def _body_com_jacobian(tree: "TreeModel", q, body):
    nb, nv = tree.nb, tree.nv
    R = [None] * nb
    p = [None] * nb
    dof_of_body = {int(b): j for j, b in enumerate(tree.jnt_body)}
    axis_w = np.zeros((nv, 3))
    anchor = np.zeros((nv, 3))
    chain_dofs = [None] * nb
    for b in range(nb):
        pa = int(tree.parent[b])
        Rp = np.eye(3) if pa < 0 else R[pa]
        pp = np.zeros(3) if pa < 0 else p[pa]
        p[b] = pp + Rp @ tree.pos[b]
        R[b] = Rp
        chain_dofs[b] = [] if pa < 0 else list(chain_dofs[pa])
        if b in dof_of_body:
            j = dof_of_body[b]
            R[b] = Rp @ _rot_axis(tree.jnt_axis[j], q[j])
            axis_w[j] = Rp @ tree.jnt_axis[j]
            anchor[j] = p[b]
            chain_dofs[b].append(j)
    c = p[body] + R[body] @ tree.ipos[body]
    J = np.zeros((6, nv))
    for j in chain_dofs[body]:
        J[:3, j] = np.cross(axis_w[j], c - anchor[j])
        J[3:, j] = axis_w[j]
    return J


def compile_model(spec: ModelSpec) -> CompiledModel:
    bodies = spec.bodies
    nb = len(bodies)
    joints = [(i, b.joint) for i, b in enumerate(bodies) if b.joint is not None]
    nv = len(joints)
    gear = np.zeros(nv)
    ctrl = np.zeros((nv, 2))
    jname = [j.name for _, j in joints]
    for a in spec.actuators:
        k = jname.index(a.joint)
        gear[k] = a.gear
        ctrl[k] = a.ctrlrange if a.ctrllimited else (-np.inf, np.inf)

    mass = np.zeros(nb)
    ipos = np.zeros((nb, 3))
    inertia = np.zeros((nb, 3, 3))
    for i, b in enumerate(bodies):
        mass[i], ipos[i], inertia[i] = body_inertial(b.geoms)

    hand_body, hand_pos = -1, np.zeros(3)
    for i, b in enumerate(bodies):
        if "finger" in b.sites:
            hand_body, hand_pos = i, np.array(b.sites["finger"], float)
    con_body, con_pos, con_r = -1, np.zeros(3), -1.0
    for i, b in enumerate(bodies):
        for g in b.geoms:
            if g.contact and g.kind == "sphere":
                con_body, con_pos, con_r = i, np.array(g.pos, float), g.size
    plane_z = 0.0
    has_plane = False
    for g in spec.world_geoms:
        if g.kind == "plane" and g.contact:
            plane_z, has_plane = g.pos[2], True
    if not has_plane:
        con_r = -1.0

    tree = TreeModel(
        parent=np.array([b.parent for b in bodies], np.int32),
        pos=np.array([b.pos for b in bodies], float),
        mass=mass, ipos=ipos, inertia=inertia,
        jnt_body=np.array([i for i, _ in joints], np.int32),
        jnt_axis=np.array([j.axis for _, j in joints], float),
        jnt_range=np.array([j.range for _, j in joints], float),
        jnt_limited=np.array([int(j.limited) for _, j in joints], np.int32),
        armature=np.array([j.armature for _, j in joints], float),
        damping=np.array([j.damping for _, j in joints], float),
        gear=gear, ctrlrange=ctrl,
        dof_invweight0=np.zeros(nv),
        timestep=spec.timestep, frame_skip=spec.frame_skip,
        solref=np.array(spec.solref, float), solimp=np.array(spec.solimp, float),
        hand_body=hand_body, hand_pos=hand_pos,
        con_body=con_body, con_pos=con_pos, con_radius=con_r,
        con_plane_z=plane_z, con_margin=spec.geom_margin, con_invweight=0.0)
    finalize_constants(tree)
    chain, body_to_link = _merge_chain(spec, tree)
    return CompiledModel(spec, tree, chain, [b.name for b in bodies], jname, body_to_link)


def finalize_constants(tree: TreeModel) -> None:
    """qpos0-derived constants (MuJoCo ``mj_setConst``): dof_invweight0 = diag(M0^-1);
    body_invweight0 (translational) = mean of the first 3 diagonal entries of
    J M0^-1 J^T at the body COM.  Recomputed after dynamics randomisation is NOT what
    MuJoCo does (mujoco_py writes model arrays in place without re-running
    mj_setConst, reference gym_env_wrapper.py:413), so callers keep the originals."""
    q0 = np.zeros(tree.nv)
    M0 = tree_mass_matrix(tree, q0)
    Minv = np.linalg.inv(M0)
    tree.dof_invweight0 = np.diag(Minv).copy()
    if tree.con_radius > 0:
        J = _body_com_jacobian(tree, q0, tree.con_body)
        A = J @ Minv @ J.T
        tree.con_invweight = float((A[0, 0] + A[1, 1] + A[2, 2]) / 3.0)


def _merge_chain(spec: ModelSpec, tree: TreeModel):
    """Fold welded (joint-less) bodies into the link of their nearest jointed
    ancestor and flatten to the kernel's parameter block.  Requires a serial
    chain of hinge joints with axis-aligned axes."""
    bodies = spec.bodies
    nb = len(bodies)
    link_of = [-1] * nb
    off_in_link = [np.zeros(3) for _ in range(nb)]   # body frame origin in its link frame
    links = []   # list of dict(offset, parts[(m,com,I)])
    axes = []
    for i, b in enumerate(bodies):
        if b.joint is not None:
            if b.parent >= 0:
                pl = link_of[b.parent]
                if pl != len(links) - 1:
                    raise ValueError("kernel supports serial chains only")
                offset = off_in_link[b.parent] + np.array(b.pos, float)
            else:
                if links:
                    raise ValueError("kernel supports a single chain root")
                offset = np.array(b.pos, float)
            ax = np.array(b.joint.axis, float)
            k = int(np.argmax(np.abs(ax)))
            if not np.allclose(ax, np.eye(3)[k]):
                raise ValueError("kernel supports +x/+y/+z hinge axes only")
            axes.append(k)
            links.append(dict(offset=offset, parts=[]))
            link_of[i] = len(links) - 1
            off_in_link[i] = np.zeros(3)
        else:
            if b.parent < 0:
                raise ValueError("static root bodies are not supported")
            link_of[i] = link_of[b.parent]
            off_in_link[i] = off_in_link[b.parent] + np.array(b.pos, float)
        links[link_of[i]]["parts"].append(
            (tree.mass[i], off_in_link[i] + tree.ipos[i], tree.inertia[i]))
    nj = len(links)
    if nj != CH_NJ:
        raise ValueError("the sm_100a rollout kernel is built for %d joints" % CH_NJ)
    P = np.zeros(CH_NPARAM)
    for l, L in enumerate(links):
        m = sum(p[0] for p in L["parts"])
        com = sum(p[0] * p[1] for p in L["parts"]) / m
        I = np.zeros((3, 3))
        for mg, c, Ig in L["parts"]:
            d = c - com
            I += Ig + mg * (d @ d * np.eye(3) - np.outer(d, d))
        P[CH_OFF + 3 * l: CH_OFF + 3 * l + 3] = L["offset"]
        P[CH_MASS + l] = m
        P[CH_COM + 3 * l: CH_COM + 3 * l + 3] = com
        P[CH_INERTIA + 6 * l: CH_INERTIA + 6 * l + 6] = (
            I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2])
    P[CH_ARMATURE:CH_ARMATURE + 7] = tree.armature
    P[CH_DAMPING:CH_DAMPING + 7] = tree.damping
    P[CH_GEAR:CH_GEAR + 7] = tree.gear
    P[CH_CTRL_LO:CH_CTRL_LO + 7] = tree.ctrlrange[:, 0]
    P[CH_CTRL_HI:CH_CTRL_HI + 7] = tree.ctrlrange[:, 1]
    P[CH_RANGE_LO:CH_RANGE_LO + 7] = tree.jnt_range[:, 0]
    P[CH_RANGE_HI:CH_RANGE_HI + 7] = tree.jnt_range[:, 1]
    P[CH_INVW0:CH_INVW0 + 7] = tree.dof_invweight0
    S = CH_SCALARS
    P[S + 0] = tree.timestep
    P[S + 1], P[S + 2] = solref_to_kb(tree.solref, tree.solimp, tree.timestep)
    P[S + 3] = tree.solimp[0]
    P[S + 4] = tree.solimp[1]
    P[S + 5] = tree.solimp[2]
    P[S + 6] = tree.solimp[3]
    P[S + 7] = tree.solimp[4]
    if tree.hand_body != nb - 1 and link_of[tree.hand_body] != nj - 1:
        raise ValueError("hand site must live on the last link")
    P[S + 8:S + 11] = off_in_link[tree.hand_body] + tree.hand_pos
    if tree.con_radius > 0:
        if link_of[tree.con_body] != nj - 1:
            raise ValueError("contact sphere must live on the last link")
        P[S + 11:S + 14] = off_in_link[tree.con_body] + tree.con_pos
    P[S + 14] = tree.con_radius
    P[S + 15] = tree.con_plane_z
    P[S + 16] = tree.con_margin
    P[S + 17] = tree.con_invweight
    P[S + 18] = float(sum(1 << j for j in range(nj) if tree.jnt_limited[j]))
    P[S + 19] = float(tree.frame_skip)
    return ChainModel(tuple(axes), P), link_of


def solref_to_kb(solref, solimp, timestep):
    """Stiffness / damping of the constraint reference acceleration
    (MuJoCo ``mj_makeImpedance``; REFSAFE clamps the time constant to 2*timestep)."""
    tc = max(float(solref[0]), 2.0 * timestep)
    dr = float(solref[1])
    dmax = float(solimp[1])
    K = 1.0 / max(1e-15, dmax * dmax * tc * tc * dr * dr)
    B = 2.0 / max(1e-15, dmax * tc)
    return K, B


# --------------------------------------------------------------------------
# dynamics randomisation (reference: gym_env_wrapper.py:367-416)
# --------------------------------------------------------------------------
def randomized_copy(cm: CompiledModel, param_dict: dict, rng: np.random.RandomState,
                    defaults: Optional[dict] = None):
    """Return (CompiledModel', defaults, randomized) with the named model fields
    perturbed exactly as the reference does: ``biased = (1+bias)*default``,
    ``value ~ U(biased*(1-noise), biased*(1+noise))``, written into the model
    arrays in place (qpos0 constants are NOT recomputed -- neither does mujoco_py).

    Supported fields on this model family: body_mass, body_inertia (principal
    moments, body axes), dof_damping.  dof_frictionloss / geom_* / sensor_noise
    are accepted by the reference but have no effect on (or do not exist in) the
    reacher model; a non-zero request for them raises."""
    spec = cm.spec
    tree = copy.deepcopy(cm.tree)
    defaults = {} if defaults is None else defaults
    randomized = {}
    for param_id, entries in param_dict.items():
        defaults.setdefault(param_id, {})
        randomized.setdefault(param_id, {})
        for name, (noise_scale, bias_scale) in entries.items():
            if param_id == "body_mass":
                idx = cm.body_names.index(name)
                cur = tree.mass[idx]
            elif param_id == "body_inertia":
                idx = cm.body_names.index(name)
                cur = np.linalg.eigvalsh(tree.inertia[idx])[::-1]   # MuJoCo orders principal moments descending
            elif param_id == "dof_damping":
                idx = cm.joint_names.index(name)
                cur = tree.damping[idx]
            elif param_id in ("dof_frictionloss", "geom_size", "geom_friction", "sensor_noise"):
                if noise_scale != 0.0 or bias_scale != 0.0:
                    raise ValueError("dynamics field %s is not modelled by the GPU rollout" % param_id)
                # the reference still draws for such an entry (gym_env_wrapper.py:406-408: one uniform per element of the
                # field -- geom_size and geom_friction are 3-vectors): consume the same draws, or every parameter
                # listed after it would get a different value than the reference's worker with this seed
                rng.uniform(size=3 if param_id in ("geom_size", "geom_friction") else None)
                continue
            else:
                raise ValueError("Unknown dynamics field")
            if name not in defaults[param_id]:
                defaults[param_id][name] = copy.deepcopy(cur)
            cur = copy.deepcopy(defaults[param_id][name])
            biased = (1.0 + bias_scale) * cur
            val = rng.uniform(biased - biased * noise_scale, biased + biased * noise_scale)
            randomized[param_id][name] = val
            if param_id == "body_mass":
                tree.mass[idx] = val
            elif param_id == "dof_damping":
                tree.damping[idx] = val
            else:
                w, V = np.linalg.eigh(tree.inertia[idx])
                V = V[:, ::-1]
                tree.inertia[idx] = V @ np.diag(np.atleast_1d(val)) @ V.T
    chain, link_of = _merge_chain(spec, tree)
    return CompiledModel(spec, tree, chain, cm.body_names, cm.joint_names, link_of), defaults, randomized
