"""MJCF-subset compiler for the runtime-parameterised tree rollout kernel (``csrc/rollout_tree.cu``).

SURVEY §8(f-3): the reference's other MuJoCo models (``mjmpc/envs/assets/xml/swimmer.xml``,
``half_cheetah.xml``) go through MuJoCo's own XML compiler inside ``mujoco_py`` (``gym.envs.mujoco.MujocoEnv``,
reference ``mjmpc/envs/basic/swimmer.py:7``).  This module restates the part of that compiler a kinematic tree of
hinge / slide joints needs and lowers the result to the kernel's parameter block.

Subset (anything else raises ``UnsupportedMjcf`` -- an unsupported model is rejected, never simulated wrongly):
  compiler   angle radian|degree, coordinate local, inertiafromgeom true|auto|false (+ <inertial>), settotalmass
  option     timestep, gravity, density, viscosity, integrator Euler (wind must be zero)
  default    nested classes for joint / geom / motor / site; ``childclass`` and ``class`` attributes
  body       pos, quat | axisangle | euler; any number of hinge / slide joints per body, any axis, any anchor
  joint      type hinge|slide, pos, axis, range + limited, damping, armature, stiffness + springref, ref = 0,
             solreflimit / solimplimit
  geom       sphere, capsule (fromto or size + pos/quat/axisangle), cylinder, box, plane (ignored): mass properties
  actuator   motor: joint, gear (first component), ctrlrange + ctrllimited
  contacts   opt-in.  The default (``allow_contacts="error"``) refuses a model whose geoms can collide; ``"model"`` keeps
             world-plane / capsule / sphere pairs as contact candidates (contype / conaffinity test, same-assembly and
             parent-child filters, condim 3, margin = gap = 0, one solref / solimp per pair, friction = the larger of the
             two) with ``body_invweight0`` -- simulated by the planar instantiation of the kernel; ``"ignore"`` drops
             them and says so in ``TreeModel.ignored``
Not in the subset: frictionloss, tendons, equality constraints, free / ball joints, RK4, boxes / meshes as colliders.

MuJoCo semantics restated here [EXT, MuJoCo 2.0 documentation: "XML reference", "Computation"]:
  * geom mass properties at density 1000 unless given; body frame = first-moment / parallel-axis composition of its
    geoms; the body inertial frame is the geom frame for a single geom, else the principal axes;
  * bodies without joints are welded into their parent; world-fixed bodies drop out;
  * dof_invweight0[i] = (M(qpos0)^-1)_ii  (hinge / slide: one dof per joint), the regulariser scale of limit rows;
  * solref (timeconst, dampratio) -> (K, B) with timeconst clamped to 2 * timestep (REFSAFE).
"""
from __future__ import annotations

import math
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np


class UnsupportedMjcf(ValueError):
    pass


HINGE, SLIDE = 0, 1


def _f(s, n=None):
    v = np.array([float(x) for x in str(s).split()], float)
    if n is not None and v.size != n:
        raise UnsupportedMjcf("expected %d numbers, got %r" % (n, s))
    return v


def quat_to_mat(q):
    q = np.asarray(q, float)
    q = q / np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def axisangle_to_mat(axis, angle):
    a = np.asarray(axis, float)
    a = a / np.linalg.norm(a)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + math.sin(angle) * K + (1 - math.cos(angle)) * (K @ K)


def _z_to(v):
    """Rotation taking the z axis onto unit vector v (MuJoCo's fromto / zaxis convention: minimal rotation)."""
    v = np.asarray(v, float) / np.linalg.norm(v)
    z = np.array([0.0, 0.0, 1.0])
    c = float(z @ v)
    ax = np.cross(z, v)
    s = np.linalg.norm(ax)
    if s < 1e-12:
        return np.eye(3) if c > 0 else np.diag([1.0, -1.0, -1.0])
    return axisangle_to_mat(ax / s, math.atan2(s, c))


@dataclass
class TreeModel:
    """MuJoCo-style compiled model of a hinge / slide tree (world = body 0 is implicit: parent -1 = world)."""
    name: str
    timestep: float
    gravity: np.ndarray
    density: float
    viscosity: float
    # bodies that move (welded bodies already merged), topological order
    body_parent: np.ndarray       # (nb,) int, -1 = world
    body_pos: np.ndarray          # (nb, 3) in the parent body frame
    body_mat: np.ndarray          # (nb, 3, 3) body frame in the parent body frame
    body_mass: np.ndarray         # (nb,)
    body_ipos: np.ndarray         # (nb, 3) centre of mass in the body frame
    body_imat: np.ndarray         # (nb, 3, 3) inertial frame in the body frame (columns = principal axes)
    body_inertia: np.ndarray      # (nb, 3) principal moments
    body_names: List[str]
    # joints = dofs
    jnt_type: np.ndarray          # (nv,) HINGE | SLIDE
    jnt_body: np.ndarray          # (nv,)
    jnt_pos: np.ndarray           # (nv, 3) anchor, body frame
    jnt_axis: np.ndarray          # (nv, 3) unit, body frame
    jnt_limited: np.ndarray       # (nv,) bool
    jnt_range: np.ndarray         # (nv, 2)
    jnt_damping: np.ndarray
    jnt_armature: np.ndarray
    jnt_stiffness: np.ndarray
    jnt_springref: np.ndarray
    jnt_solref: np.ndarray        # (nv, 2)
    jnt_solimp: np.ndarray        # (nv, 5)
    jnt_names: List[str]
    # actuators
    act_dof: np.ndarray           # (nu,)
    act_gear: np.ndarray
    act_ctrllimited: np.ndarray
    act_ctrlrange: np.ndarray     # (nu, 2)
    dof_invweight0: np.ndarray = None
    ignored: List[str] = field(default_factory=list)
    shapes: list = field(default_factory=list)     # colliding capsules / spheres: (body, end 0, end 1, radius), body frame
    contacts: list = field(default_factory=list)   # candidate contact pairs (allow_contacts="model"), see compile_mjcf_root

    @property
    def nv(self):
        return int(self.jnt_type.size)

    @property
    def nb(self):
        return int(self.body_parent.size)

    @property
    def nu(self):
        return int(self.act_dof.size)


# ------------------------------------------------------------------------------------------------ defaults
class _Defaults:
    """MJCF default classes: a class inherits its enclosing class; elements merge attribute-wise."""

    def __init__(self, root):
        self.cls: Dict[str, Dict[str, Dict[str, str]]] = {"main": {}}
        for d in root.findall("default"):
            self._read(d, "main", top=True)

    def _read(self, node, parent, top=False):
        name = node.get("class", "main" if top else None)
        if name is None:
            raise UnsupportedMjcf("nested <default> without class")
        table = {k: dict(v) for k, v in self.cls.get(parent, {}).items()} if name != "main" else self.cls["main"]
        for e in node:
            if e.tag == "default":
                continue
            table.setdefault(e.tag, {}).update(e.attrib)
        self.cls[name] = table
        for d in node.findall("default"):
            self._read(d, name)

    def resolve(self, elem, childclass):
        cls = elem.get("class", childclass or "main")
        if cls not in self.cls:
            raise UnsupportedMjcf("unknown default class %r" % cls)
        a = dict(self.cls[cls].get(elem.tag, {}))
        a.update({k: v for k, v in elem.attrib.items() if k != "class"})
        return a


# ------------------------------------------------------------------------------------------------ geoms
def _geom_mass_props(a, angle_scale):
    """(mass, centre, inertia about the centre, frame) of one geom in the body frame."""
    kind = a.get("type", "sphere")
    if kind == "plane":
        return None
    size = _f(a.get("size", "0"))
    R = _orientation(a, angle_scale)
    pos = _f(a.get("pos", "0 0 0"), 3)
    if "fromto" in a:
        if kind not in ("capsule", "cylinder", "box"):
            raise UnsupportedMjcf("fromto on a %s" % kind)
        ft = _f(a["fromto"], 6)
        pos = 0.5 * (ft[:3] + ft[3:])
        half = 0.5 * float(np.linalg.norm(ft[3:] - ft[:3]))
        R = _z_to(ft[3:] - ft[:3])
        size = np.array([size[0], half]) if kind != "box" else np.array([size[0], size[0], half])
    r = float(size[0])
    if kind == "sphere":
        vol = 4.0 / 3.0 * math.pi * r ** 3
        unit = np.full(3, 0.4 * r * r)
    elif kind == "capsule":
        h = 2.0 * float(size[1])
        vc, vs = math.pi * r * r * h, 4.0 / 3.0 * math.pi * r ** 3
        vol = vc + vs
        axial = (0.5 * vc * r * r + 0.4 * vs * r * r) / vol
        trans = (vc * (3 * r * r + h * h) / 12.0 + vs * (0.4 * r * r + 0.25 * h * h + 0.375 * r * h)) / vol
        unit = np.array([trans, trans, axial])
    elif kind == "cylinder":
        h = 2.0 * float(size[1])
        vol = math.pi * r * r * h
        unit = np.array([(3 * r * r + h * h) / 12.0, (3 * r * r + h * h) / 12.0, 0.5 * r * r])
    elif kind == "box":
        x, y, z = (2.0 * size[:3]).tolist()
        vol = x * y * z
        unit = np.array([(y * y + z * z) / 12.0, (x * x + z * z) / 12.0, (x * x + y * y) / 12.0])
    else:
        raise UnsupportedMjcf("geom type %s" % kind)
    mass = float(a["mass"]) if "mass" in a else float(a.get("density", "1000")) * vol
    return mass, pos, R @ np.diag(mass * unit) @ R.T, R


def _collision_record(a, mp, angle_scale):
    """What collision detection needs from one geom, or None when it cannot collide (contype = conaffinity = 0)."""
    ct, ca = int(a.get("contype", "1")), int(a.get("conaffinity", "1"))
    if ct == 0 and ca == 0:
        return None
    kind = a.get("type", "sphere")
    size = _f(a.get("size", "0"))
    rec = dict(kind=kind, contype=ct, conaffinity=ca, condim=int(a.get("condim", "3")),
               friction=float(_f(a.get("friction", "1 0.005 0.0001"))[0]), solref=_f(a.get("solref", "0.02 1"), 2),
               solimp=_solimp(a.get("solimp")), margin=float(a.get("margin", "0")), gap=float(a.get("gap", "0")),
               name=a.get("name", ""))
    if kind == "plane":
        R = _orientation(a, angle_scale)
        rec.update(p0=_f(a.get("pos", "0 0 0"), 3), p1=R[:, 2].copy(), radius=0.0)     # a point of the plane, its normal
    elif kind in ("capsule", "sphere"):
        if "fromto" in a:
            ft = _f(a["fromto"], 6)
            rec.update(p0=ft[:3], p1=ft[3:], radius=float(size[0]))
        else:
            half = float(size[1]) if kind == "capsule" else 0.0
            rec.update(p0=mp[1] - mp[3][:, 2] * half, p1=mp[1] + mp[3][:, 2] * half, radius=float(size[0]))
    else:
        rec.update(p0=None, p1=None, radius=0.0)            # a colliding shape the subset has no detection for
    rec["com"] = None
    return rec


def _orientation(a, angle_scale):
    if "quat" in a:
        return quat_to_mat(_f(a["quat"], 4))
    if "axisangle" in a:
        v = _f(a["axisangle"], 4)
        return axisangle_to_mat(v[:3], v[3] * angle_scale)
    if "euler" in a:
        e = _f(a["euler"], 3) * angle_scale        # default eulerseq "xyz", intrinsic
        R = np.eye(3)
        for ax, ang in zip(np.eye(3), e):
            R = R @ axisangle_to_mat(ax, ang)
        return R
    if "zaxis" in a:
        return _z_to(_f(a["zaxis"], 3))
    if "xyaxes" in a:
        raise UnsupportedMjcf("xyaxes orientation")
    return np.eye(3)


# ------------------------------------------------------------------------------------------------ compile
def compile_mjcf(path: str, allow_contacts: str = "error") -> TreeModel:
    """Compile an MJCF file."""
    return compile_mjcf_root(ET.parse(path).getroot(), allow_contacts)


def compile_mjcf_string(text: str, allow_contacts: str = "error") -> TreeModel:
    return compile_mjcf_root(ET.fromstring(text), allow_contacts)


def compile_mjcf_root(root, allow_contacts: str = "error") -> TreeModel:
    comp = root.find("compiler")
    comp = dict(comp.attrib) if comp is not None else {}
    if comp.get("coordinate", "local") != "local":
        raise UnsupportedMjcf("coordinate=global")
    angle_scale = 1.0 if comp.get("angle", "degree") == "radian" else math.pi / 180.0
    from_geom = comp.get("inertiafromgeom", "auto")
    opt = root.find("option")
    opt = dict(opt.attrib) if opt is not None else {}
    if opt.get("integrator", "Euler") != "Euler":
        raise UnsupportedMjcf("integrator %s" % opt["integrator"])
    if "wind" in opt and np.any(_f(opt["wind"]) != 0):
        raise UnsupportedMjcf("wind")
    for tag in ("tendon", "equality", "contact", "sensor", "keyframe"):
        if root.find(tag) is not None and len(root.find(tag)):
            raise UnsupportedMjcf("<%s>" % tag)
    timestep = float(opt.get("timestep", "0.002"))
    defaults = _Defaults(root)
    ignored: List[str] = []

    raw = []     # every body of the file: dict(parent, pos, mat, joints, parts, name)

    def walk(e, parent, childclass):
        for be in e.findall("body"):
            cc = be.get("childclass", childclass)
            joints = []
            for je in be.findall("joint"):
                a = defaults.resolve(je, cc)
                kind = a.get("type", "hinge")
                if kind not in ("hinge", "slide"):
                    raise UnsupportedMjcf("joint type %s" % kind)
                if float(a.get("frictionloss", "0")) != 0.0:
                    raise UnsupportedMjcf("frictionloss")
                if float(a.get("ref", "0")) != 0.0:
                    raise UnsupportedMjcf("joint ref")
                axis = _f(a.get("axis", "0 0 1"), 3)
                rng = _f(a.get("range", "0 0"), 2) * (angle_scale if kind == "hinge" else 1.0)
                joints.append(dict(
                    name=a.get("name", "joint%d" % (len(raw) * 8 + len(joints))), type=HINGE if kind == "hinge" else SLIDE,
                    pos=_f(a.get("pos", "0 0 0"), 3), axis=axis / np.linalg.norm(axis),
                    limited=a.get("limited", "false") == "true", range=rng,
                    damping=float(a.get("damping", "0")), armature=float(a.get("armature", "0")),
                    stiffness=float(a.get("stiffness", "0")),
                    springref=float(a.get("springref", "0")) * (angle_scale if kind == "hinge" else 1.0),
                    solref=_f(a.get("solreflimit", "0.02 1"), 2),
                    solimp=_solimp(a.get("solimplimit"))))
            if be.find("freejoint") is not None:
                raise UnsupportedMjcf("free joint")
            parts, collides, shapes = [], False, []
            for ge in be.findall("geom"):
                a = defaults.resolve(ge, cc)
                mp = _geom_mass_props(a, angle_scale)
                if mp is not None:
                    parts.append(mp)
                rec = _collision_record(a, mp, angle_scale)
                if rec is not None:
                    collides = True
                    shapes.append(rec)
            inertial = be.find("inertial")
            if inertial is not None and from_geom != "true":
                a = inertial.attrib
                if "fullinertia" in a:
                    raise UnsupportedMjcf("fullinertia")
                Ri = _orientation(a, angle_scale)
                parts = [(float(a["mass"]), _f(a["pos"], 3), Ri @ np.diag(_f(a["diaginertia"], 3)) @ Ri.T, Ri)]
            elif from_geom == "false":
                parts = []
            mtot = sum(p_[0] for p_ in parts)
            own_com = sum(p_[0] * p_[1] for p_ in parts) / mtot if mtot > 0 else np.zeros(3)
            for rec in shapes:
                rec["com"] = own_com                       # body_invweight0 is taken at the centre of mass of the geom's own body
            raw.append(dict(parent=parent, pos=_f(be.get("pos", "0 0 0"), 3), mat=_orientation(be.attrib, angle_scale),
                            joints=joints, parts=parts, name=be.get("name", "body%d" % len(raw)), collides=collides,
                            shapes=shapes))
            walk(be, len(raw) - 1, cc)

    wb = root.find("worldbody")
    walk(wb, -1, None)

    world_shapes = [r for r in (_collision_record(defaults.resolve(g, None), None, angle_scale) if
                                defaults.resolve(g, None).get("type", "sphere") == "plane" else
                                _collision_record(defaults.resolve(g, None), _geom_mass_props(defaults.resolve(g, None), angle_scale),
                                                  angle_scale) for g in wb.findall("geom")) if r is not None]

    # ---- weld joint-less bodies into their parents; drop world-fixed ones
    nraw = len(raw)
    frame_parent = [None] * nraw      # index of the moving body this raw body belongs to (-1 = world)
    frame_pos = [None] * nraw         # its frame expressed in that moving body's frame
    frame_mat = [None] * nraw
    moving = []
    for i, b in enumerate(raw):
        p = b["parent"]
        if p < 0:
            hp, hpos, hmat = -1, b["pos"], b["mat"]
        else:
            hp = frame_parent[p] if not raw[p]["joints"] else p
            if raw[p]["joints"]:
                hpos, hmat = b["pos"], b["mat"]
            else:
                hpos, hmat = frame_pos[p] + frame_mat[p] @ b["pos"], frame_mat[p] @ b["mat"]
        if b["joints"]:
            moving.append(i)
            b["host_parent"], b["host_pos"], b["host_mat"] = hp, hpos, hmat
            frame_parent[i], frame_pos[i], frame_mat[i] = i, np.zeros(3), np.eye(3)
        else:
            frame_parent[i], frame_pos[i], frame_mat[i] = hp, hpos, hmat
    index = {r: k for k, r in enumerate(moving)}
    nb = len(moving)
    if nb == 0:
        raise UnsupportedMjcf("no joints")
    parts_of = {r: [] for r in moving}
    fluid = float(opt.get("density", "0")) > 0 or float(opt.get("viscosity", "0")) > 0
    for i, b in enumerate(raw):
        host = frame_parent[i]
        if host < 0:
            continue
        if fluid and host != i and b["parts"]:
            # MuJoCo keeps a welded body as a body of its own, with its own inertia box in mj_passive's fluid model
            raise UnsupportedMjcf("body %r is welded to %r in a fluid (per-body inertia boxes)" % (b["name"], raw[host]["name"]))
        for (m, c, I, R) in b["parts"]:
            parts_of[host].append((m, frame_pos[i] + frame_mat[i] @ c, frame_mat[i] @ I @ frame_mat[i].T, frame_mat[i] @ R))

    geoms = []           # colliding geoms: record + moving body (-1 = world) with coordinates in that body's frame
    for rec in world_shapes:
        geoms.append(dict(rec, body=-1))
    for i, b in enumerate(raw):
        fp, fm = (frame_pos[i], frame_mat[i])
        for rec in b["shapes"]:
            g = dict(rec, body=index[frame_parent[i]] if frame_parent[i] >= 0 else -1)
            if rec["p0"] is not None:
                g["p0"] = fp + fm @ rec["p0"]
                g["p1"] = fm @ rec["p1"] if rec["kind"] == "plane" else fp + fm @ rec["p1"]
                g["com"] = fp + fm @ rec["com"]
            geoms.append(g)
    shapes = [(g["body"], g["p0"], g["p1"], g["radius"]) for g in geoms if g["body"] >= 0 and g["kind"] in ("capsule", "sphere")]
    bp = [index.get(raw[r]["host_parent"], -1) for r in moving]
    pairs = []
    for i in range(len(geoms)):
        for j in range(i + 1, len(geoms)):
            a, b = geoms[i], geoms[j]
            if a["body"] == b["body"]:
                continue                                         # same rigid assembly (or both fixed to the world)
            if a["body"] >= 0 and b["body"] >= 0 and (bp[a["body"]] == b["body"] or bp[b["body"]] == a["body"]):
                continue                                         # parent and child (the world is nobody's filtered parent)
            if not ((a["contype"] & b["conaffinity"]) or (b["contype"] & a["conaffinity"])):
                continue
            pairs.append((a, b) if (a["kind"] == "plane" or b["kind"] != "plane") else (b, a))    # the plane is geom 1
    contacts = []
    if pairs:
        msg = "%d geom pairs can collide" % len(pairs)
        if allow_contacts == "error":
            raise UnsupportedMjcf("contacts are outside the default subset: " + msg +
                                  " (allow_contacts='ignore' drops them, 'model' keeps plane / capsule / sphere pairs)")
        if allow_contacts == "ignore":
            ignored.append("contacts dropped: " + msg)
        else:
            for a, b in pairs:
                ok = a["kind"] in ("plane", "capsule", "sphere") and b["kind"] in ("capsule", "sphere") and \
                    (a["kind"] != "plane" or a["body"] < 0)
                if not ok:
                    raise UnsupportedMjcf("no collision detection for a %s against a %s" % (a["kind"], b["kind"]))
                if a["condim"] != 3 or b["condim"] != 3 or a["margin"] or b["margin"] or a["gap"] or b["gap"]:
                    raise UnsupportedMjcf("contacts: condim 3, margin 0, gap 0 only")
                if not (np.array_equal(a["solref"], b["solref"]) and np.array_equal(a["solimp"], b["solimp"])):
                    raise UnsupportedMjcf("contacts: geoms with different solref / solimp (solmix) are not supported")
                contacts.append(dict(kind="plane" if a["kind"] == "plane" else "capsule", body1=a["body"], body2=b["body"],
                                     a0=a["p0"], a1=a["p1"], ra=a["radius"], b0=b["p0"], b1=b["p1"], rb=b["radius"],
                                     com1=a.get("com"), com2=b.get("com"), mu=max(a["friction"], b["friction"]),
                                     solref=a["solref"].copy(), solimp=a["solimp"].copy()))
    body_parent = np.array([index.get(raw[r]["host_parent"], -1) for r in moving], int)
    body_pos = np.array([raw[r]["host_pos"] for r in moving])
    body_mat = np.array([raw[r]["host_mat"] for r in moving])
    mass, ipos, imat, inertia = np.zeros(nb), np.zeros((nb, 3)), np.zeros((nb, 3, 3)), np.zeros((nb, 3))
    for k, r in enumerate(moving):
        ps = parts_of[r]
        m = sum(p[0] for p in ps)
        if m <= 0:
            imat[k] = np.eye(3)
            continue
        c = sum(p[0] * p[1] for p in ps) / m
        I = np.zeros((3, 3))
        for (mi, ci, Ii, _) in ps:
            d = ci - c
            I += Ii + mi * ((d @ d) * np.eye(3) - np.outer(d, d))
        mass[k], ipos[k] = m, c
        if len(ps) == 1:
            imat[k] = ps[0][3]
            inertia[k] = np.diag(ps[0][3].T @ I @ ps[0][3])
        else:
            w, V = np.linalg.eigh(I)
            order = np.argsort(-w)                   # MuJoCo sorts principal moments in decreasing order
            V = V[:, order]
            if np.linalg.det(V) < 0:
                V[:, 2] = -V[:, 2]
            imat[k], inertia[k] = V, w[order]
    if "settotalmass" in comp:
        s = float(comp["settotalmass"]) / mass.sum()
        mass, inertia = mass * s, inertia * s

    J = [(index[r], j) for r in moving for j in raw[r]["joints"]]
    nv = len(J)
    names = [j["name"] for _, j in J]
    act_dof, gear, cl, cr = [], [], [], []
    act = root.find("actuator")
    for me in (list(act) if act is not None else []):
        if me.tag != "motor":
            raise UnsupportedMjcf("actuator %s" % me.tag)
        a = defaults.resolve(me, None)
        if a.get("joint") not in names:
            raise UnsupportedMjcf("motor on unknown joint %r" % a.get("joint"))
        act_dof.append(names.index(a["joint"]))
        gear.append(_f(a.get("gear", "1"))[0])
        cl.append(a.get("ctrllimited", "false") == "true")
        cr.append(_f(a.get("ctrlrange", "0 0"), 2))
    model = TreeModel(
        name=root.get("model", "model"), timestep=timestep, gravity=_f(opt.get("gravity", "0 0 -9.81"), 3),
        density=float(opt.get("density", "0")), viscosity=float(opt.get("viscosity", "0")),
        body_parent=body_parent, body_pos=body_pos, body_mat=body_mat, body_mass=mass, body_ipos=ipos, body_imat=imat,
        body_inertia=inertia, body_names=[raw[r]["name"] for r in moving],
        jnt_type=np.array([j["type"] for _, j in J], int), jnt_body=np.array([b for b, _ in J], int),
        jnt_pos=np.array([j["pos"] for _, j in J]), jnt_axis=np.array([j["axis"] for _, j in J]),
        jnt_limited=np.array([j["limited"] for _, j in J], bool), jnt_range=np.array([j["range"] for _, j in J]),
        jnt_damping=np.array([j["damping"] for _, j in J]), jnt_armature=np.array([j["armature"] for _, j in J]),
        jnt_stiffness=np.array([j["stiffness"] for _, j in J]), jnt_springref=np.array([j["springref"] for _, j in J]),
        jnt_solref=np.array([j["solref"] for _, j in J]), jnt_solimp=np.array([j["solimp"] for _, j in J]),
        jnt_names=names, act_dof=np.array(act_dof, int), act_gear=np.array(gear, float),
        act_ctrllimited=np.array(cl, bool), act_ctrlrange=np.array(cr, float).reshape(-1, 2), ignored=ignored,
        shapes=shapes, contacts=contacts)
    M0 = mass_matrix(model, np.zeros(nv))
    Minv = np.linalg.inv(M0)
    model.dof_invweight0 = np.diag(Minv).copy()
    if contacts:
        # body_invweight0 (translational) = tr(Jp M^-1 Jp') / 3 at the centre of mass of the geom's body, qpos0 (mj_setConst)
        xpos, xmat, anchor, axis = kinematics(model, np.zeros(nv))
        for c in contacts:
            w = 0.0
            for body, com in ((c["body1"], c["com1"]), (c["body2"], c["com2"])):
                if body < 0:
                    continue
                Jp = point_jacobian(model, body, xpos[body] + xmat[body] @ com, anchor, axis)
                w += float(np.trace(Jp @ Minv @ Jp.T)) / 3.0
            c["invweight"] = w
    return model


def _solimp(s):
    """solimp (d0, dmax, width, midpoint, power) with MuJoCo's clamps [EXT: getsolparam]: impedances and the midpoint
    into [mjMINIMP, mjMAXIMP] = [0.0001, 0.9999], width >= 0, power >= 1 (half_cheetah.xml asks for d0 = 0)."""
    v = np.array([0.9, 0.95, 0.001, 0.5, 2.0])
    if s is not None:
        g = _f(s)
        v[:g.size] = g
    v[0], v[1], v[3] = (min(0.9999, max(0.0001, x)) for x in (v[0], v[1], v[3]))
    v[2], v[4] = max(0.0, v[2]), max(1.0, v[4])
    return v


# ------------------------------------------------------------------------------------------------ qpos0 quantities
def kinematics(model: TreeModel, q):
    """Body frames (origin, rotation), joint anchors and axes in the world frame (MuJoCo's mj_kinematics order:
    the joints of a body act one after the other on the body frame)."""
    nb = model.nb
    xpos, xmat = np.zeros((nb, 3)), np.zeros((nb, 3, 3))
    anchor, axis = np.zeros((model.nv, 3)), np.zeros((model.nv, 3))
    for b in range(nb):
        p = model.body_parent[b]
        Pp, Rp = (np.zeros(3), np.eye(3)) if p < 0 else (xpos[p], xmat[p])
        pos, R = Pp + Rp @ model.body_pos[b], Rp @ model.body_mat[b]
        for j in np.nonzero(model.jnt_body == b)[0]:
            anchor[j], axis[j] = pos + R @ model.jnt_pos[j], R @ model.jnt_axis[j]
            if model.jnt_type[j] == SLIDE:
                pos = pos + axis[j] * q[j]
            else:
                R = R @ axisangle_to_mat(model.jnt_axis[j], q[j])
                pos = anchor[j] - R @ model.jnt_pos[j]
        xpos[b], xmat[b] = pos, R
    return xpos, xmat, anchor, axis


def point_jacobian(model: TreeModel, body: int, point, anchor, axis):
    """Translational Jacobian (3, nv) of a world point moving with ``body``."""
    Jp = np.zeros((3, model.nv))
    a = body
    while a >= 0:
        for j in np.nonzero(model.jnt_body == a)[0]:
            Jp[:, j] = axis[j] if model.jnt_type[j] == SLIDE else np.cross(axis[j], point - anchor[j])
        a = model.body_parent[a]
    return Jp


def mass_matrix(model: TreeModel, q):
    """M(q) from geometric Jacobians (used once, at qpos0, for dof_invweight0)."""
    xpos, xmat, anchor, axis = kinematics(model, q)
    nv = model.nv
    M = np.diag(model.jnt_armature.astype(float))
    for b in range(model.nb):
        if model.body_mass[b] <= 0:
            continue
        com = xpos[b] + xmat[b] @ model.body_ipos[b]
        Ri = xmat[b] @ model.body_imat[b]
        Iw = Ri @ np.diag(model.body_inertia[b]) @ Ri.T
        Jv, Jw = np.zeros((3, nv)), np.zeros((3, nv))
        a = b
        while a >= 0:
            for j in np.nonzero(model.jnt_body == a)[0]:
                if model.jnt_type[j] == SLIDE:
                    Jv[:, j] = axis[j]
                else:
                    Jv[:, j], Jw[:, j] = np.cross(axis[j], com - anchor[j]), axis[j]
            a = model.body_parent[a]
        M += model.body_mass[b] * Jv.T @ Jv + Jw.T @ Iw @ Jw
    return M


def _segment_distance(p1, q1, p2, q2):
    """Distance between segments p1-q1 and p2-q2 (closest points by clamping, Ericson 5.1.9)."""
    d1, d2, r = q1 - p1, q2 - p2, p1 - p2
    a, e, f = d1 @ d1, d2 @ d2, d2 @ r
    if a <= 1e-18 and e <= 1e-18:
        return float(np.linalg.norm(r))
    if a <= 1e-18:
        s, t = 0.0, min(max(f / e, 0.0), 1.0)
    else:
        c = d1 @ r
        if e <= 1e-18:
            t, s = 0.0, min(max(-c / a, 0.0), 1.0)
        else:
            b = d1 @ d2
            den = a * e - b * b
            s = min(max((b * f - c * e) / den, 0.0), 1.0) if den > 1e-18 else 0.0
            t = (b * s + f) / e
            if t < 0.0:
                t, s = 0.0, min(max(-c / a, 0.0), 1.0)
            elif t > 1.0:
                t, s = 1.0, min(max((b - c) / a, 0.0), 1.0)
    return float(np.linalg.norm(p1 + d1 * s - p2 - d2 * t))


def self_clearance(model: TreeModel, q) -> float:
    """Smallest gap between colliding shapes of bodies that are not parent and child (MuJoCo filters those pairs):
    positive = the configuration is inside the contact-free subset the kernel simulates.  inf without such pairs."""
    xpos, xmat, _, _ = kinematics(model, np.asarray(q, float))
    world = [(b, xpos[b] + xmat[b] @ p0, xpos[b] + xmat[b] @ p1, r) for (b, p0, p1, r) in model.shapes]
    best = np.inf
    for i in range(len(world)):
        for j in range(i + 1, len(world)):
            bi, bj = world[i][0], world[j][0]
            if bi == bj or model.body_parent[bi] == bj or model.body_parent[bj] == bi:
                continue
            best = min(best, _segment_distance(world[i][1], world[i][2], world[j][1], world[j][2]) - world[i][3] - world[j][3])
    return best


def solref_to_kb(solref, solimp, timestep):
    tc = max(float(solref[0]), 2.0 * timestep)
    dr, dmax = float(solref[1]), float(solimp[1])
    return 1.0 / max(1e-15, dmax * dmax * tc * tc * dr * dr), 2.0 / max(1e-15, dmax * tc)


# ------------------------------------------------------------------------------------------------ kernel parameters
# One LINK per dof (a body with several joints becomes a run of links, all but the last massless).  Per link, doubles:
LK_RFIX, LK_OFF, LK_AXIS, LK_MASS, LK_COM, LK_IC, LK_RIN, LK_BOX = 0, 9, 12, 15, 16, 19, 25, 34
LK_ARM, LK_DAMP, LK_STIFF, LK_SREF, LK_LO, LK_HI, LK_INVW, LK_SOLK, LK_SOLB, LK_SOLIMP = 37, 38, 39, 40, 41, 42, 43, 44, 45, 46
LK_GEAR, LK_CLO, LK_CHI, LK_STRIDE = 51, 52, 53, 54
# per link, ints: parent link, joint type, limited, actuator index (-1 = none), flags (1: carries a body, 2: Rfix = I)
LI_PARENT, LI_TYPE, LI_LIMITED, LI_ACT, LI_BODY, LI_STRIDE = 0, 1, 2, 3, 4, 5
# globals (doubles): timestep, gravity[3], density, viscosity
G_DT, G_GRAV, G_RHO, G_VISC, G_STRIDE = 0, 1, 4, 5, 6
MAX_LINKS = 12


def pack_links(model: TreeModel):
    """Lower the body / joint tree to the kernel's link arrays (see ``csrc/tree_model.h`` for the same offsets)."""
    nv = model.nv
    if nv > MAX_LINKS:
        raise UnsupportedMjcf("%d dofs > %d" % (nv, MAX_LINKS))
    P = np.zeros((nv, LK_STRIDE))
    I = np.zeros((nv, LI_STRIDE), np.int32)
    last_of_body = {}
    for j in range(nv):
        last_of_body[int(model.jnt_body[j])] = j
    first_seen = set()
    for j in range(nv):
        b = int(model.jnt_body[j])
        if b not in first_seen:                       # first joint of its body: hangs off the parent body's last link
            first_seen.add(b)
            pb = int(model.body_parent[b])
            if pb < 0:
                parent, base = -1, np.zeros(3)
            else:
                parent, base = last_of_body[pb], model.jnt_pos[last_of_body[pb]]
            Rfix = model.body_mat[b]
            off = model.body_pos[b] - base + Rfix @ model.jnt_pos[j]
        else:                                          # next joint of the same body
            parent, Rfix, off = j - 1, np.eye(3), model.jnt_pos[j] - model.jnt_pos[j - 1]
        P[j, LK_RFIX:LK_RFIX + 9] = Rfix.reshape(9)
        P[j, LK_OFF:LK_OFF + 3] = off
        P[j, LK_AXIS:LK_AXIS + 3] = model.jnt_axis[j]
        I[j, LI_PARENT], I[j, LI_TYPE], I[j, LI_LIMITED], I[j, LI_ACT] = parent, model.jnt_type[j], model.jnt_limited[j], -1
        if np.array_equal(Rfix, np.eye(3)):
            I[j, LI_BODY] |= 2                          # flag: the kernel skips the multiplication by Rfix
        if last_of_body[b] == j and model.body_mass[b] > 0:
            I[j, LI_BODY] |= 1
            P[j, LK_MASS] = model.body_mass[b]
            P[j, LK_COM:LK_COM + 3] = model.body_ipos[b] - model.jnt_pos[j]
            Rin = model.body_imat[b]
            Ic = Rin @ np.diag(model.body_inertia[b]) @ Rin.T
            P[j, LK_IC:LK_IC + 6] = [Ic[0, 0], Ic[1, 1], Ic[2, 2], Ic[0, 1], Ic[0, 2], Ic[1, 2]]
            P[j, LK_RIN:LK_RIN + 9] = Rin.T.reshape(9)           # rows = inertial axes in link coordinates
            In = model.body_inertia[b]
            P[j, LK_BOX:LK_BOX + 3] = [math.sqrt(max(1e-15, In[1] + In[2] - In[0]) / model.body_mass[b] * 6.0),
                                       math.sqrt(max(1e-15, In[0] + In[2] - In[1]) / model.body_mass[b] * 6.0),
                                       math.sqrt(max(1e-15, In[0] + In[1] - In[2]) / model.body_mass[b] * 6.0)]
        K, B = solref_to_kb(model.jnt_solref[j], model.jnt_solimp[j], model.timestep)
        P[j, LK_ARM], P[j, LK_DAMP], P[j, LK_STIFF], P[j, LK_SREF] = (model.jnt_armature[j], model.jnt_damping[j],
                                                                      model.jnt_stiffness[j], model.jnt_springref[j])
        P[j, LK_LO], P[j, LK_HI], P[j, LK_INVW], P[j, LK_SOLK], P[j, LK_SOLB] = (model.jnt_range[j, 0], model.jnt_range[j, 1],
                                                                               model.dof_invweight0[j], K, B)
        P[j, LK_SOLIMP:LK_SOLIMP + 5] = model.jnt_solimp[j]
    for a in range(model.nu):
        j = int(model.act_dof[a])
        if I[j, LI_ACT] >= 0:
            raise UnsupportedMjcf("two motors on one joint")
        I[j, LI_ACT] = a
        P[j, LK_GEAR] = model.act_gear[a]
        lo, hi = (model.act_ctrlrange[a] if model.act_ctrllimited[a] else (-np.inf, np.inf))
        P[j, LK_CLO], P[j, LK_CHI] = lo, hi
    G = np.zeros(G_STRIDE)
    G[G_DT], G[G_GRAV:G_GRAV + 3], G[G_RHO], G[G_VISC] = model.timestep, model.gravity, model.density, model.viscosity
    return P, I, G


# ------------------------------------------------------------------------------------------------ domain randomisation
def randomized_copy(model: TreeModel, param_dict: dict, rng: np.random.RandomState, defaults: dict = None):
    """(model', defaults, randomized): the named fields perturbed as the reference does (gym_env_wrapper.py:367-416):
    ``biased = (1 + bias) * default``, ``value ~ U(biased * (1 - noise), biased * (1 + noise))``, written into the model
    arrays (qpos0 constants such as dof_invweight0 are NOT recomputed -- neither does mujoco_py).  Fields of this model
    family: body_mass, body_inertia (the principal moments), dof_damping; dof_frictionloss / geom_* / sensor_noise do not
    exist in the subset: a non-zero request raises, a zero one still consumes the reference's random draws."""
    import copy
    m = copy.deepcopy(model)
    defaults = {} if defaults is None else defaults
    randomized = {}
    for param_id, entries in param_dict.items():
        defaults.setdefault(param_id, {})
        randomized.setdefault(param_id, {})
        for name, (noise_scale, bias_scale) in entries.items():
            if param_id == "body_mass":
                idx, field_ = m.body_names.index(name), m.body_mass
            elif param_id == "body_inertia":
                idx, field_ = m.body_names.index(name), m.body_inertia
            elif param_id == "dof_damping":
                idx, field_ = m.jnt_names.index(name), m.jnt_damping
            elif param_id in ("dof_frictionloss", "geom_size", "geom_friction", "sensor_noise"):
                if noise_scale != 0.0 or bias_scale != 0.0:
                    raise ValueError("dynamics field %s is not modelled by the GPU rollout" % param_id)
                rng.uniform(size=3 if param_id in ("geom_size", "geom_friction") else None)
                continue
            else:
                raise ValueError("Unknown dynamics field")
            if name not in defaults[param_id]:
                defaults[param_id][name] = copy.deepcopy(field_[idx])
            cur = copy.deepcopy(defaults[param_id][name])
            biased = (1.0 + bias_scale) * cur
            val = rng.uniform(biased - biased * noise_scale, biased + biased * noise_scale)
            field_[idx] = val
            randomized[param_id][name] = val
    return m, defaults, randomized


# ------------------------------------------------------------------------------------------------ planar mechanisms
# A tree whose hinge axes are all parallel (to n) and whose slides are all perpendicular to n moves in a plane: the
# reference's swimmer (n = z) and half-cheetah (n = y) both do.  The kernel has a planar instantiation with 3-vectors
# (omega; vx, vy) in place of spatial 6-vectors; the reduction is exact for any inertia tensor (only n'I n enters the
# in-plane equations) and for mj_passive's inertia-box fluid model (the projections of the inertial axes on the plane
# turn with the body, their components along n are constant).  Per link, doubles (frames are WORLD-ALIGNED at q = 0):
PK_OFF, PK_DIR, PK_MASS, PK_COM, PK_INN, PK_CLIN, PK_KV1, PK_KV2, PK_E, PK_AK, PK_STRIDE = 0, 2, 4, 5, 7, 8, 9, 10, 11, 17, 20
# PK_DIR: hinge: (sign of the axis along n, 0); slide: unit direction in the plane.  PK_E: the three inertial axes
# projected on the plane (x0 y0 x1 y1 x2 y2); PK_AK: 1/2 rho * face area seen along each inertial axis.
# per link, ints: bit mask of the link's ancestors
PKG_GRAV, PKG_STRIDE = 0, 2           # planar globals: gravity projected on the plane


def pack_planar(model: TreeModel, tol: float = 1e-12):
    """Planar parameter block (P, anc, G) of a planar mechanism, or None."""
    nv = model.nv
    hinges = np.nonzero(model.jnt_type == HINGE)[0]
    if hinges.size == 0:
        return None
    xpos, xmat, anchor, axis = kinematics(model, np.zeros(nv))
    n = axis[hinges[0]]
    sign = np.zeros(nv)
    for j in range(nv):
        c = float(axis[j] @ n)
        if model.jnt_type[j] == HINGE:
            if abs(abs(c) - 1.0) > tol:
                return None
            sign[j] = 1.0 if c > 0 else -1.0
        elif abs(c) > tol:
            return None
    k = int(np.argmin(np.abs(n)))
    ex = np.eye(3)[k] - n[k] * n
    ex /= np.linalg.norm(ex)
    ey = np.cross(n, ex)
    plane = lambda w: np.array([ex @ w, ey @ w])
    _, I, _ = pack_links(model)
    P = np.zeros((nv, PK_STRIDE))
    anc = np.zeros(nv, np.int32)
    last_of_body = {int(model.jnt_body[j]): j for j in range(nv)}
    for j in range(nv):
        p = int(I[j, LI_PARENT])
        anc[j] = 0 if p < 0 else (anc[p] | (1 << p))
        P[j, PK_OFF:PK_OFF + 2] = plane(anchor[j] - (anchor[p] if p >= 0 else 0.0))
        P[j, PK_DIR:PK_DIR + 2] = (sign[j], 0.0) if model.jnt_type[j] == HINGE else plane(axis[j])
        b = int(model.jnt_body[j])
        if last_of_body[b] == j and model.body_mass[b] > 0:
            m = model.body_mass[b]
            Ri = xmat[b] @ model.body_imat[b]                    # columns: inertial axes in the world at q = 0
            In = model.body_inertia[b]
            P[j, PK_MASS] = m
            P[j, PK_COM:PK_COM + 2] = plane(xpos[b] + xmat[b] @ model.body_ipos[b] - anchor[j])
            P[j, PK_INN] = float(n @ Ri @ np.diag(In) @ Ri.T @ n)
            box = np.sqrt(np.maximum(1e-15, np.array([In[1] + In[2] - In[0], In[0] + In[2] - In[1], In[0] + In[1] - In[2]])) / m * 6.0)
            d = box.sum() / 3.0
            nk = Ri.T @ n
            quart = np.array([box[1] ** 4 + box[2] ** 4, box[0] ** 4 + box[2] ** 4, box[0] ** 4 + box[1] ** 4])
            P[j, PK_CLIN] = 3.0 * math.pi * d * model.viscosity
            P[j, PK_KV1] = math.pi * d ** 3 * model.viscosity * float(nk @ nk)
            P[j, PK_KV2] = model.density * float(np.sum(box * quart * np.abs(nk) ** 3)) / 64.0
            for a in range(3):
                P[j, PK_E + 2 * a:PK_E + 2 * a + 2] = plane(Ri[:, a])
            P[j, PK_AK:PK_AK + 3] = 0.5 * model.density * np.array([box[1] * box[2], box[0] * box[2], box[0] * box[1]])
    return P, anc, plane(model.gravity)


# contact candidates of a planar mechanism (csrc/rollout_tree_planar.cuh): per candidate, doubles / ints
CT_A, CT_HA, CT_RA, CT_B, CT_HB, CT_RB, CT_MU, CT_K, CT_BB, CT_SOLIMP, CT_INVW, CT_BOUND, CT_STRIDE, CTI_STRIDE, MAX_CAND = \
    0, 2, 4, 5, 7, 9, 10, 11, 12, 13, 18, 19, 20, 3, 16


def pack_planar_contacts(model: TreeModel, tol: float = 1e-9):
    """(ints (n, 3), doubles (n, CT_STRIDE)) for the planar kernel, or raises when the contacts do not live in the plane
    of motion (shapes at different heights along the plane normal, a floor that is not perpendicular to the plane)."""
    nv = model.nv
    if not model.contacts:
        return np.zeros((0, CTI_STRIDE), np.int32), np.zeros((0, CT_STRIDE))
    if pack_planar(model) is None:
        raise UnsupportedMjcf("contacts are only simulated for planar mechanisms")
    if len(model.contacts) > MAX_CAND:
        raise UnsupportedMjcf("%d contact candidates > %d" % (len(model.contacts), MAX_CAND))
    xpos, xmat, anchor, axis = kinematics(model, np.zeros(nv))
    hinges = np.nonzero(model.jnt_type == HINGE)[0]
    n = axis[hinges[0]]
    k = int(np.argmin(np.abs(n)))
    ex = np.eye(3)[k] - n[k] * n
    ex /= np.linalg.norm(ex)
    ey = np.cross(n, ex)
    plane = lambda w: np.array([ex @ w, ey @ w])
    last = {int(model.jnt_body[j]): j for j in range(nv)}
    I, D = np.zeros((len(model.contacts), CTI_STRIDE), np.int32), np.zeros((len(model.contacts), CT_STRIDE))
    heights = []
    for c, ct in enumerate(model.contacts):
        l2 = last[ct["body2"]]
        w0, w1 = (xpos[ct["body2"]] + xmat[ct["body2"]] @ ct[e] for e in ("b0", "b1"))
        heights += [n @ w0, n @ w1]
        D[c, CT_B:CT_B + 2], D[c, CT_HB:CT_HB + 2], D[c, CT_RB] = plane(0.5 * (w0 + w1) - anchor[l2]), plane(0.5 * (w1 - w0)), ct["rb"]
        if ct["kind"] == "plane":
            if abs(n @ ct["a1"]) > tol:
                raise UnsupportedMjcf("a contact plane must be perpendicular to the plane of motion")
            I[c] = (1, -1, l2)
            D[c, CT_A:CT_A + 2], D[c, CT_HA:CT_HA + 2] = plane(ct["a0"]), plane(ct["a1"])
        else:
            # MuJoCo orients the friction pyramid of a capsule-capsule contact by mju_makeFrame (first tangent from the world
            # y or z axis), and a pyramid is not invariant under a turn about the normal.  When the plane of motion is a
            # world coordinate plane the tangents come out in the plane / along its normal -- the three merged rows the kernel
            # builds; in a tilted plane they would be a mixture, which is not implemented.  (Plane-capsule contacts take their
            # first tangent from the capsule axis, which lies in the plane of motion whatever its orientation.)
            if np.sort(np.abs(n))[1] > tol:
                raise UnsupportedMjcf("capsule-capsule contacts need the plane of motion to be a world coordinate plane")
            l1 = last[ct["body1"]]
            v0, v1 = (xpos[ct["body1"]] + xmat[ct["body1"]] @ ct[e] for e in ("a0", "a1"))
            heights += [n @ v0, n @ v1]
            I[c] = (0, l1, l2)
            D[c, CT_A:CT_A + 2], D[c, CT_HA:CT_HA + 2], D[c, CT_RA] = plane(0.5 * (v0 + v1) - anchor[l1]), plane(0.5 * (v1 - v0)), ct["ra"]
            # centres farther apart than this cannot touch (a hair of slack for the rounding of the kernel's own centres)
            D[c, CT_BOUND] = (np.linalg.norm(D[c, CT_HA:CT_HA + 2]) + np.linalg.norm(D[c, CT_HB:CT_HB + 2]) + ct["ra"] + ct["rb"]) * (1 + 1e-9)
        K, B = solref_to_kb(ct["solref"], ct["solimp"], model.timestep)
        D[c, CT_MU], D[c, CT_K], D[c, CT_BB], D[c, CT_INVW] = ct["mu"], K, B, ct["invweight"]
        D[c, CT_SOLIMP:CT_SOLIMP + 5] = ct["solimp"]
    if max(heights) - min(heights) > tol:
        raise UnsupportedMjcf("colliding shapes at different heights along the plane normal")
    return I, D


# ------------------------------------------------------------------------------------------------ shipped models
def swimmer_mjcf(radii=(0.07, 0.065, 0.06, 0.055, 0.05), half_length=0.15, spacing=0.3, joint_range=1.5, gear=20.0,
                 timestep=0.005, viscosity=0.000894, density=1000.0, height=0.03) -> str:
    """MJCF text of the reference's ``Swimmer-v0`` model (``mjmpc/envs/assets/xml/swimmer.xml``), physics only: a
    torso on two slides and a hinge (planar base) followed by ``len(radii) - 1`` capsule links on limited z hinges
    with geared motors, in a fluid.  ``/root/reference`` does not travel to the GPU box, so the model is restated
    as this parameter list; tests/test_tree_cpu.py compares its compilation with the reference file's."""
    n = len(radii)
    geom = '<geom type="capsule" pos="%g 0 0" quat="0.707 0 -0.707 0" size="%g %g"/>'
    parts = ['<mujoco model="swimmer"><compiler inertiafromgeom="true" angle="radian"/>',
             '<default><joint limited="true" range="%g %g"/><motor ctrllimited="true" ctrlrange="-1 1"/></default>'
             % (-joint_range, joint_range),
             '<option timestep="%g" viscosity="%g" density="%g"/>' % (timestep, viscosity, density),
             '<worldbody><body name="torso" pos="0 0 %g">' % height,
             '<joint type="slide" axis="1 0 0" limited="false"/><joint type="slide" axis="0 1 0" limited="false"/>',
             '<joint type="hinge" axis="0 0 1" limited="false"/>', geom % (half_length, radii[0], half_length)]
    for i in range(1, n):
        parts.append('<body name="link%d" pos="%g 0 0"><joint name="j%d" type="hinge" axis="0 0 1"/>' % (i, spacing, i))
        parts.append(geom % (half_length, radii[i], half_length))
    parts.append("</body>" * n + "</worldbody><actuator>")
    parts += ['<motor joint="j%d" gear="%g"/>' % (i, gear) for i in range(1, n)]
    parts.append("</actuator></mujoco>")
    return "".join(parts)


def half_cheetah_mjcf() -> str:
    """MJCF text of the reference's ``HalfCheetah-v0`` model (``mjmpc/envs/assets/xml/half_cheetah.xml``), physics only:
    a torso (two capsules) on slide x / slide z / hinge y, a back and a front leg of three spring-loaded, damped, limited
    hinges each, geared motors, total mass set to 14, a floor the capsules collide with (friction 0.4, condim 3).  Restated
    as a parameter list because ``/root/reference`` does not travel to the GPU box; tests/test_tree_cpu.py compares its
    compilation with the reference file's."""
    legs = [  # body, pos, joint range, stiffness, damping, geom axisangle-y, geom pos, geom half length, gear
        [("bthigh", "-.5 0 0", "-.52 1.05", 240, 6, -3.8, ".1 0 -.13", .145, 120),
         ("bshin", ".16 0 -.25", "-.785 .785", 180, 4.5, -2.03, "-.14 0 -.07", .15, 90),
         ("bfoot", "-.28 0 -.14", "-.4 .785", 120, 3, -.27, ".03 0 -.097", .094, 60)],
        [("fthigh", ".5 0 0", "-1 .7", 180, 4.5, .52, "-.07 0 -.12", .133, 120),
         ("fshin", "-.14 0 -.24", "-1.2 .87", 120, 3, -.6, ".065 0 -.09", .106, 60),
         ("ffoot", ".13 0 -.18", "-.5 .5", 60, 1.5, -.6, ".045 0 -.07", .07, 30)]]
    parts = ['<mujoco model="cheetah"><compiler angle="radian" coordinate="local" inertiafromgeom="true" settotalmass="14"/>',
             '<default><joint armature=".1" damping=".01" limited="true" solimplimit="0 .8 .03" solreflimit=".02 1" stiffness="8"/>',
             '<geom conaffinity="0" condim="3" contype="1" friction=".4 .1 .1" solimp="0.0 0.8 0.01" solref="0.02 1"/>',
             '<motor ctrllimited="true" ctrlrange="-1 1"/></default><option gravity="0 0 -9.81" timestep="0.01"/><worldbody>',
             '<geom conaffinity="1" condim="3" name="floor" pos="0 0 0" size="40 40 40" type="plane"/><body name="torso" pos="0 0 .7">']
    for name, axis, kind in (("rootx", "1 0 0", "slide"), ("rootz", "0 0 1", "slide"), ("rooty", "0 1 0", "hinge")):
        parts.append('<joint armature="0" axis="%s" damping="0" limited="false" name="%s" pos="0 0 0" stiffness="0" type="%s"/>'
                     % (axis, name, kind))
    parts.append('<geom fromto="-.5 0 0 .5 0 0" name="torso" size="0.046" type="capsule"/>')
    parts.append('<geom axisangle="0 1 0 .87" name="head" pos=".6 0 .1" size="0.046 .15" type="capsule"/>')
    for leg in legs:
        for name, pos, rng, k, d, ang, gpos, half, _ in leg:
            parts.append('<body name="%s" pos="%s"><joint axis="0 1 0" damping="%g" name="%s" pos="0 0 0" range="%s" stiffness="%g" '
                         'type="hinge"/><geom axisangle="0 1 0 %g" name="%s" pos="%s" size="0.046 %g" type="capsule"/>'
                         % (name, pos, d, name, rng, k, ang, name, gpos, half))
        parts.append("</body>" * 3)
    parts.append("</body></worldbody><actuator>")
    parts += ['<motor gear="%g" joint="%s" name="%s"/>' % (g, n, n) for leg in legs for (n, _, _, _, _, _, _, _, g) in leg]
    parts.append("</actuator></mujoco>")
    return "".join(parts)
