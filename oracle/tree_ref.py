"""oracle/tree_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A SECOND, independent restatement of one MuJoCo 2.0 ``mj_step`` for hinge / slide trees, written to pin
``oracle/tree_step.c`` (and ``csrc/rollout_tree.cu``) against something that shares no code, no compiled constant and
no derivation with either (the counterpart of ``oracle/efc_ref.py`` for SURVEY §8 f-3):

  * reads the MJCF itself (xml.etree; quaternion algebra, not rotation matrices) and applies MuJoCo's compiler rules
    (``inertiafromgeom``, density 1000, capsule = cylinder + two half spheres, parallel axis) -- it does NOT import
    ``mjmpc_b200.envs.mjcf_tree``;
  * kinematics as one differentiable function q -> (centre of mass, orientation) of every body; the translational
    Jacobians are its COMPLEX-STEP derivatives, the rotational ones are the joint axes;
  * M(q) = sum_b m Jv'Jv + Jw' I_world Jw (+ armature): no recursion, no spatial algebra;
  * Coriolis / centrifugal bias from the Christoffel symbols with dM/dq by complex step; gravity from the derivative
    of the potential energy;
  * fluid forces (mj_passive's inertia-box model) from the body's centre-of-mass velocity Jv qdot, Jw qdot, mapped
    back through the same Jacobians;
  * dof_invweight0 = diag(M(0)^-1); limit rows per ``mj_instantiateLimit`` / ``mj_makeImpedance`` /
    ``mj_referenceConstraint``; the convex problem solved by ENUMERATING active sets and keeping the KKT-consistent one;
  * contacts (read_model(..., contacts=True)): its own candidate pairs and body_invweight0 (complex-step Jacobians), a TRUE
    closest-point search between segments (not MuJoCo's sequential clamping: the same points in general position),
    contact Jacobians by complex step of the material point, pyramidal rows, active-set enumeration over dense rows;
  * mj_Euler with implicit joint damping.

Reference call sites: mjmpc/envs/basic/swimmer.py:7-24 (frame_skip 4, reward, observation),
mjmpc/envs/assets/xml/swimmer.xml (the model).  MuJoCo itself is absent from the reference tree and this image: this
file narrows "parity unpinned" to "two independent restatements of MuJoCo's documented algorithm agree"; the committed
vectors ``tests/golden/tree_pin.npz`` are ITS outputs (tests/golden/gen_tree_pin.py).
"""
from __future__ import annotations

import itertools
import math
import xml.etree.ElementTree as ET

import numpy as np

MINVAL = 1e-15


def _v(s):
    return np.array([float(x) for x in s.split()])


def _clamp_solimp(v):
    """MuJoCo clamps the impedances and the midpoint to [1e-4, 1 - 1e-4], the width to >= 0, the power to >= 1."""
    v = np.array(v, float)
    v[[0, 1, 3]] = np.clip(v[[0, 1, 3]], 1e-4, 1 - 1e-4)
    v[2], v[4] = max(0.0, v[2]), max(1.0, v[4])
    return v


# ------------------------------------------------------------------------------------------------ quaternions
def qmul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                     a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
                     a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def qrot(q, v):
    """Rotate v by unit quaternion q (complex-step safe: no conjugation of the imaginary perturbation)."""
    w, u = q[0], q[1:]
    t = 2.0 * np.cross(u, v)
    return v + w * t + np.cross(u, t)


def qaxis(axis, angle):
    return np.concatenate([[np.cos(0.5 * angle)], np.sin(0.5 * angle) * axis])


def qfrom_z(v):
    v = v / np.linalg.norm(v)
    c = v[2]
    ax = np.array([-v[1], v[0], 0.0])
    s = np.linalg.norm(ax)
    if s < 1e-12:
        return np.array([1.0, 0, 0, 0]) if c > 0 else np.array([0.0, 1, 0, 0])
    return qaxis(ax / s, math.atan2(s, c))


def qmat(q):
    return np.stack([qrot(q, e) for e in np.eye(3)], axis=1)


# ------------------------------------------------------------------------------------------------ model
def read_model(xml_path: str, contacts: bool = False, text: str = None) -> dict:
    root = ET.fromstring(text) if text is not None else ET.parse(xml_path).getroot()
    comp = root.find("compiler").attrib if root.find("compiler") is not None else {}
    deg = comp.get("angle", "degree") != "radian"
    opt = root.find("option").attrib if root.find("option") is not None else {}
    classes = {}

    def read_defaults(node, inherited, name):
        table = {k: dict(v) for k, v in inherited.items()}
        for e in node:
            if e.tag != "default":
                table.setdefault(e.tag, {}).update(e.attrib)
        classes[name] = table
        for sub in node.findall("default"):
            read_defaults(sub, table, sub.get("class"))

    for top in root.findall("default"):
        read_defaults(top, {}, "main")
    classes.setdefault("main", {})

    def attrs(e, cc):
        a = dict(classes[e.get("class", cc or "main")].get(e.tag, {}))
        a.update({k: v for k, v in e.attrib.items() if k != "class"})
        return a

    def orient(a):
        if "quat" in a:
            q = _v(a["quat"])
            return q / np.linalg.norm(q)
        if "axisangle" in a:
            t = _v(a["axisangle"])
            return qaxis(t[:3] / np.linalg.norm(t[:3]), t[3] * (math.pi / 180 if deg else 1.0))
        if "euler" in a:                 # default eulerseq "xyz": rotations about the body's own x, then y, then z
            e = _v(a["euler"]) * (math.pi / 180 if deg else 1.0)
            q = np.array([1.0, 0, 0, 0])
            for ax, ang in zip(np.eye(3), e):
                q = qmul(q, qaxis(ax, ang))
            return q
        return np.array([1.0, 0, 0, 0])

    bodies = []

    def walk(e, parent, cc):
        for be in e.findall("body"):
            c2 = be.get("childclass", cc)
            geoms = []
            for ge in be.findall("geom"):
                a = attrs(ge, c2)
                kind, size = a.get("type", "sphere"), _v(a.get("size", "0"))
                rho = float(a.get("density", "1000"))
                if "fromto" in a:
                    ft = _v(a["fromto"])
                    pos, quat, half = 0.5 * (ft[:3] + ft[3:]), qfrom_z(ft[3:] - ft[:3]), 0.5 * np.linalg.norm(ft[3:] - ft[:3])
                else:
                    pos, quat, half = _v(a.get("pos", "0 0 0")), orient(a), (size[1] if size.size > 1 else 0.0)
                r = size[0]
                if kind == "sphere":
                    m = rho * 4 / 3 * math.pi * r ** 3
                    I = np.full(3, 0.4 * m * r * r)
                elif kind == "capsule":
                    hh = 2 * half
                    mc, ms = rho * math.pi * r * r * hh, rho * 4 / 3 * math.pi * r ** 3
                    m = mc + ms
                    I = np.array([mc * (3 * r * r + hh * hh) / 12 + ms * (0.4 * r * r + 0.25 * hh * hh + 0.375 * r * hh)] * 2
                                 + [0.5 * mc * r * r + 0.4 * ms * r * r])
                else:
                    raise ValueError(kind)
                geoms.append((m, pos, quat, I))
                coll = dict(r=r, half=half, contype=int(a.get("contype", "1")), conaffinity=int(a.get("conaffinity", "1")),
                            mu=_v(a.get("friction", "1 0.005 0.0001"))[0], solref=_v(a.get("solref", "0.02 1")),
                            solimp=_clamp_solimp(np.concatenate([_v(a.get("solimp", "0.9 0.95 0.001")), [0.5, 2.0]])[:5]
                                                 if len(_v(a.get("solimp", "0.9 0.95 0.001"))) < 5 else _v(a["solimp"])))
            joints = []
            for je in be.findall("joint"):
                a = attrs(je, c2)
                ax = _v(a.get("axis", "0 0 1"))
                joints.append(dict(type=a.get("type", "hinge"), pos=_v(a.get("pos", "0 0 0")), axis=ax / np.linalg.norm(ax),
                                   limited=a.get("limited", "false") == "true", range=_v(a.get("range", "0 0")),
                                   damping=float(a.get("damping", "0")), armature=float(a.get("armature", "0")),
                                   stiffness=float(a.get("stiffness", "0")), springref=float(a.get("springref", "0")),
                                   solref=_v(a.get("solreflimit", "0.02 1")),
                                   solimp=_clamp_solimp(np.concatenate([_v(a["solimplimit"]), [0.9, 0.95, 0.001, 0.5, 2.0][len(_v(a["solimplimit"])):]])
                                                        if "solimplimit" in a else np.array([0.9, 0.95, 0.001, 0.5, 2.0])),
                                   name=a.get("name", "")))
            if len(geoms) != 1:
                raise ValueError("tree_ref reads one geom per body")       # principal-axis composition not needed here
            m, c, gq, I = geoms[0]
            bodies.append(dict(parent=parent, pos=_v(be.get("pos", "0 0 0")), quat=orient(be.attrib), joints=joints,
                               mass=m, ipos=c, iquat=gq, inertia=I, coll=coll))
            walk(be, len(bodies) - 1, c2)

    walk(root.find("worldbody"), -1, None)
    dofs = [(b, j) for b, body in enumerate(bodies) for j in body["joints"]]
    names = [j["name"] for _, j in dofs]
    motors = []
    for me in root.findall("actuator/motor"):
        a = attrs(me, None)
        motors.append(dict(dof=names.index(a["joint"]), gear=_v(a.get("gear", "1"))[0],
                           range=_v(a["ctrlrange"]) if a.get("ctrllimited", "false") == "true" else np.array([-np.inf, np.inf])))
    model = dict(bodies=bodies, dofs=dofs, motors=motors, nv=len(dofs), h=float(opt.get("timestep", "0.002")),
                 gravity=_v(opt.get("gravity", "0 0 -9.81")), rho=float(opt.get("density", "0")),
                 mu=float(opt.get("viscosity", "0")))
    Minv0 = np.linalg.inv(mass_matrix(model, np.zeros(model["nv"])))
    model["invweight0"] = np.diag(Minv0).copy()
    # contacts (read only when asked: read_model(..., contacts=True)): world planes and the bodies' capsules / spheres
    model["planes"], model["pairs"] = [], []
    if contacts:
        for ge in root.find("worldbody").findall("geom"):
            a = attrs(ge, None)
            if a.get("type", "sphere") == "plane" and (int(a.get("contype", "1")) or int(a.get("conaffinity", "1"))):
                model["planes"].append(dict(point=_v(a.get("pos", "0 0 0")), normal=qrot(orient(a), np.array([0.0, 0, 1])),
                                            contype=int(a.get("contype", "1")), conaffinity=int(a.get("conaffinity", "1")),
                                            mu=_v(a.get("friction", "1 0.005 0.0001"))[0]))
        Jv0, _, _ = jacobians(model, np.zeros(model["nv"]))
        bw = [np.trace(Jv0[b] @ Minv0 @ Jv0[b].T) / 3.0 for b in range(len(bodies))]       # body_invweight0, translational
        can = lambda x, y: bool((x["contype"] & y["conaffinity"]) or (y["contype"] & x["conaffinity"]))
        for b, body in enumerate(bodies):
            for pl in model["planes"]:
                if can(pl, body["coll"]):
                    model["pairs"].append(dict(plane=pl, b2=b, mu=max(pl["mu"], body["coll"]["mu"]), invw=bw[b]))
            for b1 in range(b):
                o = bodies[b1]
                if body["parent"] == b1 or o["parent"] == b or not can(o["coll"], body["coll"]):
                    continue
                model["pairs"].append(dict(plane=None, b1=b1, b2=b, mu=max(o["coll"]["mu"], body["coll"]["mu"]), invw=bw[b1] + bw[b]))
    return model


# ------------------------------------------------------------------------------------------------ kinematics
def forward(model, q):
    """(com positions, inertial-frame quaternions, joint axes in the world, joint anchors) -- works for complex q."""
    pos, quat = [], []
    axes, anchors = [None] * model["nv"], [None] * model["nv"]
    k = 0
    for body in model["bodies"]:
        p = body["parent"]
        P, Q = (np.zeros(3), np.array([1.0, 0, 0, 0])) if p < 0 else (pos[p], quat[p])
        P = P + qrot(Q, body["pos"])
        Q = qmul(Q, body["quat"])
        for j in body["joints"]:
            anchors[k], axes[k] = P + qrot(Q, j["pos"]), qrot(Q, j["axis"])
            if j["type"] == "slide":
                P = P + axes[k] * q[k]
            else:
                Q = qmul(Q, qaxis(j["axis"], q[k]))
                P = anchors[k] - qrot(Q, j["pos"])
            k += 1
        pos.append(P)
        quat.append(Q)
    com = [pos[b] + qrot(quat[b], body["ipos"]) for b, body in enumerate(model["bodies"])]
    iq = [qmul(quat[b], body["iquat"]) for b, body in enumerate(model["bodies"])]
    return com, iq, axes, anchors


def jacobians(model, q):
    """Per body: Jv (3, nv) = d com / dq by complex step, Jw (3, nv) = axes of the hinges above the body."""
    nv, nb = model["nv"], len(model["bodies"])
    Jv = np.zeros((nb, 3, nv))
    for i in range(nv):
        qc = np.array(q, complex)
        qc[i] += 1e-30j
        com, _, _, _ = forward(model, qc)
        for b in range(nb):
            Jv[b, :, i] = np.imag(com[b]) / 1e-30
    _, iq, axes, _ = forward(model, np.asarray(q, float))
    Jw = np.zeros((nb, 3, nv))
    for b in range(nb):
        a = b
        chain = []
        while a >= 0:
            chain.append(a)
            a = model["bodies"][a]["parent"]
        for i, (bi, j) in enumerate(model["dofs"]):
            if bi in chain and j["type"] == "hinge":
                Jw[b, :, i] = np.real(axes[i])
    return Jv, Jw, iq


def mass_matrix(model, q):
    """Complex-step safe in q (used for dM/dq): Jv by finite complex step is not available inside a complex
    evaluation, so M is assembled from analytic cross products here and from the complex-step Jacobians in
    ``mass_matrix_cs`` -- the two must agree (tests)."""
    com, iq, axes, anchors = forward(model, q)
    nv = model["nv"]
    M = np.zeros((nv, nv), dtype=np.result_type(np.asarray(q).dtype, float))
    for b, body in enumerate(model["bodies"]):
        chain, a = [], b
        while a >= 0:
            chain.append(a)
            a = model["bodies"][a]["parent"]
        Jv = np.zeros((3, nv), M.dtype)
        Jw = np.zeros((3, nv), M.dtype)
        for i, (bi, j) in enumerate(model["dofs"]):
            if bi not in chain:
                continue
            if j["type"] == "slide":
                Jv[:, i] = axes[i]
            else:
                Jv[:, i], Jw[:, i] = np.cross(axes[i], com[b] - anchors[i]), axes[i]
        R = qmat(iq[b])
        Iw = R @ np.diag(body["inertia"]) @ R.T
        M = M + body["mass"] * Jv.T @ Jv + Jw.T @ Iw @ Jw
    return M + np.diag([j["armature"] for _, j in model["dofs"]])


def mass_matrix_cs(model, q):
    Jv, Jw, iq = jacobians(model, q)
    nv = model["nv"]
    M = np.diag([float(j["armature"]) for _, j in model["dofs"]])
    for b, body in enumerate(model["bodies"]):
        R = qmat(iq[b])
        M = M + body["mass"] * Jv[b].T @ Jv[b] + Jw[b].T @ (R @ np.diag(body["inertia"]) @ R.T) @ Jw[b]
    return M


def bias_forces(model, q, v):
    """c(q, v) - gravity force:  c_i = sum_jk (dM_ij/dq_k - 1/2 dM_jk/dq_i) v_j v_k;  gravity = -dV/dq."""
    nv = model["nv"]
    dM = np.zeros((nv, nv, nv))
    dV = np.zeros(nv)
    for k in range(nv):
        qc = np.array(q, complex)
        qc[k] += 1e-30j
        dM[:, :, k] = np.imag(mass_matrix(model, qc)) / 1e-30
        com, _, _, _ = forward(model, qc)
        V = sum(-body["mass"] * (model["gravity"] @ com[b]) for b, body in enumerate(model["bodies"]))
        dV[k] = np.imag(V) / 1e-30
    c = np.einsum("ijk,j,k->i", dM, v, v) - 0.5 * np.einsum("jki,j,k->i", dM, v, v)
    return c + dV


def fluid_forces(model, q, v):
    out = np.zeros(model["nv"])
    if model["rho"] <= 0 and model["mu"] <= 0:
        return out
    Jv, Jw, iq = jacobians(model, q)
    for b, body in enumerate(model["bodies"]):
        m, I = body["mass"], body["inertia"]
        if m < MINVAL:
            continue
        box = np.sqrt(np.maximum(MINVAL, np.array([I[1] + I[2] - I[0], I[0] + I[2] - I[1], I[0] + I[1] - I[2]])) / m * 6.0)
        R = qmat(iq[b])
        lw, lv = R.T @ (Jw[b] @ v), R.T @ (Jv[b] @ v)
        tq, fr = np.zeros(3), np.zeros(3)
        if model["mu"] > 0:
            d = box.sum() / 3.0
            tq += -math.pi * d ** 3 * model["mu"] * lw
            fr += -3.0 * math.pi * d * model["mu"] * lv
        if model["rho"] > 0:
            area = np.array([box[1] * box[2], box[0] * box[2], box[0] * box[1]])
            fr -= 0.5 * model["rho"] * area * np.abs(lv) * lv
            quart = np.array([box[1] ** 4 + box[2] ** 4, box[0] ** 4 + box[2] ** 4, box[0] ** 4 + box[1] ** 4])
            tq -= model["rho"] * box * quart * np.abs(lw) * lw / 64.0
        out += Jv[b].T @ (R @ fr) + Jw[b].T @ (R @ tq)
    return out


# ------------------------------------------------------------------------------------------------ constraints
def impedance(solimp, dist):
    d0, d1, width, mid, power = solimp
    x = abs(dist) / width
    if x >= 1:
        return d1
    if x <= 0:
        return d0
    if power == 1:
        y = x
    elif x <= mid:
        y = x ** power / mid ** (power - 1)
    else:
        y = 1 - (1 - x) ** power / (1 - mid) ** (power - 1)
    return d0 + y * (d1 - d0)


def limit_rows(model, q, v):
    rows = []
    for i, (_, j) in enumerate(model["dofs"]):
        if not j["limited"]:
            continue
        tc, dr, dmax = max(j["solref"][0], 2 * model["h"]), j["solref"][1], j["solimp"][1]
        Kk, Bb = 1.0 / (dmax * dmax * tc * tc * dr * dr), 2.0 / (dmax * tc)
        for sign, dist in ((1.0, q[i] - j["range"][0]), (-1.0, j["range"][1] - q[i])):
            if dist < 0:
                imp = impedance(j["solimp"], dist)
                R = max(MINVAL, (1 - imp) / imp * model["invweight0"][i])
                rows.append(dict(dof=i, sign=sign, D=1.0 / R, aref=-Bb * sign * v[i] - Kk * imp * dist, dist=dist))
    return rows


def solve_rows(M, f, rows):
    """Enumerate active sets; the minimiser is the one whose residual signs agree with its own set."""
    n = len(rows)
    for active in itertools.product((False, True), repeat=n):
        H, rhs = M.copy(), f.copy()
        for r, on in zip(rows, active):
            if on:
                H[r["dof"], r["dof"]] += r["D"]
                rhs[r["dof"]] += r["sign"] * r["D"] * r["aref"]
        a = np.linalg.solve(H, rhs)
        res = [r["sign"] * a[r["dof"]] - r["aref"] for r in rows]
        if all((x < 0) == on for x, on in zip(res, active)):
            fc = np.zeros_like(f)
            for r, x, on in zip(rows, res, active):
                if on:
                    fc[r["dof"]] += r["sign"] * (-r["D"] * x)
            return a, fc
    raise RuntimeError("no consistent active set")


def _material_point_jacobian(model, q, body, world_point):
    """d/dq of the world position of the point of ``body`` that currently sits at ``world_point`` (complex step)."""
    def frames(qq):
        pos, quat = [], []
        k = 0
        for bd in model["bodies"]:
            p = bd["parent"]
            P, Q = (np.zeros(3), np.array([1.0, 0, 0, 0])) if p < 0 else (pos[p], quat[p])
            P = P + qrot(Q, bd["pos"])
            Q = qmul(Q, bd["quat"])
            for j in bd["joints"]:
                anchor, axis = P + qrot(Q, j["pos"]), qrot(Q, j["axis"])
                if j["type"] == "slide":
                    P = P + axis * qq[k]
                else:
                    Q = qmul(Q, qaxis(j["axis"], qq[k]))
                    P = anchor - qrot(Q, j["pos"])
                k += 1
            pos.append(P)
            quat.append(Q)
        return pos, quat
    pos, quat = frames(np.asarray(q, float))
    qc = np.array([quat[body][0], *(-quat[body][1:])])
    local = qrot(qc, world_point - pos[body])
    J = np.zeros((3, model["nv"]))
    for i in range(model["nv"]):
        qq = np.array(q, complex)
        qq[i] += 1e-30j
        P, Q = frames(qq)
        J[:, i] = np.imag(P[body] + qrot(Q[body], local)) / 1e-30
    return J


def _closest_points(c1, a1, c2, a2):
    """Closest points of segments c1 + x1 a1, c2 + x2 a2, x in [-1, 1]: interior stationary point, else the best of the
    four edges (each a point-to-segment projection) -- not MuJoCo's sequential clamping, the same points in general position."""
    best = None
    def consider(x1, x2):
        nonlocal best
        d = np.linalg.norm(c1 + x1 * a1 - c2 - x2 * a2)
        if best is None or d < best[0]:
            best = (d, x1, x2)
    A = np.array([[a1 @ a1, -a1 @ a2], [-a1 @ a2, a2 @ a2]])
    rhs = np.array([-(c1 - c2) @ a1, (c1 - c2) @ a2])
    if abs(np.linalg.det(A)) > 1e-14:
        x = np.linalg.solve(A, rhs)
        if np.all(np.abs(x) <= 1):
            consider(x[0], x[1])
    for x1 in (-1.0, 1.0):
        p = c1 + x1 * a1
        x2 = np.clip((p - c2) @ a2 / max(a2 @ a2, 1e-300), -1, 1) if a2 @ a2 > 0 else 0.0
        consider(x1, x2)
    for x2 in (-1.0, 1.0):
        p = c2 + x2 * a2
        x1 = np.clip((p - c1) @ a1 / max(a1 @ a1, 1e-300), -1, 1) if a1 @ a1 > 0 else 0.0
        consider(x1, x2)
    return best[1], best[2]


def contact_rows(model, q, v):
    """Pyramidal rows of every active contact: 4 per contact (n +- mu t1, n +- mu t2), R = 2 mu^2 (1 - imp)/imp (1 + mu^2) w."""
    if not model["pairs"]:
        return []
    com, iq, _, _ = forward(model, np.asarray(q, float))
    rows = []
    for pr in model["pairs"]:
        b2 = model["bodies"][pr["b2"]]
        c2, z2 = com[pr["b2"]], qrot(iq[pr["b2"]], np.array([0.0, 0, 1]))          # one geom per body: geom frame = inertial frame
        r2, a2 = b2["coll"]["r"], z2 * b2["coll"]["half"]
        found = []
        if pr["plane"] is not None:
            n = pr["plane"]["normal"]
            for end in ((1.0, -1.0) if b2["coll"]["half"] > 0 else (0.0,)):
                p = c2 + end * a2
                dist = (p - pr["plane"]["point"]) @ n - r2
                if dist < 0:
                    found.append((dist, p - n * (r2 + 0.5 * dist), n, a2 if b2["coll"]["half"] > 0 else None))
            solref, solimp = b2["coll"]["solref"], b2["coll"]["solimp"]
            bodies = (None, pr["b2"])
        else:
            b1 = model["bodies"][pr["b1"]]
            c1, z1 = com[pr["b1"]], qrot(iq[pr["b1"]], np.array([0.0, 0, 1]))
            r1, a1 = b1["coll"]["r"], z1 * b1["coll"]["half"]
            x1, x2 = _closest_points(c1, a1, c2, a2)
            p1, p2 = c1 + x1 * a1, c2 + x2 * a2
            cd = np.linalg.norm(p2 - p1)
            dist = cd - r1 - r2
            if dist < 0 and cd > 1e-15:
                n = (p2 - p1) / cd
                found.append((dist, p1 + n * (r1 + 0.5 * dist), n, None))
            solref, solimp = b2["coll"]["solref"], b2["coll"]["solimp"]
            bodies = (pr["b1"], pr["b2"])
        tc, dr, dmax = max(solref[0], 2 * model["h"]), solref[1], solimp[1]
        Kk, Bb = 1.0 / (dmax * dmax * tc * tc * dr * dr), 2.0 / (dmax * tc)
        for dist, pos, n, pref in found:
            y = None
            if pref is not None:
                y = pref - (pref @ n) * n
                y = None if np.linalg.norm(y) < 1e-12 else y
            if y is None:
                y = np.array([0.0, 1, 0]) if -0.5 < n[1] < 0.5 else np.array([0.0, 0, 1])
                y = y - (y @ n) * n
            t1 = y / np.linalg.norm(y)
            t2 = np.cross(n, t1)
            Jd = _material_point_jacobian(model, q, bodies[1], pos)
            if bodies[0] is not None:
                Jd = Jd - _material_point_jacobian(model, q, bodies[0], pos)
            imp = impedance(solimp, dist)
            mu = pr["mu"]
            R = 2 * mu * mu * max(MINVAL, (1 - imp) / imp * (1 + mu * mu) * pr["invw"])
            for t, sg in ((t1, 1.0), (t1, -1.0), (t2, 1.0), (t2, -1.0)):
                Jr = (n + sg * mu * t) @ Jd
                rows.append(dict(J=Jr, D=1.0 / R, aref=-Bb * (Jr @ v) - Kk * imp * dist, dist=dist))
    return rows


def solve_dense_rows(M, f, rows):
    """Enumerate active sets of dense rows (identical rows switch together: they have the same residual)."""
    groups = []
    for r in rows:
        for g in groups:
            if np.allclose(g[0]["J"], r["J"], rtol=0, atol=1e-14) and g[0]["aref"] == r["aref"]:
                g.append(r)
                break
        else:
            groups.append([r])
    for active in itertools.product((False, True), repeat=len(groups)):
        H, rhs = M.copy(), f.copy()
        for g, on in zip(groups, active):
            if on:
                for r in g:
                    H += r["D"] * np.outer(r["J"], r["J"])
                    rhs += r["D"] * r["aref"] * r["J"]
        a = np.linalg.solve(H, rhs)
        res = [g[0]["J"] @ a - g[0]["aref"] for g in groups]
        if all((x < 0) == on for x, on in zip(res, active)):
            fc = np.zeros_like(f)
            for g, x, on in zip(groups, res, active):
                if on:
                    for r in g:
                        fc += r["J"] * (-r["D"] * x)
            return a, fc
    raise RuntimeError("no consistent active set")


def step(model, q, v, u):
    """One mj_step; returns (q', v', info)."""
    q, v = np.asarray(q, float), np.asarray(v, float)
    M = mass_matrix_cs(model, q)
    bias = bias_forces(model, q, v)
    damp = np.array([j["damping"] for _, j in model["dofs"]])
    stiff = np.array([j["stiffness"] for _, j in model["dofs"]])
    sref = np.array([j["springref"] for _, j in model["dofs"]])
    passive = -stiff * (q - sref) - damp * v + fluid_forces(model, q, v)
    act = np.zeros(model["nv"])
    for mtr, ui in zip(model["motors"], u):
        act[mtr["dof"]] += mtr["gear"] * min(max(ui, mtr["range"][0]), mtr["range"][1])
    f = passive + act - bias
    rows = limit_rows(model, q, v)
    crows = contact_rows(model, q, v) if model.get("pairs") else []
    fc = np.zeros(model["nv"])
    if crows:
        dense = [dict(J=r["sign"] * np.eye(model["nv"])[r["dof"]], D=r["D"], aref=r["aref"]) for r in rows] + crows
        _, fc = solve_dense_rows(M, f, dense)
        rows = dense
    elif rows:
        _, fc = solve_rows(M, f, rows)
    qacc = np.linalg.solve(M + model["h"] * np.diag(damp), f + fc)
    v2 = v + model["h"] * qacc
    return q + model["h"] * v2, v2, dict(M=M, bias=bias, passive=passive, actuation=act, constraint=fc, qacc=qacc,
                                         nefc=len(rows))
