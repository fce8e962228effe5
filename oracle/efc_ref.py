"""oracle/efc_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A SECOND, independent restatement of one MuJoCo 2.0 ``mj_step`` of the reference's reacher model, written to pin
the soft-constraint half of ``oracle/mjstep.c`` (and of the CUDA kernel) against something that shares no code
and no pre-computed constant with either:

  * it reads ``mjmpc/envs/assets/xml/sawyer.xml`` itself (xml.etree) and applies MuJoCo's compiler rules
    (``inertiafromgeom``: density 1000, sphere / capsule inertia, parallel axis) -- it does NOT import
    ``mjmpc_b200.envs.model`` / ``mjcf`` and does not take their ``dof_invweight0`` / ``con_invweight`` / ``K`` / ``B``;
  * M(q) comes from geometric Jacobians (sum_b m Jv'Jv + Jw' R I R' Jw + armature), the Coriolis / centrifugal
    bias from the Christoffel form with dM/dq by complex-step differentiation -- no recursion, no spatial algebra;
  * qpos0 constants as ``mj_setConst`` documents them: dof_invweight0 = diag(M(qpos0)^-1), body_invweight0
    (translation) = tr(Jcom M0^-1 Jcom')/3;
  * rows as ``mj_instantiateLimit`` / ``mj_instantiateContact`` / ``mj_makeImpedance`` / ``mj_referenceConstraint``
    document them (file-level defaults solref = (0.02, 1), solimp = (0.9, 0.95, 0.001, 0.5, 2); REFSAFE);
  * the convex problem  min_a 1/2 (a-a0)'M(a-a0) + sum_r 1/2 D_r min(0, J_r a - aref_r)^2  is solved by ENUMERATING
    the 2^nefc active sets and keeping the one that satisfies its own KKT conditions (mjstep.c: Newton + exact
    line search; the kernel: guessed active set + rank-one repair) -- a third method;
  * mj_Euler with implicit joint damping.

Reference call sites this follows: mjmpc/envs/basic/reacher_env.py:21,29-39 (frame_skip 2, stale site_xpos
cost), mjmpc/envs/gym_env_wrapper.py:133-151 (u = mean[t] + noise, unclipped action recorded),
mjmpc/envs/assets/xml/sawyer.xml (the model).  MuJoCo itself is absent from the reference tree and this image:
this file narrows "parity unpinned" to "two independent restatements of MuJoCo's documented algorithm agree";
the committed vectors ``tests/golden/efc_pin.npz`` are ITS outputs (tests/golden/gen_efc_pin.py).
"""
from __future__ import annotations

import itertools
import math
import xml.etree.ElementTree as ET

import numpy as np

MJMINVAL = 1e-15
DEFAULT_SOLREF = (0.02, 1.0)
DEFAULT_SOLIMP = (0.9, 0.95, 0.001, 0.5, 2.0)


def _f(s):
    return np.array([float(x) for x in s.split()], float)


# ------------------------------------------------------------------------------------------------ model
def read_model(xml_path: str, frame_skip: int = 2) -> dict:
    root = ET.parse(xml_path).getroot()
    opt = root.find("option").attrib
    jd = dict(root.find("default/joint").attrib)
    gd = dict(root.find("default/geom").attrib)
    bodies = []

    def geom_inertial(a):
        """(mass, centre, inertia about the centre) of one geom in the body frame; MuJoCo compiler rules."""
        rho = float(a.get("density", "1000"))
        kind = a.get("type", "sphere")
        r = _f(a["size"])[0]
        if kind == "sphere":
            m = rho * 4.0 / 3.0 * math.pi * r ** 3
            return m, _f(a.get("pos", "0 0 0")), 0.4 * m * r * r * np.eye(3)
        if kind == "capsule":
            ft = _f(a["fromto"])
            p0, p1 = ft[:3], ft[3:]
            h = float(np.linalg.norm(p1 - p0))
            ax = (p1 - p0) / h
            mc = rho * math.pi * r * r * h
            ms = rho * 4.0 / 3.0 * math.pi * r ** 3
            ia = 0.5 * mc * r * r + 0.4 * ms * r * r
            it = mc * (3 * r * r + h * h) / 12.0 + ms * (0.4 * r * r + 0.25 * h * h + 0.375 * r * h)
            P = np.outer(ax, ax)
            return mc + ms, 0.5 * (p0 + p1), it * (np.eye(3) - P) + ia * P
        raise ValueError(kind)

    def walk(e, parent):
        for be in e.findall("body"):
            geoms = []
            for g in be.findall("geom"):
                a = dict(gd)
                a.update(g.attrib)
                geoms.append(a)
            parts = [geom_inertial(a) for a in geoms]
            m = sum(p[0] for p in parts)
            com = sum(p[0] * p[1] for p in parts) / m
            I = np.zeros((3, 3))
            for mg, c, Ig in parts:
                d = c - com
                I += Ig + mg * (d @ d * np.eye(3) - np.outer(d, d))
            j = None
            js = be.findall("joint")
            if js:
                a = dict(jd)
                a.update(js[0].attrib)
                j = dict(axis=_f(a["axis"]), range=_f(a["range"]), damping=float(a["damping"]),
                         armature=float(a["armature"]), limited=a.get("limited") == "true", name=a["name"])
            con = [a for a in geoms if int(a["contype"]) and int(a["conaffinity"])]
            sites = {s.get("name"): _f(s.get("pos", "0 0 0")) for s in be.findall("site")}
            bodies.append(dict(name=be.get("name"), parent=parent, pos=_f(be.get("pos", "0 0 0")), mass=m, com=com,
                               inertia=I, joint=j, contact=con, sites=sites))
            walk(be, len(bodies) - 1)

    wb = root.find("worldbody")
    walk(wb, -1)
    plane = None
    for g in wb.findall("geom"):
        a = dict(gd)
        a.update(g.attrib)
        if a.get("type") == "plane" and int(a["contype"]) and int(a["conaffinity"]):
            plane = dict(z=_f(a["pos"])[2], margin=float(a.get("margin", "0")))
    jbody = [i for i, b in enumerate(bodies) if b["joint"] is not None]
    jname = [bodies[i]["joint"]["name"] for i in jbody]
    gear = np.zeros(len(jbody)); lo = np.zeros(len(jbody)); hi = np.zeros(len(jbody))
    for mtr in root.findall("actuator/motor"):
        k = jname.index(mtr.get("joint"))
        gear[k] = float(mtr.get("gear"))
        lo[k], hi[k] = _f(mtr.get("ctrlrange"))
        assert mtr.get("ctrllimited") == "true"
    m = dict(bodies=bodies, jbody=jbody, nv=len(jbody), gear=gear, ctrl_lo=lo, ctrl_hi=hi, plane=plane,
             timestep=float(opt["timestep"]), frame_skip=frame_skip, solref=DEFAULT_SOLREF, solimp=DEFAULT_SOLIMP)
    assert _f(opt["gravity"]).tolist() == [0.0, 0.0, 0.0] and opt.get("integrator", "Euler") == "Euler"
    m["armature"] = np.array([bodies[i]["joint"]["armature"] for i in jbody])
    m["damping"] = np.array([bodies[i]["joint"]["damping"] for i in jbody])
    m["range"] = np.array([bodies[i]["joint"]["range"] for i in jbody])
    m["limited"] = np.array([bodies[i]["joint"]["limited"] for i in jbody])
    m["hand"] = next((i, b["sites"]["finger"]) for i, b in enumerate(bodies) if "finger" in b["sites"])
    cb = [(i, a) for i, b in enumerate(bodies) for a in b["contact"]]
    assert len(cb) == 1 and cb[0][1].get("type", "sphere") == "sphere" and int(cb[0][1]["condim"]) == 1
    m["sphere"] = dict(body=cb[0][0], pos=_f(cb[0][1]["pos"]), r=_f(cb[0][1]["size"])[0], margin=float(cb[0][1]["margin"]))
    set_const(m)
    return m


def pack(m: dict) -> dict:
    """Flat numpy view of a model (what tests/golden/efc_pin.npz stores so the pin also runs without the XML)."""
    b = m["bodies"]
    return dict(parent=np.array([x["parent"] for x in b]), pos=np.array([x["pos"] for x in b]),
                mass=np.array([x["mass"] for x in b]), com=np.array([x["com"] for x in b]),
                inertia=np.array([x["inertia"] for x in b]), jbody=np.array(m["jbody"]),
                axis=np.array([b[i]["joint"]["axis"] for i in m["jbody"]]), armature=m["armature"], damping=m["damping"],
                range=m["range"], limited=m["limited"], gear=m["gear"], ctrl_lo=m["ctrl_lo"], ctrl_hi=m["ctrl_hi"],
                dof_invweight0=m["dof_invweight0"], con_invweight=np.float64(m["con_invweight"]),
                solK=np.float64(m["K"]), solB=np.float64(m["B"]))


# ------------------------------------------------------------------------------------------------ kinematics
def _rodrigues(axis, q):
    a = np.asarray(axis, float)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(q) * K + (1 - np.cos(q)) * (K @ K)


def kinematics(m, q):
    """Per body: world rotation, origin, list of ancestor dofs; per dof: world axis and anchor."""
    nb = len(m["bodies"])
    R, p, dofs = [None] * nb, [None] * nb, [None] * nb
    ax, an = [None] * m["nv"], [None] * m["nv"]
    for i, b in enumerate(m["bodies"]):
        pa = b["parent"]
        Rp = np.eye(3) if pa < 0 else R[pa]
        pp = np.zeros(3) if pa < 0 else p[pa]
        p[i] = pp + Rp @ b["pos"]
        R[i] = Rp
        dofs[i] = [] if pa < 0 else list(dofs[pa])
        if b["joint"] is not None:
            j = m["jbody"].index(i)
            ax[j] = Rp @ b["joint"]["axis"]
            an[j] = p[i]
            R[i] = Rp @ _rodrigues(b["joint"]["axis"], q[j])
            dofs[i].append(j)
    return R, p, dofs, ax, an


def point_jacobian(m, kin, body, point_world):
    R, p, dofs, ax, an = kin
    J = np.zeros((3, m["nv"]), dtype=np.result_type(point_world, *[ax[j] for j in dofs[body]]))
    for j in dofs[body]:
        J[:, j] = np.cross(ax[j], point_world - an[j])
    return J


def mass_matrix(m, q):
    kin = kinematics(m, q)
    R, p, dofs, ax, an = kin
    M = np.diag(m["armature"]).astype(np.asarray(q).dtype if np.iscomplexobj(q) else float)
    for i, b in enumerate(m["bodies"]):
        c = p[i] + R[i] @ b["com"]
        Jv = point_jacobian(m, kin, i, c)
        Jw = np.zeros((3, m["nv"]), dtype=M.dtype)
        for j in dofs[i]:
            Jw[:, j] = ax[j]
        M += b["mass"] * Jv.T @ Jv + Jw.T @ (R[i] @ b["inertia"] @ R[i].T) @ Jw
    return M


def bias(m, q, v):
    """c_i = sum_jk (dM_ij/dq_k - 1/2 dM_jk/dq_i) v_j v_k with dM/dq by complex-step differentiation of the
    Jacobian-sum mass matrix (M is analytic in q; step 1e-30, so the derivative is exact to rounding)."""
    nv = m["nv"]
    q = np.asarray(q, float)
    v = np.asarray(v, float)
    dM = np.zeros((nv, nv, nv))                                  # [i, j, k] = dM_ij / dq_k
    for k in range(nv):
        qc = q.astype(complex)
        qc[k] += 1e-30j
        dM[:, :, k] = mass_matrix(m, qc).imag / 1e-30
    return np.einsum("ijk,j,k->i", dM, v, v) - 0.5 * np.einsum("jki,j,k->i", dM, v, v)


# ------------------------------------------------------------------------------------------------ mj_setConst
def set_const(m):
    q0 = np.zeros(m["nv"])
    Minv = np.linalg.inv(mass_matrix(m, q0))
    m["dof_invweight0"] = np.diag(Minv).copy()                 # hinge: one dof, no averaging
    kin = kinematics(m, q0)
    R, p = kin[0], kin[1]
    b = m["sphere"]["body"]
    J = point_jacobian(m, kin, b, p[b] + R[b] @ m["bodies"][b]["com"])      # mj_jacBodyCom, translational rows
    A = J @ Minv @ J.T
    m["con_invweight"] = float(np.trace(A) / 3.0)              # + 0 for the world body that owns the plane
    # mj_makeImpedance, standard solref; mjDSBL_REFSAFE off: time constant >= 2 * timestep
    tc = max(m["solref"][0], 2.0 * m["timestep"])
    dmax = m["solimp"][1]
    m["K"] = 1.0 / max(MJMINVAL, dmax * dmax * tc * tc * m["solref"][1] ** 2)
    m["B"] = 2.0 / max(MJMINVAL, dmax * tc)


# ------------------------------------------------------------------------------------------------ constraint rows
def impedance(solimp, x_signed):
    d0, dw, width, mid, power = solimp
    if d0 == dw or width <= MJMINVAL:
        return 0.5 * (d0 + dw)
    x = abs(x_signed / width)
    if x >= 1.0:
        return dw
    if x <= 0.0:
        return d0
    if power == 1.0:
        y = x
    elif x <= mid:
        y = x ** power / mid ** (power - 1.0)
    else:
        y = 1.0 - (1.0 - x) ** power / (1.0 - mid) ** (power - 1.0)
    return d0 + y * (dw - d0)


def rows(m, q, v):
    """efc rows in MuJoCo's order (limits by joint, then contacts): J, pos - margin, diagApprox, D, aref."""
    nv = m["nv"]
    out = []
    for j in range(nv):
        if not m["limited"][j]:
            continue
        lo, hi = m["range"][j]
        for dist, sg in ((q[j] - lo, 1.0), (hi - q[j], -1.0)):
            if dist < 0.0:                                     # jnt_margin = 0
                J = np.zeros(nv)
                J[j] = sg
                out.append(dict(J=J, pos=dist, diag=m["dof_invweight0"][j], kind="limit", dof=j))
    if m["plane"] is not None:
        kin = kinematics(m, q)
        R, p = kin[0], kin[1]
        s = m["sphere"]
        c = p[s["body"]] + R[s["body"]] @ s["pos"]
        dist = c[2] - m["plane"]["z"] - s["r"]
        margin = max(s["margin"], m["plane"]["margin"])        # MuJoCo 2.0: max of the two geom margins
        if dist < margin:
            cp = c - np.array([0.0, 0.0, 1.0]) * (s["r"] + 0.5 * dist)      # mid-surface contact point
            J = point_jacobian(m, kin, s["body"], cp)[2]       # frame normal = plane normal = +z; plane body is static
            out.append(dict(J=J, pos=dist - margin, diag=m["con_invweight"], kind="contact"))
    for r in out:
        imp = impedance(m["solimp"], r["pos"])
        R_ = max(MJMINVAL, (1.0 - imp) * r["diag"] / imp)
        r["imp"], r["D"] = imp, 1.0 / R_
        r["aref"] = -m["B"] * float(r["J"] @ v) - m["K"] * imp * r["pos"]
    return out


def solve_active_sets(M, f, rws):
    """Exact minimiser by enumeration: the active set A is optimal iff the solution of
    (M + sum_A D J'J) a = f + sum_A D aref J'  has J_r a - aref_r < 0 exactly for r in A."""
    n = len(rws)
    for mask in itertools.product((False, True), repeat=n):
        H, g = M.copy(), f.copy()
        for on, r in zip(mask, rws):
            if on:
                H += r["D"] * np.outer(r["J"], r["J"])
                g += r["D"] * r["aref"] * r["J"]
        a = np.linalg.solve(H, g)
        if all(((r["J"] @ a - r["aref"]) < 0.0) == on for on, r in zip(mask, rws)):
            force = np.array([-r["D"] * (r["J"] @ a - r["aref"]) if on else 0.0 for on, r in zip(mask, rws)])
            return a, force, np.array(mask)
    raise RuntimeError("no self-consistent active set (degenerate tie)")


def step(m, q, v, u):
    """One mj_step.  Returns (q', v', info)."""
    q, v = np.asarray(q, float), np.asarray(v, float)
    M = mass_matrix(m, q)
    c = bias(m, q, v)
    f = m["gear"] * np.clip(u, m["ctrl_lo"], m["ctrl_hi"]) - m["damping"] * v - c        # qfrc_smooth (no gravity)
    rws = rows(m, q, v)
    fc = np.zeros(m["nv"])
    info = dict(M=M, bias=c, nefc=len(rws), rows=rws)
    if rws:
        a, force, mask = solve_active_sets(M, f, rws)
        fc = sum(fr * r["J"] for fr, r in zip(force, rws))
        info.update(qacc_constrained=a, efc_force=force, active=mask)
    h = m["timestep"]
    qacc = np.linalg.solve(M + h * np.diag(m["damping"]), f + fc)                        # mj_Euler, implicit damping
    v2 = v + h * qacc
    q2 = q + h * v2
    info.update(qfrc_constraint=fc, qacc=qacc)
    return q2, v2, info


def hand_position(m, q):
    R, p = kinematics(m, q)[:2]
    b, s = m["hand"]
    return p[b] + R[b] @ s


def rollout(m, q0, v0, target, mean, noise):
    """The reference rollout of ONE particle (gym_env_wrapper.py:125-153 around reacher_env.py:29-39)."""
    H = mean.shape[0]
    q, v = np.array(q0, float), np.array(v0, float)
    costs, qv, nrow = np.zeros(H), np.zeros((H, 2 * m["nv"])), 0
    for t in range(H):
        u = mean[t] + noise[t]
        for s in range(m["frame_skip"]):
            hand = hand_position(m, q)                          # site_xpos of the last forward pass: before the integration
            q, v, info = step(m, q, v, u)
            nrow += info["nefc"] > 0
        d = hand - target
        costs[t] = np.abs(d).sum() + 5.0 * np.sqrt((d * d).sum())
        qv[t] = np.concatenate([q, v])
    return costs, qv, nrow
