"""Random MJCF hinge / slide trees for the tests of the generic path: random topology (2-6 bodies, one or
two joints each, <= 12 dofs), axes, anchors, body orientations, capsule / sphere geoms (one per body), gravity, springs,
dampers, armature, limits with random solref / solimp, motors, with or without a fluid.  Deterministic per seed."""
import numpy as np


def random_tree_xml(seed: int, planar: bool = False) -> str:
    """planar: every hinge axis along one random direction n, every slide perpendicular to it, body frames turned about n
    only -- a planar mechanism in a tilted plane, with bodies, anchors and centres of mass OFF the plane and arbitrary
    inertia tensors (the planar reduction must be exact for all of that); at most 9 dofs."""
    rng = np.random.default_rng(seed)
    n = rng.normal(0, 1, 3)
    n /= np.linalg.norm(n)
    f = lambda v: " ".join(("%.17g" if planar else "%.6g") % x for x in np.atleast_1d(v))
    nbody = int(rng.integers(2, 6 if planar else 7))
    fluid = rng.random() < 0.5
    opt = '<option timestep="%g" gravity="%s"%s/>' % (rng.choice([0.002, 0.005, 0.01]), f(rng.normal(0, 5, 3)),
                                                      ' density="%g" viscosity="%g"' % (rng.uniform(5, 500), rng.uniform(0, 0.05)) if fluid else "")
    bodies, joints, nv = [], [], 0
    parent = [-1] + [int(rng.integers(0, i)) for i in range(1, nbody)]
    for b in range(nbody):
        nj = 1 if nv >= (7 if planar else 10) else int(rng.integers(1, 3))
        js = ""
        for k in range(nj):
            kind = "slide" if rng.random() < 0.3 else "hinge"
            ax = rng.normal(0, 1, 3)
            if planar:
                ax = n * rng.choice([-1.0, 1.0]) if kind == "hinge" else ax - (ax @ n) * n
            ax /= np.linalg.norm(ax)
            name = "j%d_%d" % (b, k)
            lim = rng.random() < 0.6
            lo, hi = sorted(rng.uniform(-0.8, 0.8, 2))
            js += ('<joint name="%s" type="%s" axis="%s" pos="%s" damping="%g" armature="%g" stiffness="%g" springref="%g" limited="%s" '
                   'range="%g %g" solreflimit="%g %g" solimplimit="%g %g %g %g %g"/>'
                   % (name, kind, f(ax), f(rng.normal(0, 0.05, 3)), rng.choice([0.0, rng.uniform(0, 1)]), rng.uniform(0, 0.05),
                      rng.choice([0.0, rng.uniform(0, 10)]), rng.uniform(-0.2, 0.2), "true" if lim else "false", lo, hi + 0.05,
                      rng.uniform(0.02, 0.05), rng.uniform(0.7, 1.2), rng.uniform(0.5, 0.9), rng.uniform(0.9, 0.99),
                      rng.uniform(0.001, 0.02), rng.uniform(0.3, 0.7), rng.choice([1.0, 2.0, 3.0])))
            joints.append(name)
            nv += 1
        if rng.random() < 0.75:
            geom = '<geom type="capsule" fromto="%s %s" size="%g" density="%g"/>' % (f(rng.normal(0, 0.1, 3)), f(rng.normal(0, 0.2, 3) + 0.15),
                                                                                   rng.uniform(0.02, 0.06), rng.uniform(500, 2000))
        else:
            geom = '<geom type="sphere" pos="%s" size="%g"/>' % (f(rng.normal(0, 0.1, 3)), rng.uniform(0.04, 0.08))
        q = rng.normal(0, 1, 4)
        if planar:                                   # a turn about n
            ang = rng.uniform(-np.pi, np.pi)
            q = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * n])
        bodies.append(dict(open='<body name="b%d" pos="%s" quat="%s">%s%s' % (b, f(rng.normal(0, 0.3, 3)), f(q / np.linalg.norm(q)), js, geom)))
    children = {i: [j for j in range(nbody) if parent[j] == i] for i in range(-1, nbody)}

    def emit(i):
        return bodies[i]["open"] + "".join(emit(c) for c in children[i]) + "</body>"
    motors = ""
    for name in joints:
        if rng.random() < 0.5:
            motors += '<motor joint="%s" gear="%g" ctrllimited="%s" ctrlrange="-1 1"/>' % (name, rng.uniform(0.5, 5), "true" if rng.random() < 0.5 else "false")
    if not motors:
        motors = '<motor joint="%s" gear="1"/>' % joints[0]
    return ('<mujoco model="rand%d"><compiler angle="radian" inertiafromgeom="true"/><default><geom contype="0" conaffinity="0"/></default>%s'
            '<worldbody>%s</worldbody><actuator>%s</actuator></mujoco>' % (seed, opt, "".join(emit(c) for c in children[-1]), motors))


def random_contact_mechanism_xml(seed: int) -> str:
    """A random planar mechanism WITH contacts: 5 bodies / 9 dofs (the kernel's contact instantiation for trees), everything
    in one world coordinate plane (x-y, y-z or z-x by seed), one capsule per body that collides with the capsules of non-adjacent bodies and
    with a floor whose normal lies in the plane; random friction, solref, solimp (one set for all geoms), gravity towards
    the floor, limits, dampers."""
    rng = np.random.default_rng(seed)
    f = lambda v: " ".join("%.17g" % x for x in np.atleast_1d(v))
    n = np.eye(3)[seed % 3] * (1.0 if seed % 2 else -1.0)       # a world coordinate plane (capsule-capsule contacts need one)
    ex = np.cross(n, [0.0, 0.0, 1.0]) if abs(n[2]) < 0.9 else np.cross(n, [1.0, 0.0, 0.0])
    ex /= np.linalg.norm(ex)
    ey = np.cross(n, ex)
    P = lambda a, b: a * ex + b * ey                            # a point / direction of the plane
    up = P(*rng.normal(0, 1, 2))
    up /= np.linalg.norm(up)                                     # floor normal, in the plane
    geomdef = ('<geom condim="3" friction="%g 0.1 0.1" solref="%g %g" solimp="%g %g %g"/>'
               % (rng.uniform(0.2, 1.2), rng.uniform(0.02, 0.04), rng.uniform(0.8, 1.1), rng.uniform(0.0, 0.9), rng.uniform(0.9, 0.98),
                  rng.uniform(0.001, 0.02)))
    parent = [-1, 0, int(rng.integers(0, 2)), int(rng.integers(0, 3)), int(rng.integers(1, 4))]
    njoint = [2, 2, 2, 2, 1]
    bodies, joints = [], []
    for b in range(5):
        js = ""
        for k in range(njoint[b]):
            kind = "slide" if (b == 0 or rng.random() < 0.2) and k == 0 else "hinge"
            ax = n * rng.choice([-1.0, 1.0]) if kind == "hinge" else P(*rng.normal(0, 1, 2))
            ax = ax / np.linalg.norm(ax)
            name = "j%d_%d" % (b, k)
            lim = b > 0 and rng.random() < 0.6
            js += ('<joint name="%s" type="%s" axis="%s" pos="%s" damping="%g" armature="%g" limited="%s" range="%g %g"/>'
                   % (name, kind, f(ax), f(P(*rng.normal(0, 0.03, 2))), rng.choice([0.0, rng.uniform(0, 0.5)]), rng.uniform(0, 0.05),
                      "true" if lim else "false", rng.uniform(-0.9, -0.2), rng.uniform(0.2, 0.9)))
            joints.append(name)
        a0, d = P(*rng.normal(0, 0.05, 2)), P(*rng.normal(0, 1, 2))
        d = d / np.linalg.norm(d) * rng.uniform(0.15, 0.35)
        geom = '<geom type="capsule" fromto="%s %s" size="%g"/>' % (f(a0), f(a0 + d), rng.uniform(0.03, 0.06))
        ang = rng.uniform(-np.pi, np.pi)
        q = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * n])
        pos = P(*rng.normal(0, 0.25, 2)) + (up * 0.35 if b == 0 else 0.0)
        bodies.append('<body name="b%d" pos="%s" quat="%s">%s%s' % (b, f(pos), f(q), js, geom))
    children = {i: [j for j in range(5) if parent[j] == i] for i in range(-1, 5)}

    def emit(i):
        return bodies[i] + "".join(emit(c) for c in children[i]) + "</body>"
    R = np.stack([np.cross(up, n), -n if False else np.cross(up, np.cross(up, n)) * 0 + n, up], axis=1)     # columns x, y, z = up
    R[:, 0] = np.cross(R[:, 1], R[:, 2])
    w = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    if w > 1e-6:
        fq = np.array([w, (R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w)])
        floor = '<geom type="plane" pos="%s" quat="%s" size="5 5 1"/>' % (f(-0.1 * up), f(fq))
    else:
        floor = '<geom type="plane" pos="%s" zaxis="%s" size="5 5 1"/>' % (f(-0.1 * up), f(up))
    motors = "".join('<motor joint="%s" gear="%g"/>' % (name, rng.uniform(0.5, 4)) for name in joints[2:] if rng.random() < 0.6) \
        or '<motor joint="%s" gear="1"/>' % joints[2]
    return ('<mujoco model="randc%d"><compiler angle="radian" inertiafromgeom="true"/><default>%s</default>'
            '<option timestep="0.005" gravity="%s"/><worldbody>%s%s</worldbody><actuator>%s</actuator></mujoco>'
            % (seed, geomdef, f(-9.0 * up + P(*rng.normal(0, 1, 2))), floor, emit(0), motors))
