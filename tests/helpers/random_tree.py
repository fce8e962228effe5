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
