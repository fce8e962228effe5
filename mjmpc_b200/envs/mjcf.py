"""Minimal MJCF reader for the model family the rollout kernel supports.

Reads the subset of MJCF that the reference's reacher model uses (reference:
``mjmpc/envs/assets/xml/sawyer.xml``): ``compiler inertiafromgeom``, ``option
timestep/gravity``, ``default/joint`` and ``default/geom``, nested ``body`` with
``geom`` (sphere, capsule ``fromto``, plane), one hinge ``joint`` per body,
``site`` and ``actuator/motor``.  Anything else raises, so an unsupported model
is rejected instead of being simulated wrongly.
"""
from __future__ import annotations

import xml.etree.ElementTree as ET

from .model import ActuatorSpec, BodySpec, GeomSpec, JointSpec, ModelSpec


def _floats(s):
    return tuple(float(x) for x in s.split())


def load_mjcf(path: str, frame_skip: int = 2) -> ModelSpec:
    root = ET.parse(path).getroot()
    comp = root.find("compiler")
    if comp is None or comp.get("inertiafromgeom", "false") != "true":
        raise ValueError("only inertiafromgeom=true models are supported")
    if comp.get("angle", "degree") != "radian" or comp.get("coordinate", "local") != "local":
        raise ValueError("only angle=radian coordinate=local models are supported")
    opt = root.find("option")
    timestep = float(opt.get("timestep", "0.002"))
    gravity = _floats(opt.get("gravity", "0 0 -9.81"))
    if opt.get("integrator", "Euler") != "Euler":
        raise ValueError("only the Euler integrator is supported")
    if any(g != 0.0 for g in gravity):
        raise ValueError("the rollout kernel is built for gravity-free models")
    dj = root.find("default/joint")
    dg = root.find("default/geom")
    jdef = dict(dj.attrib) if dj is not None else {}
    gdef = dict(dg.attrib) if dg is not None else {}

    def geom(e):
        a = dict(gdef)
        a.update(e.attrib)
        kind = a.get("type", "sphere")
        contact = int(a.get("contype", "1")) != 0 and int(a.get("conaffinity", "1")) != 0
        size = _floats(a.get("size", "0"))
        g = GeomSpec(kind, size[0], pos=_floats(a.get("pos", "0 0 0")), contact=contact,
                     name=a.get("name", ""), density=float(a.get("density", "1000")))
        if "fromto" in a:
            g.fromto = _floats(a["fromto"])
        elif kind == "capsule":
            raise ValueError("capsules must use fromto")
        if kind not in ("sphere", "capsule", "plane"):
            raise ValueError("unsupported geom type %s" % kind)
        return g

    def joint(e):
        a = dict(jdef)
        a.update(e.attrib)
        if a.get("type", "hinge") != "hinge":
            raise ValueError("only hinge joints are supported")
        if _floats(a.get("pos", "0 0 0")) != (0.0, 0.0, 0.0):
            raise ValueError("joint anchors must sit at the body origin")
        return JointSpec(a["name"], _floats(a["axis"]), _floats(a.get("range", "0 0")),
                         float(a.get("damping", "0")), float(a.get("armature", "0")),
                         a.get("limited", "false") == "true", float(a.get("frictionloss", "0")))

    bodies = []

    def walk(e, parent):
        for be in e.findall("body"):
            if any(k in be.attrib for k in ("quat", "euler", "axisangle", "xyaxes", "zaxis")):
                raise ValueError("rotated body frames are not supported")
            js = be.findall("joint")
            if len(js) > 1:
                raise ValueError("at most one joint per body")
            b = BodySpec(be.get("name"), _floats(be.get("pos", "0 0 0")), parent,
                         joint(js[0]) if js else None,
                         [geom(g) for g in be.findall("geom")],
                         {s.get("name"): _floats(s.get("pos", "0 0 0")) for s in be.findall("site")})
            bodies.append(b)
            walk(be, len(bodies) - 1)

    wb = root.find("worldbody")
    walk(wb, -1)
    acts = []
    for m in root.findall("actuator/motor"):
        acts.append(ActuatorSpec(m.get("joint"), float(m.get("gear", "1").split()[0]),
                                 _floats(m.get("ctrlrange", "0 0")),
                                 m.get("ctrllimited", "false") == "true"))
    margin = float(gdef.get("margin", "0"))
    return ModelSpec(bodies, acts, timestep=timestep, gravity=gravity, frame_skip=frame_skip,
                     geom_margin=margin,
                     world_geoms=[geom(g) for g in wb.findall("geom")],
                     world_sites={s.get("name"): _floats(s.get("pos", "0 0 0"))
                                  for s in wb.findall("site")})
