"""Generates tests/golden/tree_pin.npz: outputs of oracle/tree_ref.py -- the SECOND, independent restatement of
MuJoCo's mj_step for hinge / slide trees (own MJCF reader on quaternions, complex-step Jacobians, Jacobian-sum mass
matrix, Christoffel bias, potential-gradient gravity, inertia-box fluid forces, active-set enumeration) -- on

  * the reference's swimmer.xml                      (interior states, joints beyond +-1.5, clamped controls)
  * tests/fixtures/tree3d.xml, tree3d_weld.xml        (branches, skew axes, anchors, rotated frames, gravity, springs,
                                                       dampers, per-joint solref / solimp, a fluid / a welded body)

    python tests/golden/gen_tree_pin.py        (needs /root/reference for swimmer.xml; run in the build container)

tests/test_tree_cpu.py pins oracle/tree_step.c and mjmpc_b200/envs/mjcf_tree.py's compiled constants against these
vectors; tests/test_tree_gpu.py pins the CUDA kernel against tree_step.c.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import tree_ref  # noqa: E402

MODELS = {"swimmer": "/root/reference/mjmpc/envs/assets/xml/swimmer.xml",
          "tree3d": os.path.join(ROOT, "tests", "fixtures", "tree3d.xml"),
          "tree3d_weld": os.path.join(ROOT, "tests", "fixtures", "tree3d_weld.xml"),
          # the contact half: the same swimmer with its capsule-capsule contacts (folded poses), a walker on a floor
          "swimmer_contact": "/root/reference/mjmpc/envs/assets/xml/swimmer.xml",
          "walker": os.path.join(ROOT, "tests", "fixtures", "planar_walker.xml")}


def main():
    out = {}
    rng = np.random.default_rng(20260101)
    for name, xml in MODELS.items():
        contact = name in ("swimmer_contact", "walker")
        m = tree_ref.read_model(xml, contacts=contact)
        nv, nu = m["nv"], len(m["motors"])
        n = 24 if contact else 48
        Q, V, U = np.zeros((n, nv)), np.zeros((n, nv)), np.zeros((n, nu))
        keys = ("M", "bias", "passive", "actuation", "constraint", "qacc")
        rec = {k: [] for k in keys}
        Q2, V2, NE = np.zeros((n, nv)), np.zeros((n, nv)), np.zeros(n, int)
        for i in range(n):
            q = rng.uniform(-1.0, 1.0, nv)
            if i % 3:                       # push some limited joints beyond their range, shallow and deep
                for k, (_, j) in enumerate(m["dofs"]):
                    if j["limited"] and rng.random() < 0.5:
                        d = rng.uniform(0, 0.003 if i % 3 == 1 else 0.2)
                        q[k] = j["range"][0] - d if rng.random() < 0.5 else j["range"][1] + d
            if name == "swimmer_contact":     # three or four joints folded to one side: non-adjacent capsules overlap
                q[3:7] = rng.choice([-1.0, 1.0]) * rng.uniform(1.3, 1.6, 4)
                if i % 4 == 0:
                    q[6] = -q[6]
            if name == "walker":              # feet at / into the floor (rest height of the feet is ~0.2 m above it)
                q[:3] = [rng.uniform(-1, 1), rng.uniform(-0.3, -0.1), rng.uniform(-0.4, 0.4)]
            v = rng.normal(0, 2.0 if i % 2 else 0.2, nv)
            u = rng.normal(0, 1.2, nu)
            q2, v2, info = tree_ref.step(m, q, v, u)
            Q[i], V[i], U[i], Q2[i], V2[i], NE[i] = q, v, u, q2, v2, info["nefc"]
            for k in keys:
                rec[k].append(info[k])
        out.update({name + "_q": Q, name + "_v": V, name + "_u": U, name + "_q2": Q2, name + "_v2": V2, name + "_nefc": NE,
                    name + "_invweight0": m["invweight0"],
                    name + "_mass": np.array([b["mass"] for b in m["bodies"]]),
                    name + "_inertia": np.array([b["inertia"] for b in m["bodies"]])})
        out.update({name + "_" + k: np.array(rec[k]) for k in keys})
        # a short free rollout (frame_skip 4, forward-progress reward) for the swimmer
        if name == "swimmer":
            H = 6
            q, v = rng.uniform(-.1, .1, nv), rng.uniform(-.1, .1, nv)
            acts = rng.normal(0, 0.8, (H, nu))
            traj, rew = [], []
            out["swimmer_roll_state0"] = np.concatenate([q, v])
            for t in range(H):
                x0 = q[0]
                for _ in range(4):
                    q, v, _ = tree_ref.step(m, q, v, acts[t])
                traj.append(np.concatenate([q, v]))
                rew.append((q[0] - x0) / (4 * m["h"]) - 1e-4 * np.square(acts[t]).sum())
            out["swimmer_roll_actions"], out["swimmer_roll_states"], out["swimmer_roll_rewards"] = acts, np.array(traj), np.array(rew)
    path = os.path.join(ROOT, "tests", "golden", "tree_pin.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.endswith("_nefc")},
          {k: int((v > 0).sum()) for k, v in out.items() if k.endswith("_nefc")})


if __name__ == "__main__":
    main()
