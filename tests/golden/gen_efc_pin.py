"""Generates tests/golden/efc_pin.npz: outputs of oracle/efc_ref.py -- the SECOND, independent restatement of
MuJoCo's mj_step for the reference's reacher model (own MJCF reader, Jacobian-sum mass matrix, complex-step
Coriolis terms, qpos0 constants, constraint rows, active-set enumeration) -- on states that exercise the
soft-constraint path: joints beyond their limits, the env's reset state, the end-effector sphere on the table.

    python tests/golden/gen_efc_pin.py        (needs /root/reference for sawyer.xml; run in the build container)

tests/test_efc_pin.py pins oracle/mjstep.c, mjmpc_b200/envs/model.py's compiled constants and (on the GPU) the
CUDA rollout against these vectors.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import efc_ref  # noqa: E402

XML = "/root/reference/mjmpc/envs/assets/xml/sawyer.xml"
MAXROW = 15


def states(m, rng, n):
    lo, hi = m["range"][:, 0], m["range"][:, 1]
    out = []
    for i in range(n):
        kind = i % 6
        if kind == 0:                       # interior
            q = rng.uniform(lo + 0.1 * (hi - lo), hi - 0.1 * (hi - lo))
        elif kind == 1:                     # the env's reset pose (elbow / wrist-flex sit ON their upper limit), nudged
            q = np.zeros(7) + rng.normal(0, 0.01, 7)
        elif kind in (2, 3):                # 2..5 joints beyond their range: shallow (inside the impedance width) / deep
            q = rng.uniform(lo + 0.1 * (hi - lo), hi - 0.1 * (hi - lo))
            for j in rng.choice(7, size=rng.integers(2, 6), replace=False):
                d = rng.uniform(0, 0.002 if kind == 2 else 0.15)
                q[j] = lo[j] - d if rng.random() < 0.5 else hi[j] + d
        else:                               # sphere at the table: best of 400 random poses, then a limit violation on top
            best = None
            for _ in range(400):
                qq = rng.uniform(lo, hi)
                R, p = efc_ref.kinematics(m, qq)[:2]
                c = p[m["sphere"]["body"]] + R[m["sphere"]["body"]] @ m["sphere"]["pos"]
                d = c[2] - m["plane"]["z"] - m["sphere"]["r"]
                if best is None or abs(d + 0.001) < abs(best[0] + 0.001):
                    best = (d, qq)
            q = best[1]
            if kind == 5:
                q[6] = hi[6] + 0.05
        out.append((q, rng.normal(0, 1.5, 7), rng.normal(0, 1.0, 7)))
    return out


def main():
    m = efc_ref.read_model(XML)
    rng = np.random.default_rng(20261017)
    S = states(m, rng, 72)
    n = len(S)
    g = {"model_" + k: v for k, v in efc_ref.pack(m).items()}
    g.update(q=np.zeros((n, 7)), v=np.zeros((n, 7)), u=np.zeros((n, 7)), nefc=np.zeros(n, np.int32),
             J=np.zeros((n, MAXROW, 7)), aref=np.zeros((n, MAXROW)), D=np.zeros((n, MAXROW)), imp=np.zeros((n, MAXROW)),
             force=np.zeros((n, MAXROW)), qfrc_constraint=np.zeros((n, 7)), qacc=np.zeros((n, 7)),
             q2=np.zeros((n, 7)), v2=np.zeros((n, 7)), M=np.zeros((n, 7, 7)), bias=np.zeros((n, 7)),
             has_contact=np.zeros(n, bool))
    for i, (q, v, u) in enumerate(S):
        q2, v2, info = efc_ref.step(m, q, v, u)
        r = info["rows"]
        g["q"][i], g["v"][i], g["u"][i], g["nefc"][i] = q, v, u, len(r)
        for k, row in enumerate(r):
            g["J"][i, k], g["aref"][i, k], g["D"][i, k], g["imp"][i, k] = row["J"], row["aref"], row["D"], row["imp"]
        if r:
            g["force"][i, :len(r)] = info["efc_force"]
        g["has_contact"][i] = any(x["kind"] == "contact" for x in r)
        g["qfrc_constraint"][i], g["qacc"][i], g["q2"][i], g["v2"][i] = info["qfrc_constraint"], info["qacc"], q2, v2
        g["M"][i], g["bias"][i] = info["M"], info["bias"]
    # short rollouts of the reference loop (gym_env_wrapper.py:125-153): K particles x H env steps from three starts
    K, H = 8, 6
    starts = [S[0][0], np.zeros(7), np.array([0.0, 0.57, 0, -0.2, 0, -0.3, 0.0])]     # third: sphere 1 cm above the table
    vels = [S[0][1] / 3.0, np.zeros(7), np.array([0.0, 0.5, 0, 0, 0, 0, 0.0])]
    target = np.array([0.1, 0.1, 0.1])
    noise = rng.normal(0, 1.0, (3, K, H, 7))
    noise[2] *= 0.3
    mean = np.zeros((3, H, 7))
    mean[2, :, 1] = 1.0                     # shoulder lift pushes the sphere into the table
    costs, qv, nrow = np.zeros((3, K, H)), np.zeros((3, K, H, 14)), np.zeros((3, K), np.int32)
    for s in range(3):
        for k in range(K):
            costs[s, k], qv[s, k], nrow[s, k] = efc_ref.rollout(m, starts[s], vels[s], target, mean[s], noise[s, k])
    print("substeps with constraint rows per rollout:", nrow.tolist())
    g.update(ro_q0=np.array(starts), ro_v0=np.array(vels), ro_target=target, ro_noise=noise, ro_mean=mean, ro_costs=costs, ro_qv=qv)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "efc_pin.npz")
    np.savez_compressed(out, **g)
    print(out, "states", n, "with rows", int((g["nefc"] > 0).sum()), "with contact", int(g["has_contact"].sum()),
          "max nefc", int(g["nefc"].max()))


if __name__ == "__main__":
    main()
