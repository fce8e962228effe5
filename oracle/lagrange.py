"""Independent derivation of M(q) and the Coriolis/centrifugal bias c(q,v) from the
Lagrangian, used to cross-check oracle/mjstep.c (TEST INFRASTRUCTURE).

T(q,v) = 1/2 v' M(q) v with M(q) = diag(armature) + sum_b (m_b Jv'Jv + Jw' R I R' Jw) built
from geometric Jacobians in torch float64;  c = d/dq(M v) v - dT/dq  by autograd.
No spatial algebra, no recursion -- nothing shared with mjstep.c but the model arrays.
"""
import numpy as np
import torch


def _rot(axis, q):
    a = torch.as_tensor(axis, dtype=torch.float64)
    K = torch.zeros(3, 3, dtype=torch.float64)
    K[0, 1], K[0, 2], K[1, 0], K[1, 2], K[2, 0], K[2, 1] = -a[2], a[1], a[2], -a[0], -a[1], a[0]
    return torch.eye(3, dtype=torch.float64) + torch.sin(q) * K + (1 - torch.cos(q)) * (K @ K)


def mass_matrix(tree, q):
    nb, nv = tree.nb, tree.nv
    dof_of_body = {int(b): j for j, b in enumerate(tree.jnt_body)}
    R, p, dofs = [None] * nb, [None] * nb, [None] * nb
    axis_w, anchor = [None] * nv, [None] * nv
    I3 = torch.eye(3, dtype=torch.float64)
    for b in range(nb):
        pa = int(tree.parent[b])
        Rp = I3 if pa < 0 else R[pa]
        pp = torch.zeros(3, dtype=torch.float64) if pa < 0 else p[pa]
        p[b] = pp + Rp @ torch.as_tensor(tree.pos[b])
        R[b] = Rp
        dofs[b] = [] if pa < 0 else list(dofs[pa])
        if b in dof_of_body:
            j = dof_of_body[b]
            R[b] = Rp @ _rot(tree.jnt_axis[j], q[j])
            axis_w[j] = Rp @ torch.as_tensor(tree.jnt_axis[j])
            anchor[j] = p[b]
            dofs[b].append(j)
    M = torch.diag(torch.as_tensor(tree.armature))
    for b in range(nb):
        c = p[b] + R[b] @ torch.as_tensor(tree.ipos[b])
        Iw = R[b] @ torch.as_tensor(tree.inertia[b]) @ R[b].T
        Jv = torch.zeros(3, nv, dtype=torch.float64)
        Jw = torch.zeros(3, nv, dtype=torch.float64)
        cols_v, cols_w = [], []
        for j in range(nv):
            if j in dofs[b]:
                cols_w.append(axis_w[j])
                cols_v.append(torch.linalg.cross(axis_w[j], c - anchor[j]))
            else:
                cols_w.append(torch.zeros(3, dtype=torch.float64))
                cols_v.append(torch.zeros(3, dtype=torch.float64))
        Jv = torch.stack(cols_v, 1)
        Jw = torch.stack(cols_w, 1)
        M = M + tree.mass[b] * Jv.T @ Jv + Jw.T @ Iw @ Jw
    return M


def mass_bias(tree, q, v):
    q = torch.tensor(np.asarray(q, float), dtype=torch.float64, requires_grad=True)
    v = torch.tensor(np.asarray(v, float), dtype=torch.float64)
    M = mass_matrix(tree, q)
    T = 0.5 * v @ M @ v
    dTdq, = torch.autograd.grad(T, q, retain_graph=True)
    Mv = M @ v
    rows = [torch.autograd.grad(Mv[i], q, retain_graph=True)[0] for i in range(len(v))]
    dMv = torch.stack(rows, 0)              # d(Mv)_i / dq_j
    c = dMv @ v - dTdq
    return M.detach().numpy(), c.detach().numpy()
