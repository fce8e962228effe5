"""ctypes front-end of ``oracle/mjstep.c`` (TEST INFRASTRUCTURE -- see oracle/__init__.py).

PARITY UNPINNED AGAINST MuJoCo ITSELF (pinned against the independent restatement oracle/efc_ref.py,
tests/test_efc_pin.py) for the MuJoCo dynamics (no MuJoCo binary, no golden vectors in the
reference); see the header of mjstep.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "mjstep.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.ora_model_create.restype = C.c_void_p
        L.ora_substep.restype = C.c_int
        L.ora_substep_efc.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


class OracleModel:
    """Wraps one ``ora_model`` built from a ``mjmpc_b200.envs.model.TreeModel``."""

    def __init__(self, tree):
        L = lib()
        f = lambda a: np.ascontiguousarray(a, np.float64)
        i = lambda a: np.ascontiguousarray(a, np.int32)
        K, B = _solref_kb(tree)
        self._keep = [i(tree.parent), f(tree.pos), f(tree.mass), f(tree.ipos), f(tree.inertia),
                      i(tree.jnt_body), f(tree.jnt_axis), f(tree.jnt_range), i(tree.jnt_limited),
                      f(tree.armature), f(tree.damping), f(tree.gear), f(tree.ctrlrange),
                      f(tree.dof_invweight0), f(tree.solimp), f(tree.hand_pos), f(tree.con_pos)]
        k = self._keep
        self.nv = tree.nv
        self.h = L.ora_model_create(
            C.c_int(tree.nb), C.c_int(tree.nv), _p(k[0], C.c_int), _p(k[1]), _p(k[2]), _p(k[3]), _p(k[4]),
            _p(k[5], C.c_int), _p(k[6]), _p(k[7]), _p(k[8], C.c_int), _p(k[9]), _p(k[10]), _p(k[11]),
            _p(k[12]), _p(k[13]), C.c_double(tree.timestep), C.c_int(tree.frame_skip),
            C.c_double(K), C.c_double(B), _p(k[14]), C.c_int(tree.hand_body), _p(k[15]),
            C.c_int(tree.con_body), _p(k[16]), C.c_double(tree.con_radius), C.c_double(tree.con_plane_z),
            C.c_double(tree.con_margin), C.c_double(tree.con_invweight))
        if not self.h:
            raise RuntimeError("ora_model_create failed")

    def __del__(self):
        try:
            lib().ora_model_destroy(C.c_void_p(self.h))
        except Exception:
            pass

    def mass_bias(self, q, v):
        nv = self.nv
        M = np.zeros((nv, nv)); b = np.zeros(nv); hand = np.zeros(3)
        q = np.ascontiguousarray(q, np.float64); v = np.ascontiguousarray(v, np.float64)
        lib().ora_mass_bias(C.c_void_p(self.h), _p(q), _p(v), _p(M), _p(b), _p(hand))
        return M, b, hand

    def substep(self, q, v, u):
        q = np.array(q, np.float64); v = np.array(v, np.float64); u = np.ascontiguousarray(u, np.float64)
        qacc = np.zeros(self.nv)
        n = lib().ora_substep(C.c_void_p(self.h), _p(q), _p(v), _p(u), _p(qacc))
        return q, v, qacc, n

    def substep_efc(self, q, v, u):
        """One mj_step plus its constraint rows: dict(q, v, qacc, J, aref, D, force, qfrc_constraint)."""
        nv = self.nv
        q = np.array(q, np.float64); v = np.array(v, np.float64); u = np.ascontiguousarray(u, np.float64)
        qacc, qfrc = np.zeros(nv), np.zeros(nv)
        J, aref, D, force = np.zeros((2 * nv + 1, nv)), np.zeros(2 * nv + 1), np.zeros(2 * nv + 1), np.zeros(2 * nv + 1)
        n = lib().ora_substep_efc(C.c_void_p(self.h), _p(q), _p(v), _p(u), _p(qacc), _p(J), _p(aref), _p(D), _p(force),
                                  _p(qfrc))
        return dict(q=q, v=v, qacc=qacc, J=J[:n], aref=aref[:n], D=D[:n], force=force[:n], qfrc_constraint=qfrc)


def _solref_kb(tree):
    tc = max(float(tree.solref[0]), 2.0 * tree.timestep)
    dmax = float(tree.solimp[1])
    return (1.0 / max(1e-15, dmax * dmax * tc * tc * tree.solref[1] ** 2), 2.0 / max(1e-15, dmax * tc))


def rollout(models, qpos, qvel, target, mean, noise, want_traj=False, want_obs=False, nthreads=1, policy_w=None,
            horizon=None):
    """Reference rollout on the CPU.  ``models``: one OracleModel or a list (one per
    contiguous particle block).  Returns dict(costs, actions[, qv, next_observations], ncon).
    ``policy_w`` (d_obs + 1, d_action): mode="closed_loop_linear" of gym_env_wrapper.py:135-136 (``mean`` is
    then ignored and ``horizon`` gives H when there is no noise tensor to read it from)."""
    if isinstance(models, OracleModel):
        models = [models]
    nv = models[0].nv
    noise = None if noise is None else np.ascontiguousarray(noise, np.float64)
    if policy_w is not None:
        policy_w = np.ascontiguousarray(policy_w, np.float64)
        assert policy_w.shape == (2 * nv + 7, nv)
        H = noise.shape[1] if noise is not None else int(horizon)
        mean = np.zeros((H, nv))
    mean = np.ascontiguousarray(mean, np.float64)
    H = mean.shape[0]
    K = noise.shape[0] if noise is not None else 1
    assert K % len(models) == 0
    costs = np.zeros((K, H)); actions = np.zeros((K, H, nv))
    qv = np.zeros((K, H, 2 * nv)) if want_traj else None
    nobs = np.zeros((K, H, 2 * nv + 6)) if want_obs else None
    ncon = np.zeros(K, np.int32)
    arr = (C.c_void_p * len(models))(*[m.h for m in models])
    qpos = np.ascontiguousarray(qpos, np.float64); qvel = np.ascontiguousarray(qvel, np.float64)
    target = np.ascontiguousarray(target, np.float64)
    nul = C.POINTER(C.c_double)()
    lib().ora_rollout_cl(arr, C.c_int(len(models)), _p(qpos), _p(qvel), _p(target), C.c_int(K), C.c_int(H),
                         _p(mean), _p(noise) if noise is not None else nul, _p(costs), _p(actions),
                         _p(qv) if qv is not None else nul, _p(nobs) if nobs is not None else nul,
                         _p(ncon, C.c_int), C.c_int(nthreads), _p(policy_w) if policy_w is not None else nul)
    out = dict(costs=costs, actions=actions, ncon=ncon)
    if qv is not None:
        out["qv"] = qv
    if nobs is not None:
        out["next_observations"] = nobs
    return out


def pendulum_rollout(th0, thdot0, mean, noise):
    mean = np.ascontiguousarray(mean, np.float64).reshape(-1)
    H = mean.shape[0]
    noise = np.ascontiguousarray(noise, np.float64).reshape(-1, H)
    K = noise.shape[0]
    costs = np.zeros((K, H)); actions = np.zeros((K, H)); states = np.zeros((K, H, 2))
    lib().ora_pendulum_rollout(C.c_double(th0), C.c_double(thdot0), C.c_int(K), C.c_int(H), _p(mean),
                               _p(noise), _p(costs), _p(actions), _p(states))
    return dict(costs=costs, actions=actions.reshape(K, H, 1), states=states)
