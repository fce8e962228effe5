"""ctypes front-end of ``oracle/tree_step.c`` (TEST INFRASTRUCTURE -- see oracle/__init__.py).

PARITY UNPINNED AGAINST MuJoCo ITSELF (pinned against the independent restatement oracle/tree_ref.py and closed
forms: tests/test_tree_oracle.py); see the header of tree_step.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle_tree.so")
    src = os.path.join(_HERE, "tree_step.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle_tree.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.tree_model_create.restype = C.c_void_p
        L.tree_rollout.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


class TreeOracle:
    """One ``tree_model`` built from a compiled ``mjmpc_b200.envs.mjcf_tree.TreeModel`` (arrays only)."""

    def __init__(self, model, solref_to_kb):
        L = lib()
        f = lambda a: np.ascontiguousarray(a, np.float64)
        i = lambda a: np.ascontiguousarray(a, np.int32)
        kb = np.array([solref_to_kb(model.jnt_solref[j], model.jnt_solimp[j], model.timestep) for j in range(model.nv)])
        ctrl = np.where(model.act_ctrllimited[:, None], model.act_ctrlrange, np.array([-np.inf, np.inf])[None, :])
        k = self._keep = [i(model.body_parent), f(model.body_pos), f(model.body_mat), f(model.body_mass), f(model.body_ipos),
                          f(model.body_imat), f(model.body_inertia), i(model.jnt_type), i(model.jnt_body), f(model.jnt_pos),
                          f(model.jnt_axis), i(model.jnt_limited), f(model.jnt_range), f(model.jnt_damping),
                          f(model.jnt_armature), f(model.jnt_stiffness), f(model.jnt_springref), f(model.dof_invweight0),
                          f(kb[:, 0]), f(kb[:, 1]), f(model.jnt_solimp), i(model.act_dof), f(model.act_gear), f(ctrl),
                          f(model.gravity)]
        ci = C.c_int
        self.nv, self.nu, self.nb = model.nv, model.nu, model.nb
        self.h = C.c_void_p(L.tree_model_create(
            ci(model.nb), ci(model.nv), ci(model.nu), _p(k[0], ci), _p(k[1]), _p(k[2]), _p(k[3]), _p(k[4]), _p(k[5]),
            _p(k[6]), _p(k[7], ci), _p(k[8], ci), _p(k[9]), _p(k[10]), _p(k[11], ci), _p(k[12]), _p(k[13]), _p(k[14]),
            _p(k[15]), _p(k[16]), _p(k[17]), _p(k[18]), _p(k[19]), _p(k[20]), _p(k[21], ci), _p(k[22]), _p(k[23]),
            C.c_double(model.timestep), _p(k[24]), C.c_double(model.density), C.c_double(model.viscosity)))
        if not self.h:
            raise ValueError("model too large for the oracle")
        cs = getattr(model, "contacts", [])
        if cs:
            kb = np.array([solref_to_kb(c["solref"], c["solimp"], model.timestep) for c in cs])
            arr = [i([1 if c["kind"] == "plane" else 0 for c in cs]), i([c["body1"] for c in cs]), i([c["body2"] for c in cs]),
                   f([c["a0"] for c in cs]), f([c["a1"] for c in cs]), f([c["ra"] for c in cs]), f([c["b0"] for c in cs]),
                   f([c["b1"] for c in cs]), f([c["rb"] for c in cs]), f([c["mu"] for c in cs]), f(kb[:, 0]), f(kb[:, 1]),
                   f([c["solimp"] for c in cs]), f([c["invweight"] for c in cs])]
            self._keep += arr
            if L.tree_model_set_contacts(self.h, ci(len(cs)), _p(arr[0], ci), _p(arr[1], ci), _p(arr[2], ci), *[_p(a) for a in arr[3:]]):
                raise ValueError("too many contact candidates for the oracle")

    def __del__(self):
        if getattr(self, "h", None):
            lib().tree_model_free(self.h)
            self.h = None

    def substep(self, q, v, u):
        """One mj_step; returns dict(q, v, M, bias, passive, actuation, constraint, qacc, nefc)."""
        nv = self.nv
        q, v = np.array(q, np.float64), np.array(v, np.float64)
        u = np.ascontiguousarray(u, np.float64)
        M, out = np.zeros((nv, nv)), [np.zeros(nv) for _ in range(5)]
        nefc = C.c_int(0)
        lib().tree_substep_debug(self.h, _p(q), _p(v), _p(u), _p(M), *[_p(o) for o in out], C.byref(nefc))
        return dict(q=q, v=v, M=M, bias=out[0], passive=out[1], actuation=out[2], constraint=out[3], qacc=out[4],
                    nefc=nefc.value)

    def rollout(self, state0, mean, noise, frame_skip, fwd_dof=0, w_fwd=1.0, w_ctrl=1e-4, nthreads=8):
        """state0 (2 nv,) or (K, 2 nv); mean (H, nu); noise (K, H, nu) -> costs (K, H), actions, states (K, H, 2 nv)."""
        noise = np.ascontiguousarray(noise, np.float64)
        K, H, nu = noise.shape
        mean = np.ascontiguousarray(mean, np.float64)
        s0 = np.ascontiguousarray(state0, np.float64)
        stride = 0 if s0.ndim == 1 else 2 * self.nv
        costs, actions, states = np.zeros((K, H)), np.zeros((K, H, nu)), np.zeros((K, H, 2 * self.nv))
        nefc = lib().tree_rollout(self.h, C.c_int(K), C.c_int(H), C.c_int(frame_skip), C.c_int(fwd_dof),
                                  C.c_double(w_fwd), C.c_double(w_ctrl), _p(s0), C.c_int(stride), _p(mean), _p(noise),
                                  _p(costs), _p(actions), _p(states), C.c_int(nthreads))
        return dict(costs=costs, actions=actions, states=states, nefc=nefc)
