"""GPU rollout backend for the reference's linear-quadratic toy env (``mjmpc/envs/basic/lqr.py``: cost
``x'Qx + u'Ru`` on the pre-step state, ``x <- Ax + Bu``), same adaptor surface as
:class:`GpuReacherVecEnv` / :class:`GpuPendulumVecEnv` (set_env_state / rollout_device / rollout_fn)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib

MAXN, MAXD = 8, 8          # MJB_LQR_MAXN / MJB_LQR_MAXD of include/mjmpc_b200.h


class GpuLQRVecEnv:
    def __init__(self, A, B, Q, R, device: int = 0):
        """A (n,n), B (n,d), Q (n,n), R (d,d) as in ``LQREnv.__init__`` (lqr.py:13-19)."""
        if not torch.cuda.is_available():
            raise _lib.MjbError("GpuLQRVecEnv needs a CUDA device (there is no CPU fallback)")
        _lib.lib()
        A, B, Q, R = [np.atleast_2d(np.asarray(m, np.float64)) for m in (A, B, Q, R)]
        n, d = A.shape[0], B.shape[-1]
        if A.shape != (n, n) or B.shape != (n, d) or Q.shape != (n, n) or R.shape != (d, d):
            raise ValueError("inconsistent LQR matrices: A %s B %s Q %s R %s" % (A.shape, B.shape, Q.shape, R.shape))
        if n > MAXN or d > MAXD:
            raise ValueError("d_state=%d / d_action=%d exceed the kernel's limits %d / %d" % (n, d, MAXN, MAXD))
        self.d_state = self.d_obs = n
        self.d_action = d
        self.device = torch.device("cuda", device)
        self._mats = [torch.from_numpy(np.ascontiguousarray(m)).to(self.device) for m in (A, B, Q, R)]
        self._state = torch.zeros(1, n, dtype=torch.float64, device=self.device)
        self.action_lows = np.full(d, -np.inf)        # lqr.py:22: an unbounded action space
        self.action_highs = np.full(d, np.inf)

    def set_env_state(self, state_dicts):
        """lqr.py:79-80: {'state': (n,1) column}; a list gives one state per batched controller."""
        if not isinstance(state_dicts, (list, tuple)):
            state_dicts = [state_dicts]
        rows = np.stack([np.asarray(s["state"] if isinstance(s, dict) else s, float).reshape(self.d_state) for s in state_dicts])
        self._state = torch.from_numpy(rows).to(self.device)

    def reset(self):
        pass

    def close(self):
        pass

    def rollout_device(self, num_particles, horizon, mean, noise, want_states=False):
        K, H, n, d = int(num_particles), int(horizon), self.d_state, self.d_action
        n_ctrl = self._state.shape[0]
        mean = mean.reshape(n_ctrl, H, d).contiguous()
        costs = torch.empty((H, K), dtype=torch.float64, device=self.device).t()
        actions = torch.empty((H, d, K), dtype=torch.float64, device=self.device).permute(2, 0, 1)
        a = _lib.LqrArgs()
        a.K, a.H, a.n, a.d, a.particles_per_ctrl = K, H, n, d, K // n_ctrl
        a.A, a.B, a.Q, a.R = [m.data_ptr() for m in self._mats]
        a.state, a.mean = self._state.data_ptr(), mean.data_ptr()
        if noise is not None:
            if tuple(noise.shape) != (K, H, d):
                raise ValueError("noise must have shape (K,H,%d)" % d)
            a.noise = noise.data_ptr()
            a.noise_sk, a.noise_st, a.noise_sj = noise.stride()
        a.costs = costs.data_ptr(); a.costs_sk, a.costs_st = costs.stride()
        a.actions = actions.data_ptr(); a.act_sk, a.act_st, a.act_sj = actions.stride()
        out = dict(costs=costs, actions=actions)
        if want_states:
            out["states"] = torch.empty((K, H, n), dtype=torch.float64, device=self.device)
            a.states_out = out["states"].data_ptr()
        _lib.check(_lib.lib().mjb_rollout_lqr(C.byref(a), _lib.stream_ptr()))
        return out

    def _to_device(self, x):
        if x is None or isinstance(x, torch.Tensor):
            return x
        return torch.from_numpy(np.ascontiguousarray(x, np.float64)).to(self.device)

    @property
    def rollout_fn(self):
        def fn(num_particles, horizon, mean, noise, mode="open_loop"):
            if mode != "open_loop":
                raise NotImplementedError("only mode='open_loop' runs on the GPU rollout")
            host = not isinstance(noise, torch.Tensor) and not isinstance(mean, torch.Tensor)
            out = self.rollout_device(num_particles, horizon, self._to_device(mean), self._to_device(noise))
            if host:
                return {k: np.ascontiguousarray(v.cpu().numpy()) for k, v in out.items()}
            return out
        return fn
