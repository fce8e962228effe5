"""GPU rollout backend for SimplePendulum-v0 (reference env: ``mjmpc/envs/basic/pendulum.py``),
same adaptor surface as :class:`GpuReacherVecEnv` (set_env_state / rollout / rollout_fn)."""
from __future__ import annotations

import ctypes as C
import time

import numpy as np
import torch

from .. import _lib


class GpuPendulumVecEnv:
    d_action = 1
    d_obs = 3
    d_state = 2

    def __init__(self, device: int = 0):
        if not torch.cuda.is_available():
            raise _lib.MjbError("GpuPendulumVecEnv needs a CUDA device (there is no CPU fallback)")
        _lib.lib()
        self.device = torch.device("cuda", device)
        self._state = torch.zeros(1, 2, dtype=torch.float64, device=self.device)
        self.action_lows = np.array([-2.0])
        self.action_highs = np.array([2.0])

    def set_env_state(self, state_dicts):
        """pendulum.py:96-97: {'state': (theta, thetadot)}; a list gives one state per batched controller."""
        if not isinstance(state_dicts, (list, tuple)):
            state_dicts = [state_dicts]
        rows = np.stack([np.asarray(s["state"] if isinstance(s, dict) else s, float).reshape(2) for s in state_dicts])
        self._state = torch.from_numpy(rows).to(self.device)

    def reset(self):
        pass

    def close(self):
        pass

    def rollout_device(self, num_particles, horizon, mean, noise, want_states=False):
        K, H = int(num_particles), int(horizon)
        n_ctrl = self._state.shape[0]
        mean = mean.reshape(n_ctrl, H).contiguous()
        costs = torch.empty((H, K), dtype=torch.float64, device=self.device).t()
        actions = torch.empty((H, 1, K), dtype=torch.float64, device=self.device).permute(2, 0, 1)
        a = _lib.PendulumArgs()
        a.K, a.H, a.particles_per_ctrl = K, H, K // n_ctrl
        a.state, a.mean = self._state.data_ptr(), mean.data_ptr()
        if noise is not None:
            if tuple(noise.shape) != (K, H, 1):
                raise ValueError("noise must have shape (K,H,1)")
            a.noise = noise.data_ptr()
            a.noise_sk, a.noise_st = noise.stride(0), noise.stride(1)
        a.costs = costs.data_ptr(); a.costs_sk, a.costs_st = costs.stride()
        a.actions = actions.data_ptr(); a.act_sk, a.act_st = actions.stride(0), actions.stride(1)
        out = dict(costs=costs, actions=actions)
        if want_states:
            out["states"] = torch.empty((K, H, 2), dtype=torch.float64, device=self.device)
            a.states_out = out["states"].data_ptr()
        _lib.check(_lib.lib().mjb_rollout_pendulum(C.byref(a), _lib.stream_ptr()))
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
