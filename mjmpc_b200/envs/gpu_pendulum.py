"""GPU rollout backend for SimplePendulum-v0 (reference env: ``mjmpc/envs/basic/pendulum.py``),
same adaptor surface as :class:`GpuReacherVecEnv` (set_env_state / rollout / rollout_fn)."""
from __future__ import annotations

import ctypes as C

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


class GpuPendulumEnv:
    """The PLANT for ``SimplePendulum-v0`` (``mjmpc/envs/__init__.py:5-9``, ``pendulum.py:7-103``): one pendulum
    advanced by the same kernel the planner rolls out (K = 1, H = 1), with the reference env's ``reset`` /
    ``step`` / ``get_env_state`` / ``set_env_state`` / ``get_obs`` / ``evaluate_success`` surface."""
    _max_episode_steps = 200
    d_obs, d_state, d_action = GpuPendulumVecEnv.d_obs, GpuPendulumVecEnv.d_state, GpuPendulumVecEnv.d_action

    def __init__(self, device: int = 0, seed=None):
        self.sim = GpuPendulumVecEnv(device=device)
        self.action_lows, self.action_highs = self.sim.action_lows, self.sim.action_highs
        self.np_random = np.random.RandomState(seed)
        self.state = np.zeros(2)

    def reset(self, seed=None):
        """pendulum.py:52-56: theta ~ U(-pi, pi), thetadot ~ U(-1, 1) from the env's generator."""
        if seed is not None:
            self.np_random = np.random.RandomState(seed)
        high = np.array([np.pi, 1])
        self.state = self.np_random.uniform(low=-high, high=high)
        return self.get_obs()

    def get_obs(self):
        theta, thetadot = self.state
        return np.array([np.cos(theta), np.sin(theta), thetadot])

    def get_env_state(self):
        return {'state': self.state.copy()}

    def set_env_state(self, state_dict):
        self.state = np.asarray(state_dict['state'], float).reshape(2).copy()

    def step(self, u):
        self.sim.set_env_state(self.get_env_state())
        mean = torch.as_tensor(np.asarray(u, np.float64).reshape(1, 1), device=self.sim.device)
        out = self.sim.rollout_device(1, 1, mean, None, want_states=True)
        self.state = out["states"][0, 0].cpu().numpy().copy()
        return self.get_obs(), -float(out["costs"][0, 0].item()), False, {}

    def evaluate_success(self, trajectories):
        return 0.0                      # pendulum.py:102-103

    def close(self):
        self.sim.close()
