"""MPC with closed-loop (linear-)Gaussian policies -- the reference's ``CLGaussianMPC``
(``mjmpc/control/clgaussian_mpc.py:10-145``): the distribution is over the weights of a linear policy
``u_t = W' [obs_t; 1] + noise_t`` evaluated INSIDE the rollout (``mode="closed_loop_linear"``,
``mjmpc/envs/gym_env_wrapper.py:129-136``) -- here inside the rollout kernel K1, one policy evaluation per
particle and step in registers.

Like the reference class it has no update rule of its own (``_update_distribution`` stays abstract; the reference's
only concrete subclass is the policy-gradient ``Reinforce``, outside the sampling-MPC path): subclasses provide one.
"""
from __future__ import annotations

import copy

import numpy as np
import torch

from .. import _lib
from ..utils import control_utils
from ..utils.control_utils import generate_noise
from .controller import Controller


class CLGaussianMPC(Controller):
    def __init__(self, d_state, d_obs, d_action, action_lows, action_highs, horizon, init_cov, init_mean, num_particles,
                 gamma, n_iters, filter_coeffs, set_sim_state_fn=None, rollout_fn=None, cov_type='diagonal',
                 sample_mode='mean', batch_size=1, seed=0, device=None, shard=None):
        """Parameters as in the reference (clgaussian_mpc.py:11-55); ``init_mean`` is the (d_obs + 1, d_action)
        weight matrix of the linear policy (last row: bias)."""
        super(CLGaussianMPC, self).__init__(d_state, d_obs, d_action, action_lows, action_highs, horizon, gamma, n_iters,
                                            set_sim_state_fn, rollout_fn, sample_mode, batch_size, seed, device, shard)
        if self.shard.world_size != 1 or batch_size != 1:
            raise NotImplementedError("closed-loop policies run unsharded, one instance")
        self.init_cov = np.array([init_cov] * self.d_action)
        self.init_mean = np.array(init_mean, dtype=np.float64).copy()
        if self.init_mean.shape != (self.d_obs + 1, self.d_action):
            raise ValueError("init_mean must be the (d_obs + 1, d_action) weight matrix of the linear policy")
        self.num_particles = int(num_particles)
        self.cov_type = cov_type
        self.filter_coeffs = filter_coeffs
        self._weights = self._to_device(self.init_mean).contiguous().clone()
        self._cov = self._to_device(np.diag(self.init_cov)).contiguous()
        self.curr_obs = None

    # numpy views, as the reference's attributes
    @property
    def mean_weights(self):
        return self._weights.cpu().numpy()

    @mean_weights.setter
    def mean_weights(self, value):
        self._weights = self._to_device(value).reshape(self.d_obs + 1, self.d_action).contiguous().clone()

    @property
    def cov_action(self):
        return self._cov.cpu().numpy()

    @cov_action.setter
    def cov_action(self, value):
        self._cov = self._to_device(value).reshape(self.d_action, self.d_action).contiguous()

    # clgaussian_mpc.py:62-73
    def _get_next_action(self, state, mode='mean'):
        mean_action = self.mean_weights.T @ np.append(self.curr_obs, 1.0)
        if mode == 'mean':
            next_action = mean_action.copy()
        elif mode == 'sample':
            delta = generate_noise(self._cov, self.filter_coeffs, shape=(1, 1), base_seed=self.seed_val,
                                   step=123 * self.num_steps, stream_id=control_utils.NOISE_STREAM_ACTION, device=self.device)
            next_action = mean_action.copy() + delta.reshape(self.d_action).cpu().numpy().copy()
        else:
            raise ValueError('Unidentified sampling mode in get_next_action')
        return next_action

    # clgaussian_mpc.py:84-89
    def sample_noise(self):
        return generate_noise(self._cov, self.filter_coeffs, shape=(self.num_particles, self.horizon),
                              base_seed=self.seed_val, step=self.num_steps, device=self.device)

    # clgaussian_mpc.py:91-117
    def generate_rollouts(self, state):
        self._set_sim_state_fn(copy.deepcopy(state))
        delta = self.sample_noise()
        trajectories = self._rollout_fn(self.num_particles, self.horizon, self._weights, delta, mode="closed_loop_linear")
        obs0 = trajectories["observations"][0, 0]
        self.curr_obs = obs0.cpu().numpy() if isinstance(obs0, torch.Tensor) else np.asarray(obs0)
        return trajectories

    # clgaussian_mpc.py:119-134: the policy is not shifted
    def _shift(self):
        pass

    # clgaussian_mpc.py:136-143
    def reset(self, seed=None):
        if seed is not None:
            self.seed_val = self.seed(seed)
        self.num_steps = 0
        self.mean_weights = self.init_mean.copy()
        self.cov_action = np.diag(self.init_cov)
        self.gamma_seq = np.cumprod([1.0] + [self.gamma] * (self.horizon - 1)).reshape(1, self.horizon)
        self.converged = False

    def _calc_val(self, trajectories):
        raise NotImplementedError("_calc_val not implemented")

    def _first_action(self):
        raise _lib.MjbError("closed-loop controllers return their action through _get_next_action")
