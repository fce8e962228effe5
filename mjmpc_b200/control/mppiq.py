"""MPPI with Q-function estimates -- the reference's ``MPPIQ`` (``mjmpc/control/mppiq.py:19-160``):
same constructor arguments; the TD(lambda) returns (``calculate_returns``, :104-126), the
exponential-utility weights (:91-102) and the weighted mean (:73-89) run on the GPU.
"""
from __future__ import annotations

import numpy as np

from .. import _lib
from .olgaussian_mpc import OLGaussianMPC


class MPPIQ(OLGaussianMPC):
    def __init__(self, d_state, d_obs, d_action, horizon, init_cov, base_action, beta, num_particles, step_size,
                 alpha, gamma, n_iters, td_lam, action_lows, action_highs, time_based_weights=True,
                 set_sim_state_fn=None, get_sim_state_fn=None, sim_step_fn=None, sim_reset_fn=None,
                 rollout_fn=None, sample_mode='mean', batch_size=1, filter_coeffs=[1., 0., 0.], seed=0,
                 device=None, shard=None):
        # get_sim_state_fn / sim_step_fn / sim_reset_fn are accepted and ignored, like mppiq.py:37-39
        super(MPPIQ, self).__init__(d_state, d_obs, d_action, action_lows, action_highs, horizon, init_cov,
                                    np.zeros(shape=(horizon, d_action)), base_action, num_particles, gamma,
                                    n_iters, step_size, filter_coeffs, set_sim_state_fn, rollout_fn, 'diagonal',
                                    sample_mode, batch_size, seed, False, device, shard)
        if self.batch_size > 1:
            raise NotImplementedError("batched instances are implemented for MPPI only")
        self.beta = beta
        self.td_lam = td_lam
        self.alpha = alpha  # 0 means control cost is on, 1 means off
        self.time_based_weights = time_based_weights

    def td_weight_seq(self):
        """mppiq.py:116-119: cumprod([1, gamma*td_lam, ...]) over the H-1 TD errors."""
        if self.horizon == 1:
            return np.array([1.0])
        return np.cumprod([1.0] + [self.gamma * self.td_lam] * (self.horizon - 2))

    def _qvals(self, trajectories):
        q = trajectories.get("qvals") if hasattr(trajectories, "get") else None
        if q is None:
            return None
        q = self._to_device(q)
        if tuple(q.shape) != (self.local_particles, self.horizon):
            raise ValueError("rollout_fn returned qvals %s for K=%d H=%d" % (tuple(q.shape), self.local_particles,
                                                                             self.horizon))
        return q

    def _update_distribution(self, trajectories):
        """mppiq.py:73-102: w = softmax(-q_hat/beta, axis=0) with q_hat the TD(lambda) return of
        costs + beta*control cost; mean <- (1-step)*mean + step * sum_k w_k a_k."""
        costs, actions = self._traj(trajectories)
        self._softmax_update(costs, actions, self.beta, control_cost=(self.alpha != 1),
                             time_based=bool(self.time_based_weights),
                             td=(self.td_lam, self.gamma, self.td_weight_seq(), self._qvals(trajectories)))

    def _calc_val(self, trajectories):
        """mppiq.py:138-160: -beta * logsumexp(-q_hat[:,0]/beta, b=1/K)."""
        costs, actions = self._traj(trajectories)
        stats = self._softmax_update(costs, actions, self.beta, control_cost=(self.alpha != 1), apply=False,
                                     td=(self.td_lam, self.gamma, self.td_weight_seq(), self._qvals(trajectories)))
        return float(stats[0].item())
