"""Model Predictive Path Integral controller -- the reference's ``MPPI``
(``mjmpc/control/mppi.py:15-131``): same constructor arguments, GPU update.
"""
from __future__ import annotations

import ctypes

import numpy as np

from .. import _lib
from .olgaussian_mpc import OLGaussianMPC


class MPPI(OLGaussianMPC):
    def __init__(self, d_state, d_obs, d_action, horizon, init_cov, base_action, lam, num_particles,
                 step_size, alpha, gamma, n_iters, action_lows, action_highs, time_based_weights=False,
                 set_sim_state_fn=None, get_sim_state_fn=None, sim_step_fn=None, sim_reset_fn=None,
                 rollout_fn=None, sample_mode='mean', batch_size=1, filter_coeffs=[1., 0., 0.], seed=0,
                 use_zero_control_seq=False, device=None, shard=None):
        # get_sim_state_fn / sim_step_fn / sim_reset_fn are accepted and ignored, like mppi.py:34-36
        super(MPPI, self).__init__(d_state, d_obs, d_action, action_lows, action_highs, horizon, init_cov,
                                   np.zeros(shape=(horizon, d_action)), base_action, num_particles, gamma,
                                   n_iters, step_size, filter_coeffs, set_sim_state_fn, rollout_fn, 'diagonal',
                                   sample_mode, batch_size, seed, use_zero_control_seq, device, shard)
        self.lam = lam
        self.alpha = alpha  # 0 means control cost is on, 1 means off
        self.time_based_weights = time_based_weights

    def _update_distribution(self, trajectories):
        """mppi.py:69-97: w = softmax(-(cost-to-go + lam*control cost)/lam);
        mean <- (1-step)*mean + step * sum_k w_k a_k."""
        costs, actions = self._traj(trajectories)
        if self.batch_size > 1:
            self._batched_update(costs, actions, apply=True)
            return
        self._softmax_update(costs, actions, self.lam, control_cost=(self.alpha != 1),
                             time_based=bool(self.time_based_weights))

    def _softmax_spec(self):
        return dict(lam=float(self.lam), control_cost=(self.alpha != 1), time_based=bool(self.time_based_weights),
                    cov_mode=_lib.COV_NONE, cov_shift_beta=0.0)

    def _batched_update(self, costs, actions, apply):
        """batch_size independent instances: one thread block per instance, no cross-instance reduction."""
        if self.time_based_weights:
            raise NotImplementedError("time_based_weights is not available for batched instances")
        a = _lib.MppiBatchedArgs()
        a.n_ctrl, a.K, a.H, a.d = self.batch_size, self.num_particles, self.horizon, self.d_action
        a.costs = costs.data_ptr(); a.costs_sk, a.costs_st = costs.stride()
        a.actions = actions.data_ptr(); a.act_sk, a.act_st, a.act_sj = actions.stride()
        mean = self._mean if apply else self._mean.clone()
        a.mean, a.cov = mean.data_ptr(), self._cov.data_ptr()
        g = np.ascontiguousarray(self.gamma_seq.reshape(-1))
        a.gamma_seq = g.ctypes.data
        a.lam, a.step_size, a.control_cost = float(self.lam), float(self.step_size), int(self.alpha != 1)
        value = self._buf("batched_value", (self.batch_size,))
        a.value = value.data_ptr()
        _lib.check(_lib.lib().mjb_mppi_update_batched(ctypes.byref(a), _lib.stream_ptr()))
        return value

    def _calc_val(self, trajectories):
        """mppi.py:113-131: -lam * logsumexp(-total/lam, b=1/K)."""
        costs, actions = self._traj(trajectories)
        if self.batch_size > 1:
            return self._batched_update(costs, actions, apply=False).cpu().numpy()
        stats = self._softmax_update(costs, actions, self.lam, control_cost=(self.alpha != 1), apply=False)
        return float(stats[0].item())
