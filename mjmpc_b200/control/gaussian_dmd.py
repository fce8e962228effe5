"""DMD-MPC with Gaussian sampling, exponential utility and covariance adaptation -- the
reference's ``DMDMPC`` (``mjmpc/control/gaussian_dmd.py:16-139``), GPU update.

The full-covariance update is a (d x K*H)(K*H x d) contraction with d = 7: 1.75 FLOP per byte
of the action tensor, i.e. bound by HBM, so it is accumulated on the CUDA cores in the same
pass that forms the weighted mean (tensor cores would idle).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from .olgaussian_mpc import OLGaussianMPC


class DMDMPC(OLGaussianMPC):
    def __init__(self, d_state, d_obs, d_action, horizon, init_cov, beta, base_action, lam, num_particles,
                 step_size, gamma, n_iters, action_lows, action_highs, set_sim_state_fn=None,
                 rollout_fn=None, update_cov=False, cov_type='diagonal', sample_mode='mean', batch_size=1,
                 filter_coeffs=[1., 0., 0.], seed=0, device=None, shard=None):
        super(DMDMPC, self).__init__(d_state, d_obs, d_action, action_lows, action_highs, horizon, init_cov,
                                     np.zeros(shape=(horizon, d_action)), base_action, num_particles, gamma,
                                     n_iters, step_size, filter_coeffs, set_sim_state_fn, rollout_fn, cov_type,
                                     sample_mode, batch_size, seed, False, device, shard)
        self.lam = lam
        self.beta = beta
        self.update_cov = update_cov

    def _cov_mode(self):
        if not self.update_cov:
            return _lib.COV_NONE
        if self.cov_type == 'diagonal':
            return _lib.COV_DIAG
        if self.cov_type == 'full':
            return _lib.COV_FULL
        raise ValueError('Unidentified covariance type in update_distribution')

    def _update_distribution(self, trajectories):
        """gaussian_dmd.py:65-91."""
        costs, actions = self._traj(trajectories)
        if self.batch_size > 1:
            self._batched_update(costs, actions, apply=True)
            return
        self._softmax_update(costs, actions, self.lam, cov_mode=self._cov_mode())

    def _batched_update(self, costs, actions, apply):
        """batch_size independent instances with a fixed covariance: the exponential-utility weighted mean is
        MPPI's without the control cost (gaussian_dmd.py:94-104 == mppi.py:84-97 at alpha = 1), one thread block per
        instance (mjb_mppi_update_batched)."""
        if self.update_cov:
            raise NotImplementedError("batched DMD-MPC instances are implemented for update_cov=False")
        a = _lib.MppiBatchedArgs()
        a.n_ctrl, a.K, a.H, a.d = self.batch_size, self.num_particles, self.horizon, self.d_action
        a.costs = costs.data_ptr(); a.costs_sk, a.costs_st = costs.stride()
        a.actions = actions.data_ptr(); a.act_sk, a.act_st, a.act_sj = actions.stride()
        mean = self._mean if apply else self._mean.clone()
        a.mean, a.cov = mean.data_ptr(), self._cov.data_ptr()
        g = np.ascontiguousarray(self.gamma_seq.reshape(-1))
        a.gamma_seq = g.ctypes.data
        a.lam, a.step_size, a.control_cost = float(self.lam), float(self.step_size), 0
        value = self._buf("batched_value", (self.batch_size,))
        a.value = value.data_ptr()
        _lib.check(_lib.lib().mjb_mppi_update_batched(C.byref(a), _lib.stream_ptr()))
        return value

    def _softmax_spec(self):
        return dict(lam=float(self.lam), control_cost=False, time_based=False, cov_mode=self._cov_mode(),
                    cov_shift_beta=float(self.beta) if self.update_cov else 0.0)

    def _shift(self):
        """gaussian_dmd.py:106-113: shift the mean; grow the covariance by beta*I if it is adapted."""
        super()._shift()
        if self.update_cov:
            _lib.check(_lib.lib().mjb_cov_add_diag(_lib.ptr(self._cov), C.c_int(self.d_action),
                                                   C.c_double(self.beta), None, _lib.stream_ptr()))

    def _calc_val(self, trajectories):
        """gaussian_dmd.py:126-139."""
        costs, actions = self._traj(trajectories)
        if self.batch_size > 1:
            return self._batched_update(costs, actions, apply=False).cpu().numpy()
        stats = self._softmax_update(costs, actions, self.lam, apply=False)
        return float(stats[0].item())
