"""Cross Entropy Method for MPC -- the reference's ``CEM`` (``mjmpc/control/cem.py:15-112``),
GPU update: radix top-k elite selection (ties -> lower index), then elite mean and pooled
covariance in two ordered passes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib
from .olgaussian_mpc import OLGaussianMPC


class CEM(OLGaussianMPC):
    def __init__(self, d_state, d_obs, d_action, horizon, init_cov, base_action, elite_frac, num_particles,
                 step_size, gamma, n_iters, action_lows, action_highs, set_sim_state_fn=None,
                 rollout_fn=None, beta=0.0, cov_type='diagonal', sample_mode='mean', batch_size=1,
                 filter_coeffs=[1., 0., 0.], seed=0, device=None, shard=None):
        super(CEM, self).__init__(d_state, d_obs, d_action, action_lows, action_highs, horizon, init_cov,
                                  np.zeros(shape=(horizon, d_action)), base_action, num_particles, gamma,
                                  n_iters, step_size, filter_coeffs, set_sim_state_fn, rollout_fn, cov_type,
                                  sample_mode, batch_size, seed, False, device, shard)
        self.elite_frac = elite_frac
        self.beta = beta
        self.num_elite = int(self.num_particles * self.elite_frac)
        self.elite_ids = None      # device int64 (num_elite,), ascending global particle index

    def _on_params_changed(self, reshape):
        self.num_elite = int(self.num_particles * self.elite_frac)
        super()._on_params_changed(reshape)

    def _select(self, ctg0_local):
        """Global elite flags/ids from every shard's cost-to-go (all-gathered: K doubles)."""
        L = _lib.lib()
        allc = self.shard.all_gather(ctg0_local.contiguous()).reshape(-1)
        K = self.num_particles
        flags = self._buf("elite_flags", (K,), torch.uint8)
        ids = self._buf("elite_ids", (self.num_elite,), torch.int64)
        _lib.check(L.mjb_select_elites(_lib.ptr(allc), _lib.c_ll(K), _lib.c_ll(self.num_elite), _lib.ptr(flags),
                                       _lib.ptr(ids), None, _lib.stream_ptr()))
        return flags, ids

    def _update_distribution(self, trajectories):
        """cem.py:65-86."""
        if self.cov_type not in ('diagonal', 'full'):
            raise ValueError('Unidentified covariance type in update_distribution')
        L = _lib.lib()
        costs, actions = self._traj(trajectories)
        if self.batch_size > 1:
            self._batched_update(costs, actions, apply=True)
            return
        k0, kl = self.shard.local_range(self.num_particles)
        H, d = self.horizon, self.d_action
        flags, ids = self._select(self._ctg0(costs))
        self.elite_ids = ids
        a = _lib.EliteArgs()
        a.K, a.H, a.d = kl, H, d
        a.flags = flags[k0:k0 + kl].data_ptr()
        a.actions = actions.data_ptr(); a.act_sk, a.act_st, a.act_sj = actions.stride()
        a.mean = self._mean.data_ptr()
        scratch = self._buf("elite_scratch", (int(L.mjb_elite_scratch_doubles(kl, H, d)),))
        p1 = self._buf("elite_p1", (1 + H * d + d,))
        p2 = self._buf("elite_p2", (d * (d + 1) // 2,))
        mu = self._buf("elite_mu", (d,))
        a.scratch, a.partial = scratch.data_ptr(), p1.data_ptr()
        _lib.check(L.mjb_elite_moments1(C.byref(a), _lib.stream_ptr()))
        all1 = self.shard.all_gather(p1)
        c = _lib.EliteCombineArgs()
        c.H, c.d, c.n_shards, c.full_cov = H, d, self.shard.world_size, int(self.cov_type == 'full')
        c.partial1, c.partial2, c.step_size = all1.data_ptr(), None, float(self.step_size)
        c.mu = mu.data_ptr()
        _lib.check(L.mjb_elite_combine(C.byref(c), _lib.stream_ptr()))          # pooled mean of elite deltas
        a.mu, a.partial = mu.data_ptr(), p2.data_ptr()
        _lib.check(L.mjb_elite_moments2(C.byref(a), _lib.stream_ptr()))
        all2 = self.shard.all_gather(p2)
        c.partial2, c.mean, c.cov = all2.data_ptr(), self._mean.data_ptr(), self._cov.data_ptr()
        _lib.check(L.mjb_elite_combine(C.byref(c), _lib.stream_ptr()))

    _per_instance_cov = True          # batched instances adapt their own covariance

    def _batched_update(self, costs, actions, apply):
        """batch_size independent instances: one thread block per instance does cost-to-go, elite selection and
        both moments (cem.py:65-86); every instance keeps its own covariance.  elite_ids: (batch_size, num_elite)."""
        a = _lib.InstancesArgs()
        a.n_ctrl, a.K, a.H, a.d = self.batch_size, self.num_particles, self.horizon, self.d_action
        a.mode, a.apply = (_lib.INST_CEM_FULL if self.cov_type == 'full' else _lib.INST_CEM_DIAG), int(apply)
        a.num_elite = self.num_elite
        a.costs = costs.data_ptr(); a.costs_sk, a.costs_st = costs.stride()
        a.actions = actions.data_ptr(); a.act_sk, a.act_st, a.act_sj = actions.stride()
        a.mean, a.cov = self._mean.data_ptr(), self._cov.data_ptr()
        g = np.ascontiguousarray(self.gamma_seq.reshape(-1))
        a.gamma_seq = g.ctypes.data
        a.step_size = float(self.step_size)
        ids = self._buf("elite_ids_b", (self.batch_size, self.num_elite), torch.int64)
        value = self._buf("batched_value", (self.batch_size,))
        a.ids, a.value = ids.data_ptr(), value.data_ptr()
        _lib.check(_lib.lib().mjb_instances_update_batched(C.byref(a), _lib.stream_ptr()))
        self.elite_ids = ids
        return value

    def _shift(self):
        """cem.py:89-95: shift the mean and grow the covariance by beta*diag(init_cov)."""
        super()._shift()
        if self.batch_size > 1:
            _lib.check(_lib.lib().mjb_cov_add_diag_batched(_lib.ptr(self._cov), C.c_int(self.batch_size), C.c_int(self.d_action),
                                                           C.c_double(self.beta), _lib.ptr(self._init_cov_d), _lib.stream_ptr()))
            return
        _lib.check(_lib.lib().mjb_cov_add_diag(_lib.ptr(self._cov), C.c_int(self.d_action), C.c_double(self.beta),
                                               _lib.ptr(self._init_cov_d), _lib.stream_ptr()))

    def _calc_val(self, trajectories):
        """cem.py:107-112: mean cost-to-go."""
        costs, actions = self._traj(trajectories)
        if self.batch_size > 1:
            return self._batched_update(costs, actions, apply=False).cpu().numpy()
        s = self.shard.all_gather(self._ctg0(costs).sum().reshape(1)).sum()
        return float(s.item()) / self.num_particles
