"""MPC using naive random shooting -- the reference's ``RandomShooting``
(``mjmpc/control/random_shooting.py:9-69``), GPU argmin + blend.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib
from .olgaussian_mpc import OLGaussianMPC


class RandomShooting(OLGaussianMPC):
    def __init__(self, d_state, d_obs, d_action, horizon, init_cov, base_action, num_particles, step_size,
                 gamma, n_iters, action_lows, action_highs, set_sim_state_fn=None, rollout_fn=None,
                 sample_mode='mean', filter_coeffs=[1.0, 0.0, 0.0], batch_size=1, seed=0, device=None,
                 shard=None):
        super(RandomShooting, self).__init__(d_state, d_obs, d_action, action_lows, action_highs, horizon, init_cov,
                                             np.zeros(shape=(horizon, d_action)), base_action, num_particles,
                                             gamma, n_iters, step_size, filter_coeffs, set_sim_state_fn, rollout_fn,
                                             'diagonal', sample_mode, batch_size, seed, False, device, shard)
        self.best_id = None        # device int64 (1,), global particle index

    def _update_distribution(self, trajectories):
        """random_shooting.py:52-62: mean <- (1-step)*mean + step*actions[argmin cost-to-go]."""
        L = _lib.lib()
        costs, actions = self._traj(trajectories)
        if self.batch_size > 1:
            self._batched_update(costs, actions, apply=True)
            return
        k0, kl = self.shard.local_range(self.num_particles)
        H, d = self.horizon, self.d_action
        ctg0 = self._ctg0(costs).contiguous()
        idx = self._buf("rs_idx", (1,), torch.int64)
        val = self._buf("rs_val", (1,))
        _lib.check(L.mjb_argmin(_lib.ptr(ctg0), _lib.c_ll(kl), _lib.ptr(idx), _lib.ptr(val), _lib.stream_ptr()))
        if self.shard.world_size == 1:
            self.best_id = idx
            _lib.check(L.mjb_blend_best(_lib.ptr(actions), *[_lib.c_ll(s) for s in actions.stride()], _lib.ptr(idx),
                                        _lib.c_ll(0), C.c_int(kl), C.c_int(H), C.c_int(d),
                                        C.c_double(self.step_size), _lib.ptr(self._mean), _lib.stream_ptr()))
            return
        # shards: every rank offers its best row; the winner (lowest value, then lowest global index) is
        # picked identically everywhere by a second argmin over the N candidates (rank order = index order)
        row = self._buf("rs_row", (H, d))
        _lib.check(L.mjb_gather_particles(_lib.ptr(actions), *[_lib.c_ll(s) for s in actions.stride()], _lib.ptr(idx),
                                          C.c_int(1), C.c_int(H), C.c_int(d), _lib.ptr(row), _lib.c_ll(H * d),
                                          _lib.c_ll(d), _lib.c_ll(1), _lib.stream_ptr()))
        vals = self.shard.all_gather(val).reshape(-1).contiguous()
        rows = self.shard.all_gather(row).contiguous()
        ids = self.shard.all_gather(idx + k0).reshape(-1)
        win = self._buf("rs_win", (1,), torch.int64)
        _lib.check(L.mjb_argmin(_lib.ptr(vals), _lib.c_ll(self.shard.world_size), _lib.ptr(win), None,
                                _lib.stream_ptr()))
        self.best_id = ids[win]
        _lib.check(L.mjb_blend_best(_lib.ptr(rows), _lib.c_ll(H * d), _lib.c_ll(d), _lib.c_ll(1), _lib.ptr(win),
                                    _lib.c_ll(0), C.c_int(self.shard.world_size), C.c_int(H), C.c_int(d),
                                    C.c_double(self.step_size), _lib.ptr(self._mean), _lib.stream_ptr()))

    def _batched_update(self, costs, actions, apply):
        """batch_size independent instances: one thread block per instance (argmin + blend), no reduction across
        instances; best_id holds one index per instance (within the instance's particle block)."""
        a = _lib.InstancesArgs()
        a.n_ctrl, a.K, a.H, a.d = self.batch_size, self.num_particles, self.horizon, self.d_action
        a.mode, a.apply = _lib.INST_RS, int(apply)
        a.costs = costs.data_ptr(); a.costs_sk, a.costs_st = costs.stride()
        a.actions = actions.data_ptr(); a.act_sk, a.act_st, a.act_sj = actions.stride()
        a.mean = self._mean.data_ptr()
        g = np.ascontiguousarray(self.gamma_seq.reshape(-1))
        a.gamma_seq = g.ctypes.data
        a.step_size = float(self.step_size)
        ids = self._buf("rs_ids_b", (self.batch_size,), torch.int64)
        value = self._buf("batched_value", (self.batch_size,))
        a.ids, a.value = ids.data_ptr(), value.data_ptr()
        _lib.check(_lib.lib().mjb_instances_update_batched(C.byref(a), _lib.stream_ptr()))
        self.best_id = ids
        return value

    def _calc_val(self, trajectories):
        """random_shooting.py:65-69."""
        costs, actions = self._traj(trajectories)
        if self.batch_size > 1:
            return self._batched_update(costs, actions, apply=False).cpu().numpy()
        s = self.shard.all_gather(self._ctg0(costs).sum().reshape(1)).sum()
        return float(s.item()) / self.num_particles
