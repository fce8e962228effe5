"""Particle Filter MPC -- the reference's ``PFMPC``
(``mjmpc/control/particle_filter_controller.py:15-177``) with the particle set resident in HBM.
"""
from __future__ import annotations

import copy
import ctypes as C
import random

import numpy as np
import torch

from .. import _lib
from ..utils import control_utils
from ..utils.control_utils import generate_noise
from .controller import Controller


class PFMPC(Controller):
    def __init__(self, d_state, d_obs, d_action, horizon, cov_shift, cov_resample, base_action, lam,
                 num_particles, gamma, n_iters, action_lows, action_highs, set_sim_state_fn=None,
                 rollout_fn=None, sample_mode="mean", batch_size=1, filter_coeffs=[1., 0., 0.], seed=0,
                 device=None, shard=None):
        super(PFMPC, self).__init__(d_state, d_obs, d_action, action_lows, action_highs, horizon, gamma, n_iters,
                                    set_sim_state_fn, rollout_fn, sample_mode, batch_size, seed, device, shard)
        self.lam = lam
        self.num_particles = int(num_particles)
        self.cov_shift = np.diag(np.array([cov_shift] * self.d_action))
        self.cov_resample = np.diag(np.array([cov_resample] * self.d_action))
        self.base_action = base_action
        self.filter_coeffs = filter_coeffs
        random.seed(self.seed_val)
        self._buffers = {}
        self._cov_shift_d = self._to_device(self.cov_shift)
        self._cov_resample_d = self._to_device(self.cov_resample)
        self.resample_ids = None
        # batch_size > 1: that many INDEPENDENT particle filters advance in lock step (sweeps, dynamics-
        # randomisation batches): instance b owns particles [b*K, (b+1)*K), mean row b and state row b; one
        # thread block per instance does the update (mjb_pf_update_batched), nothing crosses instances
        self.batch_size = int(batch_size)
        if self.batch_size > 1 and self.shard.world_size != 1:
            raise ValueError("independent instances are partitioned by the caller, not sharded (no collective)")
        self._particle_id_offset = 0       # first global Philox particle index of this controller's block
        self._init_particles()

    def _init_particles(self):
        k0, kl = self.shard.local_range(self.num_particles)
        B = self.batch_size
        shape = (B, self.horizon, self.d_action) if B > 1 else (self.horizon, self.d_action)
        self._mean = torch.zeros(shape, dtype=torch.float64, device=self.device)
        # particle_filter_controller.py:69-71 (base_seed = seed_val)
        self._samples = generate_noise(self._cov_resample_d, self.filter_coeffs, shape=(kl * B, self.horizon),
                                       base_seed=self.seed_val, step=0, k_offset=k0 + self._particle_id_offset,
                                       K_global=self._particle_id_offset + self.num_particles * B, device=self.device)

    def set_instance_offset(self, first_instance: int):
        """Batched independent filters partitioned over several GPUs (no collective): this controller holds
        instances [first_instance, first_instance + batch_size) of the sweep.  Moves the Philox particle ids and
        re-draws the particle sets, so every instance starts from (and keeps drawing) what it would in the
        single-GPU sweep."""
        self._particle_id_offset = int(first_instance) * self.num_particles
        self.reset()

    def _graphable(self):
        return False          # the resampler's r comes from Python's random module every step

    def _buf(self, name, shape, dtype=torch.float64, zero=False):
        key = (name, tuple(shape), dtype)
        b = self._buffers.get(key)
        if b is None:
            b = (torch.zeros if zero else torch.empty)(tuple(shape), dtype=dtype, device=self.device)
            self._buffers[key] = b
        return b

    @property
    def local_particles(self):
        return self.shard.local_range(self.num_particles)[1] * self.batch_size

    def _first_action(self):
        return self._mean[:, 0].clone() if self.batch_size > 1 else self._mean[0].clone()

    @property
    def mean_action(self):
        return self._mean.cpu().numpy()

    @mean_action.setter
    def mean_action(self, value):
        shape = (self.batch_size, self.horizon, self.d_action) if self.batch_size > 1 else (self.horizon, self.d_action)
        self._mean = self._to_device(value).reshape(shape).contiguous()

    @property
    def action_samples(self):
        """(K_local, H, d) numpy copy of this shard's particle set."""
        return np.ascontiguousarray(self._samples.cpu().numpy())

    @action_samples.setter
    def action_samples(self, value):
        v = self._to_device(value)
        self._samples = v.permute(1, 2, 0).contiguous().permute(2, 0, 1)

    # ---- particle_filter_controller.py:74-90 -----------------------------------------------------------
    def generate_rollouts(self, state):
        if state is not None:             # None: the backend already holds the (device-resident) state
            if self.batch_size > 1 and isinstance(state, dict):
                state = [state] * self.batch_size          # one state for every instance
            self._set_sim_state_fn(copy.deepcopy(state))
        kl, H, d = self.local_particles, self.horizon, self.d_action
        delta = self._buf("delta", (H, d, kl)).permute(2, 0, 1)
        s = self._samples
        _lib.check(_lib.lib().mjb_particle_sub_mean_batched(
            _lib.ptr(s), *[_lib.c_ll(x) for x in s.stride()], _lib.ptr(self._mean), C.c_int(self.batch_size),
            C.c_int(kl // self.batch_size), C.c_int(H), C.c_int(d), _lib.ptr(delta),
            *[_lib.c_ll(x) for x in delta.stride()], _lib.stream_ptr()))
        return self._rollout_fn(kl, H, self._mean, delta, mode="open_loop")

    def _weights(self, costs):
        """softmax(-ctg0/lam) over ALL particles (particle_filter_controller.py:104-113); returns this
        shard's normalised weights."""
        L = _lib.lib()
        kl, H, d = self.local_particles, self.horizon, self.d_action
        a = _lib.SoftmaxArgs()
        a.K, a.H, a.d = kl, H, d
        a.costs = costs.data_ptr(); a.costs_sk, a.costs_st = costs.stride()
        s = self._samples
        a.actions = s.data_ptr(); a.act_sk, a.act_st, a.act_sj = s.stride()
        a.mean = self._mean.data_ptr()
        g = np.ascontiguousarray(self.gamma_seq.reshape(-1))
        a.gamma_seq = g.ctypes.data
        a.lam = float(self.lam)
        total = self._buf("total", (1, kl))
        scratch = self._buf("sm_scratch", (int(L.mjb_softmax_scratch_doubles(kl, H, d, 0)),), zero=True)
        P = L.mjb_softmax_partial_doubles(H, d, 0, 0)
        partials = self._buf("sm_partials", (P,))
        a.total, a.scratch, a.partials = total.data_ptr(), scratch.data_ptr(), partials.data_ptr()
        _lib.check(L.mjb_softmax_partials(C.byref(a), _lib.stream_ptr()))
        allp = self.shard.all_gather(partials)
        stats = self._buf("sm_stats", (4,))
        c = _lib.CombineArgs()
        c.H, c.d, c.n_shards, c.K_global = H, d, self.shard.world_size, self.num_particles
        c.partials, c.lam, c.step_size = allp.data_ptr(), float(self.lam), 1.0
        c.stats = stats.data_ptr()
        _lib.check(L.mjb_softmax_combine(C.byref(c), _lib.stream_ptr()))
        w = self._buf("weights", (kl,))
        _lib.check(L.mjb_softmax_weights(_lib.ptr(total), C.c_int(kl), _lib.ptr(stats), C.c_int(0),
                                         C.c_double(self.lam), _lib.ptr(w), _lib.stream_ptr()))
        return w

    # ---- particle_filter_controller.py:92-102 ----------------------------------------------------------
    def _update_distribution(self, trajectories):
        costs = self._to_device(trajectories["costs"])
        if tuple(costs.shape) != (self.local_particles, self.horizon):
            raise ValueError("rollout_fn returned costs %s for K=%d H=%d" % (tuple(costs.shape), self.local_particles,
                                                                             self.horizon))
        random.seed(self.seed_val + self.num_steps)
        np.random.seed((self.seed_val + self.num_steps) % (2 ** 32))
        if self.batch_size > 1:
            self._batched_update(costs)
            return
        w = self._weights(costs)
        self._samples = self._resampling(self._samples, w, low_variance=True)
        self._update_mean()

    def _batched_update(self, costs):
        """batch_size independent filters: weights, resampling and mean of every instance in one launch.
        Every instance draws r = random.uniform(0, 1/M) after the same random.seed(seed_val + num_steps)
        (particle_filter_controller.py:99,163), i.e. the same r."""
        B, K, H, d = self.batch_size, self.num_particles, self.horizon, self.d_action
        r = random.uniform(0.0, 1.0 / K * 1.0)
        a = _lib.PfBatchedArgs()
        a.n_ctrl, a.K, a.H, a.d = B, K, H, d
        a.costs = costs.data_ptr(); a.costs_sk, a.costs_st = costs.stride()
        s = self._samples
        a.samples = s.data_ptr(); a.s_sk, a.s_st, a.s_sj = s.stride()
        g = np.ascontiguousarray(self.gamma_seq.reshape(-1))
        a.gamma_seq = g.ctypes.data
        a.lam = float(self.lam)
        rd = self._buf("pf_r", (B,))
        rd.fill_(r)
        w = self._buf("weights", (B * K,))
        idx = self._buf("resample_idx", (B * K,), torch.int64)
        out = torch.empty((H, d, B * K), dtype=torch.float64, device=self.device).permute(2, 0, 1)
        a.r, a.weights, a.idx = rd.data_ptr(), w.data_ptr(), idx.data_ptr()
        a.out = out.data_ptr(); a.o_sk, a.o_st, a.o_sj = out.stride()
        a.mean = self._mean.data_ptr()
        _lib.check(_lib.lib().mjb_pf_update_batched(C.byref(a), _lib.stream_ptr()))
        self._samples = out
        self.resample_ids = idx.reshape(B, K)
        self._last_weights = w.reshape(B, K)
        self._last_r = r

    def _update_mean(self):
        """mean_action = np.mean(action_samples, axis=0) over all shards."""
        L = _lib.lib()
        kl, H, d = self.local_particles, self.horizon, self.d_action
        s = self._samples
        scratch = self._buf("mean_scratch", (int(L.mjb_elite_scratch_doubles(kl, H, d)),))
        local = self._buf("mean_local", (H, d))
        _lib.check(L.mjb_particle_mean(_lib.ptr(s), *[_lib.c_ll(x) for x in s.stride()], C.c_int(kl), C.c_int(H),
                                       C.c_int(d), _lib.ptr(scratch), _lib.ptr(local), _lib.stream_ptr()))
        if self.shard.world_size == 1:
            self._mean = local.clone()
        else:
            self._mean = self.shard.all_gather(local).mean(dim=0)

    def sample_actions(self):
        return self.action_samples

    # ---- particle_filter_controller.py:118-125 ---------------------------------------------------------
    def _get_next_action(self, state, mode='mean'):
        return self._first_action().cpu().numpy().copy()

    # ---- particle_filter_controller.py:127-150 ---------------------------------------------------------
    def _shift(self):
        if self.base_action not in _lib.BASE_ACTIONS:
            raise NotImplementedError("invalid option for base action during shift")
        k0, kl = self.shard.local_range(self.num_particles)
        kl *= self.batch_size
        H, d = self.horizon, self.d_action
        delta = generate_noise(self._cov_shift_d, self.filter_coeffs, shape=(kl, H), base_seed=self.seed_val,
                               step=self.num_steps, stream_id=control_utils.NOISE_STREAM_SHIFT,
                               k_offset=k0 + self._particle_id_offset,
                               K_global=self._particle_id_offset + self.num_particles * self.batch_size,
                               out=self._buf("delta", (H, d, kl)).permute(2, 0, 1), device=self.device)
        rnd = None
        if self.base_action == 'random':
            if self.batch_size > 1:
                raise NotImplementedError("base_action 'random' is not available for batched instances")
            # np.random.normal(0, self.cov_resample, d) broadcasts one (d,) row (the matrix diagonal as std)
            rnd = generate_noise(self._cov_resample_d ** 2, [1.0, 0.0, 0.0], shape=(1, 1), base_seed=self.seed_val,
                                 step=self.num_steps, stream_id=control_utils.NOISE_STREAM_BASE,
                                 device=self.device).reshape(-1).contiguous()
        s = self._samples
        _lib.check(_lib.lib().mjb_pf_shift(
            _lib.ptr(s), *[_lib.c_ll(x) for x in s.stride()], _lib.ptr(delta), *[_lib.c_ll(x) for x in delta.stride()],
            C.c_int(kl), C.c_int(H), C.c_int(d), C.c_int(_lib.BASE_ACTIONS[self.base_action]), _lib.ptr(rnd),
            _lib.stream_ptr()))

    # ---- particle_filter_controller.py:152-157 ---------------------------------------------------------
    def reset(self):
        self.num_steps = 0
        self._buffers = {}
        self._init_particles()

    # ---- particle_filter_controller.py:159-174 ---------------------------------------------------------
    def _resampling(self, act_seq, weights, low_variance=True):
        """Systematic resampling.  r = random.uniform(0, 1/M) comes from Python's generator exactly as
        in the reference; the cumulative sum runs in the reference's sequential order on the GPU so the
        indices are bit-identical for bit-identical weights."""
        if not low_variance:
            raise NotImplementedError("only low-variance (systematic) resampling runs on the GPU")
        L = _lib.lib()
        M = self.num_particles
        k0, kl = self.shard.local_range(M)
        H, d = self.horizon, self.d_action
        r = random.uniform(0.0, 1.0 / M * 1.0)
        allw = self.shard.all_gather(weights.contiguous()).reshape(-1)
        cs = self._buf("cumsum", (M + 2,))
        idx = self._buf("resample_idx", (M,), torch.int64)
        _lib.check(L.mjb_resample_indices(_lib.ptr(allw), _lib.c_ll(M), C.c_double(r), _lib.ptr(cs), _lib.ptr(idx),
                                          _lib.stream_ptr()))
        self.resample_ids = idx
        if self.shard.world_size == 1:
            src = act_seq
        else:
            # every rank needs rows that may live anywhere: all-gather the particle set (K*H*d doubles over NVLink)
            src = self.shard.all_gather(act_seq.permute(1, 2, 0).contiguous())      # (N, H, d, kl)
            src = src.permute(1, 2, 0, 3).reshape(H, d, M).permute(2, 0, 1)
        out = torch.empty((H, d, kl), dtype=torch.float64, device=self.device).permute(2, 0, 1)
        _lib.check(L.mjb_gather_particles(_lib.ptr(src), *[_lib.c_ll(x) for x in src.stride()],
                                          _lib.ptr(idx[k0:k0 + kl]), C.c_int(kl), C.c_int(H), C.c_int(d),
                                          _lib.ptr(out), *[_lib.c_ll(x) for x in out.stride()], _lib.stream_ptr()))
        return out

    def _calc_val(self, trajectories):
        raise NotImplementedError("_calc val not implemented yet")
