"""MPC with open-loop Gaussian policies -- the reference's ``OLGaussianMPC``
(``mjmpc/control/olgaussian_mpc.py:10-139``) with its numeric bodies on the GPU: ``sample_noise``
(:88-93) is the Philox kernel, ``_shift`` (:116-129) a device kernel, and the softmax /
elite reductions used by the subclasses go through the C ABI.
"""
from __future__ import annotations

import copy
import ctypes as C
import os

import numpy as np
import torch

from .. import _lib
from ..utils import control_utils
from ..utils.control_utils import generate_noise
from .controller import Controller


class OLGaussianMPC(Controller):
    def __init__(self, d_state, d_obs, d_action, action_lows, action_highs, horizon, init_cov, init_mean,
                 base_action, num_particles, gamma, n_iters, step_size, filter_coeffs, set_sim_state_fn=None,
                 rollout_fn=None, cov_type='diagonal', sample_mode='mean', batch_size=1, seed=0,
                 use_zero_control_seq=False, device=None, shard=None):
        super(OLGaussianMPC, self).__init__(d_state, d_obs, d_action, action_lows, action_highs, horizon,
                                            gamma, n_iters, set_sim_state_fn, rollout_fn, sample_mode,
                                            batch_size, seed, device, shard)
        # reference: np.array([init_cov] * d_action) (olgaussian_mpc.py:58); a length-d sequence (as some
        # shipped configs give, reacher_7dof-v0.yml:34) is taken as the per-dimension variances.
        ic = np.asarray(init_cov, dtype=np.float64).reshape(-1)
        if ic.size == 1:
            ic = np.array([float(ic[0])] * self.d_action)
        elif ic.size != self.d_action:
            raise ValueError("init_cov must be a scalar or have d_action entries")
        self.init_cov = ic
        self.init_mean = np.array(init_mean, dtype=np.float64).copy()
        self.base_action = base_action
        self.num_particles = int(num_particles)
        if cov_type == 'diag':          # README spelling (README.md:103); the code tests 'diagonal'
            cov_type = 'diagonal'
        self.cov_type = cov_type
        self.step_size = step_size
        self.filter_coeffs = filter_coeffs
        self.use_zero_control_seq = use_zero_control_seq
        # batch_size > 1: that many INDEPENDENT controller instances (sweeps, dynamics-randomisation batches)
        # advance in lock step -- instance b owns particles [b*K, (b+1)*K), mean row b, state row b and (with
        # an n_workers = batch_size backend) its own model; there is no reduction across instances
        self.batch_size = int(batch_size)
        if self.batch_size > 1:
            if self.shard.world_size != 1:
                raise ValueError("independent instances are partitioned by the caller, not sharded (no collective)")
            if use_zero_control_seq:
                raise NotImplementedError("use_zero_control_seq is not available for batched instances")
        self._particle_id_offset = 0       # first global Philox particle index of this controller's block
        # fuse_noise = True draws the noise inside the rollout kernel when the backend can
        # (rollout_fn.accepts_noise_spec): same samples as sample_noise(), and the (K,H,d) noise tensor never
        # exists in HBM (-117 MB at K=65536).  Off by default: measured on B200 the extra Philox / Box-Muller
        # instructions lengthen the latency-bound rollout by more (+0.09 ms) than the separate, fully
        # occupied noise kernel costs (0.065 ms).
        self.fuse_noise = False
        # overlap_noise = True (native step only): the NEXT step's noise -- a function of (seed, step) alone when
        # the covariance is fixed and there is no zero control sequence -- is drawn on a side stream while this
        # step rolls out, into the second of two noise tensors.  Same samples, same results; measured on B200 at
        # K = 65536: 0.744 -> 0.712 ms per step (the FP64-bound rollout leaves the integer / SFU pipes to the noise
        # kernel).  Under CUDA-graph replay two graphs alternate (read tensor A / write B, read B / write A).
        self.overlap_noise = True
        self._mean = self._mean_from(self.init_mean)
        self._init_cov_d = self._to_device(self.init_cov)
        self._cov = self._cov_from_init()
        self._buffers = {}
        self._noise_step = None       # device step counter while a CUDA graph is active

    # ---- numpy views of the distribution (public attributes in the reference) ----------------
    @property
    def mean_action(self):
        return self._mean.cpu().numpy()

    @mean_action.setter
    def mean_action(self, value):
        self.disable_cuda_graph()
        self._mean = self._mean_from(value)

    def _mean_from(self, value):
        m = self._to_device(value)
        if self.batch_size > 1:
            if m.numel() == self.horizon * self.d_action:
                m = m.reshape(1, self.horizon, self.d_action).expand(self.batch_size, -1, -1)
            return m.reshape(self.batch_size, self.horizon, self.d_action).contiguous().clone()
        return m.reshape(self.horizon, self.d_action).contiguous().clone()

    def _first_action(self):
        """mean_action[0] (per instance when batched), device tensor."""
        return self._mean[:, 0].clone() if self.batch_size > 1 else self._mean[0].clone()

    @property
    def cov_action(self):
        return self._cov.cpu().numpy()

    @cov_action.setter
    def cov_action(self, value):
        self.disable_cuda_graph()
        v = self._to_device(value)
        shape = (self.batch_size, self.d_action, self.d_action) if self._cov.dim() == 3 else (self.d_action, self.d_action)
        self._cov = v.expand(shape).contiguous().clone() if v.dim() < len(shape) else v.reshape(shape).contiguous()

    # ---- sharding ----------------------------------------------------------------------------------
    def set_instance_offset(self, first_instance: int):
        """Batched independent instances partitioned over several GPUs (no collective): this controller holds
        instances [first_instance, first_instance + batch_size) of the sweep.  Only the Philox particle ids move,
        so every instance draws the noise it would draw in the single-GPU sweep."""
        self._particle_id_offset = int(first_instance) * self.num_particles
        self.__dict__.pop("_fused_blocks", None)

    @property
    def local_particles(self):
        return self.shard.local_range(self.num_particles)[1] * self.batch_size

    # batched instances of a controller that ADAPTS its covariance (CEM) keep one covariance per instance
    _per_instance_cov = False

    def _cov_from_init(self):
        c = self._to_device(np.diag(self.init_cov)).contiguous()
        if self._per_instance_cov and self.batch_size > 1:
            if self.num_particles % 32 != 0:
                raise ValueError("batched instances with their own covariance need num_particles in multiples of 32")
            c = c.reshape(1, self.d_action, self.d_action).repeat(self.batch_size, 1, 1).contiguous()
        return c

    def _buf(self, name, shape, dtype=torch.float64, zero=False):
        key = (name, tuple(shape), dtype)
        b = self._buffers.get(key)
        if b is None:
            b = (torch.zeros if zero else torch.empty)(tuple(shape), dtype=dtype, device=self.device)
            self._buffers[key] = b
        return b

    # ---- olgaussian_mpc.py:69-78 ---------------------------------------------------------------------
    def _get_next_action(self, state, mode='mean'):
        if mode == 'mean':
            next_action = self._first_action().cpu().numpy()
        elif mode == 'sample':
            B = self.batch_size
            delta = generate_noise(self._cov, self.filter_coeffs, shape=(B, 1), base_seed=self.seed_val,
                                   step=123 * self.num_steps, stream_id=control_utils.NOISE_STREAM_ACTION,
                                   device=self.device)
            shape = (B, self.d_action) if B > 1 else (self.d_action,)
            next_action = self._first_action().cpu().numpy() + delta.reshape(shape).cpu().numpy().copy()
        else:
            raise ValueError('Unidentified sampling mode in get_next_action')
        return next_action

    # ---- olgaussian_mpc.py:88-93 ---------------------------------------------------------------------
    def noise_spec(self):
        """What sample_noise() would draw, as parameters (for backends that generate noise in-kernel)."""
        k0, kl = self.shard.local_range(self.num_particles)
        kl *= self.batch_size
        k0 += self._particle_id_offset
        return control_utils.NoiseSpec(
            self._cov, self.filter_coeffs, (kl, self.horizon), self.seed_val,
            step=self._noise_step if self._noise_step is not None else self.num_steps,
            k_offset=k0, K_global=self._particle_id_offset + self.num_particles * self.batch_size,
            zero_last=self.use_zero_control_seq, mean=self._mean,
            particles_per_cov=self.num_particles if self._cov.dim() == 3 else 0)

    def sample_noise(self):
        spec = self.noise_spec()
        out = self._buf("noise", (self.horizon, self.d_action, spec.shape[0])).permute(2, 0, 1)
        return spec.materialize(out=out)

    # ---- olgaussian_mpc.py:95-114 --------------------------------------------------------------------
    def generate_rollouts(self, state):
        if state is not None:             # None: the backend already holds the (device-resident) state
            if self.batch_size > 1 and isinstance(state, dict):
                state = [state] * self.batch_size          # one state for every instance
            self._set_sim_state_fn(copy.deepcopy(state))
        if self.fuse_noise and getattr(self._rollout_fn, "accepts_noise_spec", False):
            delta = self.noise_spec()     # drawn inside the rollout kernel
        else:
            delta = self.sample_noise()   # use_zero_control_seq is applied inside the noise kernel
        trajectories = self._rollout_fn(self.local_particles, self.horizon, self._mean, delta, mode="open_loop")
        return trajectories

    # ---- olgaussian_mpc.py:116-129 -------------------------------------------------------------------
    def _shift(self):
        if self.base_action not in _lib.BASE_ACTIONS:
            raise NotImplementedError("invalid option for base action during shift")
        rnd = None
        B = self.batch_size
        if self.base_action == 'random':
            # np.random.normal(0, self.init_cov, d): the reference passes the variances as std devs
            rnd = generate_noise(torch.diag(self._init_cov_d ** 2), [1.0, 0.0, 0.0], shape=(B, 1),
                                 base_seed=self.seed_val, step=self.num_steps,
                                 stream_id=control_utils.NOISE_STREAM_BASE, device=self.device).reshape(B, -1).contiguous()
        if B > 1:
            _lib.check(_lib.lib().mjb_shift_mean_batched(_lib.ptr(self._mean), C.c_int(B), C.c_int(self.horizon),
                                                         C.c_int(self.d_action), C.c_int(_lib.BASE_ACTIONS[self.base_action]),
                                                         _lib.ptr(rnd), _lib.stream_ptr()))
            return
        _lib.check(_lib.lib().mjb_shift_mean(_lib.ptr(self._mean), C.c_int(self.horizon), C.c_int(self.d_action),
                                             C.c_int(_lib.BASE_ACTIONS[self.base_action]), _lib.ptr(rnd),
                                             _lib.stream_ptr()))

    # ---- olgaussian_mpc.py:131-135 -------------------------------------------------------------------
    def reset(self):
        self.disable_cuda_graph()
        self.num_steps = 0
        self._mean = self._mean_from(np.zeros((self.horizon, self.d_action)))
        self._cov = self._cov_from_init()
        self.gamma_seq = np.cumprod([1.0] + [self.gamma] * (self.horizon - 1)).reshape(1, self.horizon)
        self._buffers = {}
        self.__dict__.pop("_fused_blocks", None)

    def _calc_val(self, trajectories):
        raise NotImplementedError("_calc_val not implemented")

    # ---- shared GPU reductions -----------------------------------------------------------------------
    def _traj(self, trajectories):
        costs = self._to_device(trajectories["costs"])
        actions = self._to_device(trajectories["actions"])
        kl = self.local_particles
        if tuple(costs.shape) != (kl, self.horizon) or tuple(actions.shape) != (kl, self.horizon, self.d_action):
            raise ValueError("rollout_fn returned costs %s / actions %s for K=%d H=%d d=%d" % (
                tuple(costs.shape), tuple(actions.shape), kl, self.horizon, self.d_action))
        return costs, actions

    def _softmax_update(self, costs, actions, lam, control_cost=False, time_based=False, cov_mode=_lib.COV_NONE,
                        apply=True, td=None):
        """Phase 1 (local partials) -> all-gather over shards -> phase 2 (combine, smooth).
        Returns the (2+2T,) stats tensor: value, global min, normalisers, minima.
        td = (td_lam, gamma, weight_seq, qvals or None): weights from MPPIQ's TD(lambda) returns instead of
        the discounted cost-to-go."""
        # `keep`: host arrays the phase-1 block points into (gamma_seq, TD weights) -- alive until the calls return
        a, c, P, stats, keep = self._softmax_blocks(costs, actions, lam, control_cost, time_based, cov_mode, apply, td)
        L = _lib.lib()
        _lib.check(L.mjb_softmax_partials(C.byref(a), _lib.stream_ptr()))
        partials = self._buf("sm_partials", (P,))
        px = self._peer_exchange(P)
        if px is not None:
            # fused: P2P stores into every peer's buffer + flag wait + rank-ordered combine in ONE kernel
            _lib.check(L.mjb_softmax_exchange_combine(C.byref(c), _lib.ptr(partials), C.c_void_p(px.peer_ptrs_dev),
                                                      C.c_int(self.shard.rank), C.c_ulonglong(px.next_seq()),
                                                      _lib.stream_ptr()))
        else:
            allp = self.shard.all_gather(partials)
            c.partials = allp.data_ptr()
            _lib.check(L.mjb_softmax_combine(C.byref(c), _lib.stream_ptr()))
        del keep
        return stats

    def _softmax_blocks(self, costs, actions, lam, control_cost=False, time_based=False, cov_mode=_lib.COV_NONE,
                        apply=True, td=None):
        """The mjb_softmax_args / mjb_combine_args blocks of one update over this controller's persistent
        buffers: (phase-1 block, phase-2 block, partial length P, stats tensor, host arrays the blocks point
        into).  The phase-2 block reads this rank's own partial vector until a caller points it elsewhere."""
        if self.batch_size > 1:
            raise NotImplementedError("batched instances are implemented for MPPI without time-based weights")
        L = _lib.lib()
        kl, H, d = self.local_particles, self.horizon, self.d_action
        T = H if time_based else 1
        P = L.mjb_softmax_partial_doubles(H, d, int(time_based), cov_mode)
        a = _lib.SoftmaxArgs()
        a.K, a.H, a.d = kl, H, d
        a.costs = costs.data_ptr(); a.costs_sk, a.costs_st = costs.stride()
        a.actions = actions.data_ptr(); a.act_sk, a.act_st, a.act_sj = actions.stride()
        a.mean = self._mean.data_ptr()
        a.cov = self._cov.data_ptr()
        g = np.ascontiguousarray(self.gamma_seq.reshape(-1))
        keep = [g]
        a.gamma_seq = g.ctypes.data
        a.lam = float(lam)
        a.control_cost, a.time_based, a.cov_mode = int(control_cost), int(time_based), int(cov_mode)
        if td is not None:
            td_lam, td_gamma, wseq, qvals = td
            wseq = np.ascontiguousarray(wseq, dtype=np.float64)     # host array, kept alive until the call returns
            keep.append(wseq)
            a.returns, a.td_lam, a.td_gamma = _lib.RETURNS_TD_LAMBDA, float(td_lam), float(td_gamma)
            a.td_weight_seq = wseq.ctypes.data
            if qvals is not None:
                a.qvals = qvals.data_ptr(); a.q_sk, a.q_st = qvals.stride()
        total = self._buf("total", (T, kl))
        # zero once: scratch[0] is the "blocks done" counter of mjb_softmax_update_fused (it resets itself)
        scratch = self._buf("sm_scratch", (int(L.mjb_softmax_scratch_doubles(kl, H, d, cov_mode)),), zero=True)
        partials = self._buf("sm_partials", (P,))
        a.total, a.scratch, a.partials = total.data_ptr(), scratch.data_ptr(), partials.data_ptr()
        stats = self._buf("sm_stats", (2 + 2 * T,))
        c = _lib.CombineArgs()
        c.H, c.d, c.n_shards, c.K_global = H, d, self.shard.world_size, self.num_particles
        c.lam, c.step_size = float(lam), float(self.step_size)
        c.time_based, c.cov_mode = int(time_based), int(cov_mode)
        c.partials = partials.data_ptr()
        c.mean = self._mean.data_ptr() if apply else None
        c.cov = self._cov.data_ptr() if (apply and cov_mode != _lib.COV_NONE) else None
        c.stats = stats.data_ptr()
        self._last_total = total
        return a, c, P, stats, keep

    # ---- the whole MPC step behind one native call -----------------------------------------------------
    def _softmax_spec(self):
        """dict(lam, control_cost, time_based, cov_mode, cov_shift_beta) when this controller's update is the
        plain softmax reduction (MPPI, DMD-MPC); None otherwise."""
        return None

    def _fused_step(self, state, hotstart=True):
        """One MPC step -- n_iters x (noise, rollout, softmax update), next action, shift -- through
        ``mjb_softmax_mpc_step``: the same kernels on the same buffers as the step-by-step path (bit-identical
        results), launched from one native call instead of ~10 Python-level ones.  Returns the (d,) device
        action (a persistent buffer, overwritten by the next step), or None -- with nothing done -- when this
        controller / backend / option set needs the step-by-step path.  ``state=None`` keeps the state the
        backend already holds."""
        spec = self._softmax_spec()
        backend = getattr(self._rollout_fn, "backend", None)
        if (spec is None or backend is None or not hasattr(backend, "rollout_args") or self.batch_size > 1
                or self.fuse_noise or self.base_action not in ("null", "repeat") or self.sample_mode != 'mean'
                or type(self).check_convergence is not Controller.check_convergence
                or os.environ.get("MJB_FUSED_STEP", "1") == "0"):
            return None
        L = _lib.lib()
        kl, H, d = self.local_particles, self.horizon, self.d_action
        P = L.mjb_softmax_partial_doubles(H, d, int(spec["time_based"]), spec["cov_mode"])
        px = self._peer_exchange(P)
        if self.shard.world_size > 1 and px is None:
            return None                 # NCCL all-gather between the two phases: host work inside the step
        if state is not None:
            self._set_state(state)
        overlap = bool(self.overlap_noise) and spec["cov_mode"] == _lib.COV_NONE and not self.use_zero_control_seq
        key = (self._mean.data_ptr(), self._cov.data_ptr(), backend._state.data_ptr(), backend.model.handle.value,
               kl, H, self.n_iters, tuple(sorted(spec.items())), self.step_size, self.base_action, hotstart,
               tuple(self.filter_coeffs), self.use_zero_control_seq, self.seed_val, self.gamma, overlap)
        blk = self.__dict__.get("_fused_blocks")
        if blk is None or blk["key"] != key:
            ns = self.noise_spec()
            noise = self._buf("noise", (H, d, kl)).permute(2, 0, 1)
            na, nkeep = control_utils.noise_args(ns.cov, ns.filter_coeffs, ns.shape, ns.base_seed, step=0,
                                                 stream_id=ns.stream_id, k_offset=ns.k_offset, K_global=ns.K_global,
                                                 zero_last_mean=ns.mean if ns.zero_last else None, out=noise,
                                                 device=self.device)
            costs = self._buf("fused_costs", (H, kl)).t()
            actions = self._buf("fused_actions", (H, d, kl)).permute(2, 0, 1)
            ra, rout = backend.rollout_args(kl, H, self._mean, noise, costs=costs, actions=actions)
            sa, ca, _, stats, skeep = self._softmax_blocks(costs, actions, spec["lam"], spec["control_cost"],
                                                           spec["time_based"], spec["cov_mode"])
            st = _lib.MpcStepArgs()
            st.n_iters = int(self.n_iters)
            st.noise, st.rollout, st.softmax, st.combine = C.pointer(na), C.pointer(ra), C.pointer(sa), C.pointer(ca)
            st.model = backend.model.handle
            st.rank = self.shard.rank
            st.peer_bufs_dev = px.peer_ptrs_dev if px is not None else None
            action = self._buf("fused_action", (d,))
            st.action_out = action.data_ptr()
            st.shift = int(bool(hotstart))
            st.base_action = _lib.BASE_ACTIONS[self.base_action]
            st.cov_shift_beta = float(spec["cov_shift_beta"])
            blk = dict(key=key, st=st, na=na, ra=ra, action=action, keep=(sa, ca, nkeep, rout, skeep, stats, noise))
            if overlap:
                # second noise tensor + its argument block; `ready` = (tensor index, step) of noise already drawn
                noise_b = self._buf("noise_b", (H, d, kl)).permute(2, 0, 1)
                nb, nbkeep = control_utils.noise_args(ns.cov, ns.filter_coeffs, ns.shape, ns.base_seed, step=0,
                                                      stream_id=ns.stream_id, k_offset=ns.k_offset, K_global=ns.K_global,
                                                      out=noise_b, device=self.device)
                blk.update(nargs=(na, nb), bufs=(noise, noise_b), ready=None, keep2=(nbkeep, noise_b))
            self._fused_blocks = blk
        na = blk["na"]
        st = blk["st"]
        if overlap:
            ready = blk["ready"]
            cur = ready[0] if (ready is not None and ready[1] == self.num_steps) else 0
            n_cur, n_next = blk["nargs"][cur], blk["nargs"][1 - cur]
            if self._noise_step is not None:
                # CUDA-graph capture / replay: the step comes from the device counter, the next step is counter + 1
                n_cur.offset, n_next.offset = (control_utils.noise_offset(control_utils.NOISE_STREAM_ROLLOUT, i) for i in (0, 1))
                n_cur.step_ptr = n_next.step_ptr = self._noise_step.data_ptr()
            else:
                n_cur.offset = control_utils.noise_offset(control_utils.NOISE_STREAM_ROLLOUT, self.num_steps)
                n_next.offset = control_utils.noise_offset(control_utils.NOISE_STREAM_ROLLOUT, self.num_steps + 1)
                n_cur.step_ptr = n_next.step_ptr = None
            buf = blk["bufs"][cur]
            ra = blk["ra"]
            ra.noise = buf.data_ptr()
            ra.noise_sk, ra.noise_st, ra.noise_sj = buf.stride()
            # this step's noise: already there when the previous step drew it, otherwise drawn in line
            st.noise = None if (ready is not None and ready == (cur, self.num_steps)) else C.pointer(n_cur)
            st.noise_next = C.pointer(n_next)
            blk["ready"] = (1 - cur, self.num_steps + 1)
        elif self._noise_step is not None:
            na.offset, na.step_ptr = control_utils.noise_offset(control_utils.NOISE_STREAM_ROLLOUT, 0), self._noise_step.data_ptr()
        else:
            na.offset, na.step_ptr = control_utils.noise_offset(control_utils.NOISE_STREAM_ROLLOUT, self.num_steps), None
        if px is not None:
            st.seq = px.seq + 1
            px.seq += self.n_iters
        _lib.check(L.mjb_softmax_mpc_step(C.byref(st), _lib.stream_ptr()))
        return blk["action"]

    def _draw_step_noise(self, parity, step):
        """Overlapped-noise bookkeeping of the graph replay: draw the noise of MPC step `step` into noise tensor
        `parity` now (the graphs only ever draw the NEXT step's noise)."""
        blk = self._fused_blocks
        na = blk["nargs"][parity]
        na.offset, na.step_ptr = control_utils.noise_offset(control_utils.NOISE_STREAM_ROLLOUT, step), None
        _lib.check(_lib.lib().mjb_generate_noise(C.byref(na), _lib.stream_ptr()))

    def _peer_exchange(self, P):
        """Symmetric-memory exchange buffer for partial vectors of P doubles (None: single GPU, disabled
        with MJB_P2P=0, or symmetric memory unavailable -> NCCL all-gather).  Creation is collective and
        happens at the same call on every rank."""
        if self.shard.world_size == 1 or os.environ.get("MJB_P2P", "1") == "0":
            return None
        cache = self.__dict__.setdefault("_px", {})
        if P not in cache:
            px = None
            try:
                from ..utils.shard import PeerExchange
                px = PeerExchange(self.shard, P, self.device)
            except Exception as e:            # pragma: no cover - depends on the box
                import warnings
                warnings.warn("NVLink peer exchange unavailable (%r); using NCCL all-gather" % (e,))
            # all ranks must take the same path: agree on the minimum
            ok = torch.tensor([1.0 if px is not None else 0.0], dtype=torch.float64, device=self.device)
            if self.shard.all_gather(ok).min().item() < 1.0:
                px = None
            cache[P] = px
        return cache[P]

    def _ctg0(self, costs):
        """cost_to_go(costs, gamma_seq)[:, 0] for the local particles (device, contiguous)."""
        kl = self.local_particles
        out = self._buf("ctg", (self.horizon, kl)).t()
        control_utils.cost_to_go(costs, self.gamma_seq, out=out)
        return out[:, 0]
