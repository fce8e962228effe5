"""GPU counterparts of the reference's ``mjmpc/utils/control_utils.py`` hot-path helpers
(``generate_noise`` :24-34, ``cost_to_go`` :37-46, ``scale_ctrl`` :3-12), same names and
argument meaning, operating on CUDA tensors through the C ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib

NOISE_STREAM_ROLLOUT = 0      # stream ids (high 32 bits of the Philox offset)
NOISE_STREAM_ACTION = 1
NOISE_STREAM_SHIFT = 2
NOISE_STREAM_BASE = 3


class NoiseSpec:
    """Parameters of a noise tensor that is generated INSIDE the rollout kernel instead of being
    materialised by :func:`generate_noise` (same counters, covariance factor and filter -> bit-identical
    samples).  A rollout backend whose ``rollout_fn`` has ``accepts_noise_spec = True`` takes one in
    place of the ``noise`` argument; ``materialize()`` produces the tensor it stands for."""

    def __init__(self, cov, filter_coeffs, shape, base_seed, step=0, stream_id=0, k_offset=0, K_global=None,
                 zero_last=False, mean=None, particles_per_cov=0):
        self.particles_per_cov = int(particles_per_cov)      # cov is (n_instances, d, d): particles per instance
        self.cov, self.filter_coeffs, self.shape = cov, [float(b) for b in filter_coeffs], tuple(shape)
        self.base_seed, self.step, self.stream_id = int(base_seed), step, int(stream_id)
        self.k_offset = int(k_offset)
        self.K_global = int(K_global if K_global is not None else k_offset + shape[0])
        self.zero_last, self.mean = bool(zero_last), mean

    def materialize(self, out=None):
        return generate_noise(self.cov, self.filter_coeffs, self.shape, self.base_seed, step=self.step,
                              stream_id=self.stream_id, k_offset=self.k_offset, K_global=self.K_global,
                              zero_last_mean=self.mean if self.zero_last else None, out=out,
                              device=self.cov.device, particles_per_cov=self.particles_per_cov)

    def fill(self, a):
        """Write the fused-noise fields of a ``RolloutArgs``."""
        a.noise_cov = self.cov.data_ptr()
        a.noise_seed = self.base_seed & 0xFFFFFFFFFFFFFFFF
        if isinstance(self.step, torch.Tensor):
            a.noise_offset = (self.stream_id & 0xFFFFFFFF) << 32
            a.noise_step_ptr = self.step.data_ptr()
        else:
            a.noise_offset = ((self.stream_id & 0xFFFFFFFF) << 32) | (int(self.step) & 0xFFFFFFFF)
        a.noise_beta0, a.noise_beta1, a.noise_beta2 = self.filter_coeffs
        a.noise_k_offset, a.noise_K_global = self.k_offset, self.K_global
        a.noise_zero_last = int(self.zero_last)


def _dev(x, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.float64)
    return torch.as_tensor(np.asarray(x, np.float64), device=device)


def particle_minor(K, H, d, device):
    """(K,H,d)-shaped view of a freshly allocated (H,d,K) buffer: the layout every kernel
    reads/writes coalesced across particles."""
    return torch.empty((H, d, K), dtype=torch.float64, device=device).permute(2, 0, 1)


def generate_noise(cov, filter_coeffs, shape, base_seed, *, step=0, stream_id=NOISE_STREAM_ROLLOUT,
                   k_offset=0, K_global=None, zero_last_mean=None, out=None, device="cuda", particles_per_cov=0):
    """Correlated noise samples eps (K,H,d): N(0,cov) draws filtered along the horizon by
    eps[:,i] = b0*eps[:,i] + b1*eps[:,i-1] + b2*eps[:,i-2], i >= 2 (control_utils.py:24-34).

    ``base_seed`` keys the Philox generator and ``step`` (an int, or a 1-element int64 CUDA tensor read
    by the kernel) plays the role of the reference's ``+ num_steps`` reseeding (the same (seed, step) gives the same samples, as all n_iters of
    one MPC step do in the reference).  ``k_offset``/``K_global`` place a shard inside the
    global particle range; ``zero_last_mean`` applies olgaussian_mpc.py:110-111."""
    a, keep = noise_args(cov, filter_coeffs, shape, base_seed, step=step, stream_id=stream_id, k_offset=k_offset,
                         K_global=K_global, zero_last_mean=zero_last_mean, out=out, device=device,
                         particles_per_cov=particles_per_cov)
    _lib.check(_lib.lib().mjb_generate_noise(C.byref(a), _lib.stream_ptr()))
    return keep["out"]


def noise_args(cov, filter_coeffs, shape, base_seed, *, step=0, stream_id=NOISE_STREAM_ROLLOUT, k_offset=0,
               K_global=None, zero_last_mean=None, out=None, device="cuda", particles_per_cov=0):
    """The ``mjb_noise_args`` block :func:`generate_noise` launches with, and the tensors it points into
    (``keep``: hold on to it for as long as the block is used)."""
    K, H = int(shape[0]), int(shape[1])
    cov = _dev(cov, device).contiguous()
    d = cov.shape[-1]
    if out is None:
        out = particle_minor(K, H, d, cov.device)
    a = _lib.NoiseArgs()
    a.K, a.H, a.d = K, H, d
    if cov.dim() == 3:                      # (n_instances, d, d): batched controller instances with their own covariance
        a.cov_stride, a.particles_per_cov = d * d, int(particles_per_cov)
    a.k_offset = int(k_offset)
    a.K_global = int(K_global if K_global is not None else k_offset + K)
    a.seed = int(base_seed) & 0xFFFFFFFFFFFFFFFF
    if isinstance(step, torch.Tensor):      # device-resident step counter (CUDA-graph replay)
        a.offset = (int(stream_id) & 0xFFFFFFFF) << 32
        a.step_ptr = step.data_ptr()
    else:
        a.offset = noise_offset(stream_id, step)
    a.cov = cov.data_ptr()
    a.beta0, a.beta1, a.beta2 = [float(b) for b in filter_coeffs]
    keep = dict(cov=cov, out=out, step=step)
    if zero_last_mean is not None:
        zm = _dev(zero_last_mean, cov.device).contiguous()
        a.zero_last, a.neg_mean = 1, zm.data_ptr()
        keep["neg_mean"] = zm
    a.out = out.data_ptr()
    a.out_sk, a.out_st, a.out_sj = out.stride()
    return a, keep


def noise_offset(stream_id, step):
    """Philox offset word: stream id in the high 32 bits, the num_steps-like counter in the low 32."""
    return ((int(stream_id) & 0xFFFFFFFF) << 32) | (int(step) & 0xFFFFFFFF)


def cost_to_go(cost_seq, gamma_seq, out=None):
    """Discounted cost-to-go (control_utils.py:37-46), bit-identical to the reference's numpy
    evaluation order.  cost_seq (K,H) CUDA tensor (any strides), gamma_seq (1,H) or (H,) host array."""
    g = np.ascontiguousarray(np.asarray(gamma_seq, np.float64).reshape(-1))
    K, H = cost_seq.shape
    if g.shape[0] != H:
        raise ValueError("gamma_seq has %d entries for horizon %d" % (g.shape[0], H))
    if out is None:
        out = torch.empty((H, K), dtype=torch.float64, device=cost_seq.device).t()
    _lib.check(_lib.lib().mjb_cost_to_go(
        _lib.ptr(cost_seq), _lib.c_ll(cost_seq.stride(0)), _lib.c_ll(cost_seq.stride(1)),
        g.ctypes.data_as(_lib.c_double_p), C.c_int(K), C.c_int(H),
        _lib.ptr(out), _lib.c_ll(out.stride(0)), _lib.c_ll(out.stride(1)), _lib.stream_ptr()))
    return out


def scale_ctrl(ctrl, action_low_limit, action_up_limit, squash_fn='clip'):
    """control_utils.py:3-12 (no live caller on the hot path; kept for API parity)."""
    if len(ctrl.shape) == 1:
        ctrl = ctrl[None, :, None]
    act_half_range = (action_up_limit - action_low_limit) / 2.0
    act_mid_range = (action_up_limit + action_low_limit) / 2.0
    if squash_fn == 'clip':
        ctrl = ctrl.clip(-1.0, 1.0)
    elif squash_fn == 'tanh':
        ctrl = ctrl.tanh() if isinstance(ctrl, torch.Tensor) else np.tanh(ctrl)
    return act_mid_range[None, :] + ctrl * act_half_range[None, :]
