"""ctypes binding of ``libmjmpc_b200.so`` (the C ABI declared in ``include/mjmpc_b200.h``).

There is no CPU fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MJB_LIB_PATH") or os.path.join(_HERE, "libmjmpc_b200.so")   # override: kernel-variant experiments

MJB_OK, MJB_EINVAL, MJB_ECUDA, MJB_ENOTIMPL = 0, 1, 2, 3
MODEL_NPARAM = 167
STATE_DIM = 17
OBS_DIM = 20

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_ll = C.c_longlong


class RolloutArgs(C.Structure):
    _fields_ = [
        ("K", C.c_int), ("H", C.c_int), ("particles_per_ctrl", C.c_int), ("particles_per_model", C.c_int),
        ("state", C.c_void_p), ("mean", C.c_void_p),
        ("noise", C.c_void_p), ("noise_sk", c_ll), ("noise_st", c_ll), ("noise_sj", c_ll),
        ("costs", C.c_void_p), ("costs_sk", c_ll), ("costs_st", c_ll),
        ("actions", C.c_void_p), ("act_sk", c_ll), ("act_st", c_ll), ("act_sj", c_ll),
        ("qv_traj", C.c_void_p), ("next_obs", C.c_void_p), ("ncon", C.c_void_p),
        ("noise_cov", C.c_void_p), ("noise_seed", C.c_ulonglong), ("noise_offset", C.c_ulonglong),
        ("noise_step_ptr", C.c_void_p),
        ("noise_beta0", C.c_double), ("noise_beta1", C.c_double), ("noise_beta2", C.c_double),
        ("noise_k_offset", c_ll), ("noise_K_global", c_ll), ("noise_zero_last", C.c_int), ("closed_loop", C.c_int),
    ]


class MjbError(RuntimeError):
    pass


_lib = None


def lib():
    """Load the extension; raise loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MjbError(
                "mjmpc_b200: CUDA extension %s not found -- run `python -m mjmpc_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.mjb_last_error.restype = C.c_char_p
        for name in EXPORTS:
            getattr(L, name)        # AttributeError if the build is stale
        _setup_restypes(L)
        _lib = L
    return _lib


# every symbol include/mjmpc_b200.h declares
EXPORTS = [
    "mjb_last_error", "mjb_version", "mjb_device_info", "mjb_fp64_peak",
    "mjb_model_create", "mjb_model_update", "mjb_model_n_instances", "mjb_model_destroy",
    "mjb_rollout_reacher", "mjb_rollout_split_max_k",
]


def check(rc: int):
    if rc == MJB_OK:
        return
    msg = lib().mjb_last_error().decode()
    if rc == MJB_EINVAL:
        raise ValueError(msg)
    if rc == MJB_ENOTIMPL:
        raise NotImplementedError(msg)
    raise MjbError(msg)


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# ---- argument structs of include/mjmpc_b200.h ---------------------------------------------------
class PendulumArgs(C.Structure):
    _fields_ = [
        ("K", C.c_int), ("H", C.c_int), ("particles_per_ctrl", C.c_int),
        ("state", C.c_void_p), ("mean", C.c_void_p),
        ("noise", C.c_void_p), ("noise_sk", c_ll), ("noise_st", c_ll),
        ("costs", C.c_void_p), ("costs_sk", c_ll), ("costs_st", c_ll),
        ("actions", C.c_void_p), ("act_sk", c_ll), ("act_st", c_ll),
        ("states_out", C.c_void_p),
    ]


class LqrArgs(C.Structure):
    _fields_ = [
        ("K", C.c_int), ("H", C.c_int), ("n", C.c_int), ("d", C.c_int), ("particles_per_ctrl", C.c_int),
        ("A", C.c_void_p), ("B", C.c_void_p), ("Q", C.c_void_p), ("R", C.c_void_p),
        ("state", C.c_void_p), ("mean", C.c_void_p),
        ("noise", C.c_void_p), ("noise_sk", c_ll), ("noise_st", c_ll), ("noise_sj", c_ll),
        ("costs", C.c_void_p), ("costs_sk", c_ll), ("costs_st", c_ll),
        ("actions", C.c_void_p), ("act_sk", c_ll), ("act_st", c_ll), ("act_sj", c_ll),
        ("states_out", C.c_void_p),
    ]


class TreeRolloutArgs(C.Structure):
    _fields_ = [
        ("K", C.c_int), ("H", C.c_int), ("frame_skip", C.c_int), ("particles_per_ctrl", C.c_int),
        ("particles_per_model", C.c_int),
        ("fwd_dof", C.c_int), ("obs_qpos_start", C.c_int), ("w_fwd", C.c_double), ("w_ctrl", C.c_double),
        ("state", C.c_void_p), ("mean", C.c_void_p),
        ("noise", C.c_void_p), ("noise_sk", c_ll), ("noise_st", c_ll), ("noise_sj", c_ll),
        ("costs", C.c_void_p), ("costs_sk", c_ll), ("costs_st", c_ll),
        ("actions", C.c_void_p), ("act_sk", c_ll), ("act_st", c_ll), ("act_sj", c_ll),
        ("states_out", C.c_void_p), ("next_obs", C.c_void_p), ("nefc", C.c_void_p),
    ]


class NoiseArgs(C.Structure):
    _fields_ = [
        ("K", C.c_int), ("H", C.c_int), ("d", C.c_int),
        ("k_offset", c_ll), ("K_global", c_ll),
        ("seed", C.c_ulonglong), ("offset", C.c_ulonglong), ("step_ptr", C.c_void_p),
        ("cov", C.c_void_p), ("beta0", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double),
        ("zero_last", C.c_int), ("neg_mean", C.c_void_p),
        ("out", C.c_void_p), ("out_sk", c_ll), ("out_st", c_ll), ("out_sj", c_ll),
        ("cov_stride", c_ll), ("particles_per_cov", C.c_int),
    ]


class InstancesArgs(C.Structure):
    _fields_ = [
        ("n_ctrl", C.c_int), ("K", C.c_int), ("H", C.c_int), ("d", C.c_int), ("mode", C.c_int), ("apply", C.c_int),
        ("num_elite", c_ll),
        ("costs", C.c_void_p), ("costs_sk", c_ll), ("costs_st", c_ll),
        ("actions", C.c_void_p), ("act_sk", c_ll), ("act_st", c_ll), ("act_sj", c_ll),
        ("mean", C.c_void_p), ("cov", C.c_void_p), ("gamma_seq", C.c_void_p),
        ("step_size", C.c_double), ("ids", C.c_void_p), ("value", C.c_void_p),
    ]


INST_RS, INST_CEM_DIAG, INST_CEM_FULL = 0, 1, 2


class SoftmaxArgs(C.Structure):
    _fields_ = [
        ("K", C.c_int), ("H", C.c_int), ("d", C.c_int),
        ("costs", C.c_void_p), ("costs_sk", c_ll), ("costs_st", c_ll),
        ("actions", C.c_void_p), ("act_sk", c_ll), ("act_st", c_ll), ("act_sj", c_ll),
        ("mean", C.c_void_p), ("cov", C.c_void_p), ("gamma_seq", C.c_void_p),
        ("lam", C.c_double), ("control_cost", C.c_int), ("time_based", C.c_int), ("cov_mode", C.c_int),
        ("total", C.c_void_p), ("scratch", C.c_void_p), ("partials", C.c_void_p),
        ("returns", C.c_int), ("td_lam", C.c_double), ("td_gamma", C.c_double),
        ("td_weight_seq", C.c_void_p), ("qvals", C.c_void_p), ("q_sk", c_ll), ("q_st", c_ll),
    ]


class CombineArgs(C.Structure):
    _fields_ = [
        ("H", C.c_int), ("d", C.c_int), ("n_shards", C.c_int), ("K_global", c_ll),
        ("partials", C.c_void_p), ("lam", C.c_double), ("step_size", C.c_double),
        ("time_based", C.c_int), ("cov_mode", C.c_int),
        ("mean", C.c_void_p), ("cov", C.c_void_p), ("stats", C.c_void_p),
    ]


class EliteArgs(C.Structure):
    _fields_ = [
        ("K", C.c_int), ("H", C.c_int), ("d", C.c_int),
        ("flags", C.c_void_p),
        ("actions", C.c_void_p), ("act_sk", c_ll), ("act_st", c_ll), ("act_sj", c_ll),
        ("mean", C.c_void_p), ("mu", C.c_void_p), ("scratch", C.c_void_p), ("partial", C.c_void_p),
    ]


class EliteCombineArgs(C.Structure):
    _fields_ = [
        ("H", C.c_int), ("d", C.c_int), ("n_shards", C.c_int), ("full_cov", C.c_int),
        ("partial1", C.c_void_p), ("partial2", C.c_void_p), ("step_size", C.c_double),
        ("mu", C.c_void_p), ("mean", C.c_void_p), ("cov", C.c_void_p),
    ]


class MppiBatchedArgs(C.Structure):
    _fields_ = [
        ("n_ctrl", C.c_int), ("K", C.c_int), ("H", C.c_int), ("d", C.c_int),
        ("costs", C.c_void_p), ("costs_sk", c_ll), ("costs_st", c_ll),
        ("actions", C.c_void_p), ("act_sk", c_ll), ("act_st", c_ll), ("act_sj", c_ll),
        ("mean", C.c_void_p), ("cov", C.c_void_p), ("gamma_seq", C.c_void_p),
        ("lam", C.c_double), ("step_size", C.c_double), ("control_cost", C.c_int), ("value", C.c_void_p),
    ]


class PfBatchedArgs(C.Structure):
    _fields_ = [
        ("n_ctrl", C.c_int), ("K", C.c_int), ("H", C.c_int), ("d", C.c_int),
        ("costs", C.c_void_p), ("costs_sk", c_ll), ("costs_st", c_ll),
        ("samples", C.c_void_p), ("s_sk", c_ll), ("s_st", c_ll), ("s_sj", c_ll),
        ("gamma_seq", C.c_void_p), ("lam", C.c_double), ("r", C.c_void_p),
        ("weights", C.c_void_p), ("idx", C.c_void_p),
        ("out", C.c_void_p), ("o_sk", c_ll), ("o_st", c_ll), ("o_sj", c_ll),
        ("mean", C.c_void_p),
    ]


class MpcStepArgs(C.Structure):
    _fields_ = [
        ("n_iters", C.c_int),
        ("noise", C.POINTER(NoiseArgs)), ("noise_next", C.POINTER(NoiseArgs)), ("model", C.c_void_p),
        ("rollout", C.POINTER(RolloutArgs)),
        ("softmax", C.POINTER(SoftmaxArgs)), ("combine", C.POINTER(CombineArgs)),
        ("peer_bufs_dev", C.c_void_p), ("rank", C.c_int), ("seq", C.c_ulonglong),
        ("action_out", C.c_void_p), ("shift", C.c_int), ("base_action", C.c_int), ("cov_shift_beta", C.c_double),
    ]


COV_NONE, COV_DIAG, COV_FULL = 0, 1, 2
RETURNS_CTG, RETURNS_TD_LAMBDA = 0, 1
BASE_ACTIONS = {"null": 0, "repeat": 1, "random": 2}

EXPORTS += [
    "mjb_rollout_pendulum", "mjb_rollout_lqr", "mjb_tree_model_create", "mjb_tree_model_destroy", "mjb_tree_layout",
    "mjb_rollout_tree", "mjb_tree_use_planar", "mjb_generate_noise", "mjb_cost_to_go",
    "mjb_softmax_scratch_doubles", "mjb_softmax_partial_doubles", "mjb_softmax_partials", "mjb_softmax_update_fused",
    "mjb_instances_update_batched", "mjb_cov_add_diag_batched",
    "mjb_softmax_combine", "mjb_softmax_exchange_combine", "mjb_softmax_weights",
    "mjb_select_elites", "mjb_argmin", "mjb_elite_scratch_doubles", "mjb_elite_moments1",
    "mjb_elite_moments2", "mjb_elite_combine", "mjb_blend_best",
    "mjb_resample_indices", "mjb_gather_particles", "mjb_particle_mean", "mjb_particle_sub_mean",
    "mjb_shift_mean", "mjb_shift_mean_batched", "mjb_cov_add_diag", "mjb_pf_shift", "mjb_mppi_update_batched",
    "mjb_pf_update_batched", "mjb_particle_sub_mean_batched", "mjb_softmax_mpc_step",
]


def _setup_restypes(L):
    L.mjb_tree_model_create.restype = C.c_void_p
    L.mjb_tree_model_destroy.restype = None
    L.mjb_tree_model_destroy.argtypes = [C.c_void_p]
    L.mjb_tree_layout.restype = None
    L.mjb_rollout_tree.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.mjb_softmax_scratch_doubles.restype = c_ll
    L.mjb_elite_scratch_doubles.restype = c_ll
