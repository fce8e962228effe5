"""ctypes binding of ``libmjmpc_b200.so`` (the C ABI declared in ``include/mjmpc_b200.h``).

There is no CPU fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmjmpc_b200.so")

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
        _lib = L
    return _lib


# every symbol include/mjmpc_b200.h declares
EXPORTS = [
    "mjb_last_error", "mjb_version", "mjb_device_info", "mjb_fp64_peak",
    "mjb_model_create", "mjb_model_update", "mjb_model_n_instances", "mjb_model_destroy",
    "mjb_rollout_reacher",
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
