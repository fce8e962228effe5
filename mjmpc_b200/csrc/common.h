// Shared plumbing of the C-ABI translation units: error slot, CUDA checks, model handle.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include "../../include/mjmpc_b200.h"
#include "chain_model.h"

namespace mjb {
int set_error(int code, const char* fmt, ...);
}

#define MJB_CUDA(expr)                                                                             \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return mjb::set_error(MJB_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                  __FILE__, __LINE__);                                             \
    } while (0)
#define MJB_REQUIRE(cond, ...)                                   \
    do {                                                         \
        if (!(cond)) return mjb::set_error(MJB_EINVAL, __VA_ARGS__); \
    } while (0)

struct mjb_model {
    int device;
    int n_instances;
    int fits_sawyer;      // every instance has the structural zeros SawyerTraits assumes
    double* d_params;     // n_instances x CH_NDEV (device)
    double* h_params;     // host mirror
    unsigned long long serial;  // bumped on every update (constant-bank cache key)
    int uniform_frame_skip;     // every instance has the same frame_skip (the role-split rollout needs block-uniform loops)
};

namespace mjb {
// The rollout kernels read a single-instance model from ONE per-device __constant__ bank (rollout_reacher.cu).
// The bank belongs to one live model at a time: the first single-instance model that launches a rollout of
// >= 64 particles claims it and keeps it until it is destroyed; every other model reads its parameters from
// global memory.  So a captured CUDA graph never sees another model's constants, and an update of the owner
// is uploaded at update time, not lazily at the next launch.
int const_bank_on_update(mjb_model* m);      // called by mjb_model_update after the host/device blocks changed
void const_bank_release(const mjb_model* m); // called by mjb_model_destroy
}
