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
};
