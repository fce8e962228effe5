// TEST HARNESS (CPU): runs the product's one-thread-per-particle rollout KERNEL -- the actual source
// rollout_reacher.cu (K1 wrapper: noise prefetch, fused noise, trajectory / observation outputs, closed-loop
// policy) -- on the host, one simulated thread after another, so that changes made without a GPU at hand are
// executed before they reach one.  (The noise kernel K2 became a cooperating-thread kernel in round 2: it runs on
// the block emulator, libmjmpc_b200_emu.so, like the update kernels.)  Not a product path and not a CUDA emulator:
// it only covers kernels whose threads do not communicate (thread 0's shared-memory prologue runs first
// because threads are simulated in index order); MUFU approximations are replaced by libm.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>          // host-side types only (g++: the execution-space macros expand to nothing)

#define MJB_HOST_EMU 1
#undef __shared__
#define __shared__ static           // one simulated block at a time
#undef __constant__
#define __constant__ static

namespace {
struct Idx { unsigned x, y, z; };
}
static Idx threadIdx, blockIdx, blockDim, gridDim;
static inline void __syncthreads() {}
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
#define __log2f(x) log2f(x)          // glibc declares these names for its own internals
#define __sinf(x) sinf(x)
#define __cosf(x) cosf(x)
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif

#include "../../mjmpc_b200/csrc/rollout_reacher.cu"

// ---- K1 ------------------------------------------------------------------------------------------------
// params167: n_inst x MJB_MODEL_NPARAM host blocks.  One instance: constant-bank instantiation (as the library
// chooses), several: global-memory instantiation.
extern "C" int emu_rollout_reacher(const double* params167, int n_inst, const mjb_rollout_args* a) {
    static double dev[64 * CH_NDEV];
    if (n_inst < 1 || n_inst > 64) return 1;
    for (int i = 0; i < n_inst; i++) {
        double* P = dev + (size_t)i * CH_NDEV;
        for (int k = 0; k < CH_NPARAM; k++) P[k] = params167[(size_t)i * CH_NPARAM + k];
        mjb_derive_params(P);
        if (!mjb_params_fit_sawyer(P)) return 2;
    }
    for (int k = 0; k < CH_NDEV; k++) mjb::c_params[k] = dev[k];
    const bool extra = a->qv_traj || a->next_obs || a->ncon || a->closed_loop;
    const bool fused = a->noise_cov != nullptr;
    const unsigned block = MJB_ROLLOUT_BLOCK, grid = (a->K + block - 1) / block;
    blockDim = {block, 1, 1}; gridDim = {grid, 1, 1};
    for (unsigned b = 0; b < grid; b++)
        for (unsigned t = 0; t < block; t++) {
            blockIdx = {b, 0, 0}; threadIdx = {t, 0, 0};
#define EMU_LAUNCH(P, E, F) mjb::rollout_reacher_kernel<mjb::SawyerTraits, mjb::P, E, F>(dev, n_inst, *a)
#define EMU_PICK(P) do { if (extra) { if (fused) EMU_LAUNCH(P, true, true); else EMU_LAUNCH(P, true, false); } \
                         else { if (fused) EMU_LAUNCH(P, false, true); else EMU_LAUNCH(P, false, false); } } while (0)
            if (n_inst == 1) EMU_PICK(ConstParams); else EMU_PICK(GlobalParams);
        }
    return 0;
}
