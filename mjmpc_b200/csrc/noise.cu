// K2: action-noise sampling.  Counter-based Philox4x32-10 Gaussians, covariance factor transform
// and the reference's autoregressive filter along the horizon, written in whatever layout the
// caller's strides describe (particle-minor = coalesced).
// Replaces generate_noise (mjmpc/utils/control_utils.py:24-34) as called by
// OLGaussianMPC.sample_noise (mjmpc/control/olgaussian_mpc.py:88-93), PFMPC.__init__/_shift
// (mjmpc/control/particle_filter_controller.py:69-71,136-139) and the use_zero_control_seq
// overwrite of the last particle (olgaussian_mpc.py:110-111).
//
// The reference draws from numpy's global MT19937 stream reseeded with seed_val + num_steps;
// a counter-based generator cannot reproduce those samples, so parity for everything downstream
// is defined on an injected noise tensor and this kernel is validated on its statistics, its
// determinism and its independence from how particles are sharded (counters are keyed by the
// GLOBAL particle index).
#include "common.h"
#include "philox_noise.cuh"

namespace mjb {

// Tile = 32 particles x 16 horizon steps per pass.  Phase 1 (one thread per (particle, step) pair, two pairs per
// thread): Philox, Box-Muller and the covariance factor -- the 250 integer / SFU instructions per pair that made the
// one-thread-per-particle form of this kernel latency-bound at 34 % of HBM speed -- run for all 512 pairs of the
// tile at once; the shaped normals go to shared memory.  Phase 2 (one thread per (particle, action dimension)):
// the autoregressive recursion along the horizon, the only sequential part (3 FP64 operations per sample), history
// in registers across passes, stores coalesced over particles.  Same operations in the same order per sample as
// noise_normals / noise_shape: bit-identical to the fused in-rollout generation.
#define MJB_NZ_KP 32          // particles per block
#define MJB_NZ_TT 16          // horizon steps per pass
#define MJB_NZ_THREADS 256

template <int D>
__global__ void __launch_bounds__(MJB_NZ_THREADS) noise_kernel(mjb_noise_args a) {
    __shared__ double L[MJB_MAXD][MJB_MAXD];
    __shared__ double sh[MJB_NZ_TT][D][MJB_NZ_KP];      // shaped normals (L z) of the pass
    if (a.cov_stride == 0) {
        if (threadIdx.x == 0) noise_chol<D>(a.cov, L);  // ~2 us of dependent sqrt / divide: once per (persistent) block
        __syncthreads();
    }
    Philox ph{(unsigned)a.seed, (unsigned)(a.seed >> 32)};
    // counter = (global particle, step counter, t, stream id << 8 | pair index); key = seed
    const unsigned step_ctr = (a.step_ptr ? (unsigned)(*a.step_ptr) : 0u) + (unsigned)a.offset;
    const unsigned tag_hi = (unsigned)(a.offset >> 32) << 8;
    // phase-2 identity of this thread: particle kl2, dimension j2 (threads beyond 32 * D idle in phase 2)
    const int kl2 = threadIdx.x % MJB_NZ_KP, j2 = threadIdx.x / MJB_NZ_KP;
    const long long ntiles = (a.K + MJB_NZ_KP - 1) / MJB_NZ_KP;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long kb = tile * MJB_NZ_KP;
    if (a.cov_stride != 0) {                            // batched instances with their own covariance: one factor per tile
        __syncthreads();
        if (threadIdx.x == 0) noise_chol<D>(a.cov + ((a.k_offset + kb) / a.particles_per_cov) * a.cov_stride, L);
        __syncthreads();
    }
    const long long k2 = kb + kl2;
    const bool act2 = j2 < D && k2 < a.K;
    const bool zero_seq = a.zero_last && (a.k_offset + k2) == a.K_global - 1;
    double e1 = 0.0, e2 = 0.0;                          // filtered history t-1, t-2 of (k2, j2)
    for (int t0 = 0; t0 < a.H; t0 += MJB_NZ_TT) {
        // ---- phase 1
#pragma unroll
        for (int r = 0; r < MJB_NZ_KP * MJB_NZ_TT / MJB_NZ_THREADS; r++) {
            const int pair = threadIdx.x + r * MJB_NZ_THREADS;
            const int kl = pair % MJB_NZ_KP, tl = pair / MJB_NZ_KP;
            const long long k = kb + kl;
            const int t = t0 + tl;
            if (k < a.K && t < a.H) {
                double z[(D + 3) / 4 * 4];
                noise_normals<D>(ph, (unsigned long long)(a.k_offset + k), step_ctr, tag_hi, t, z);
#pragma unroll
                for (int j = 0; j < D; j++) {
                    double s = 0.0;
#pragma unroll
                    for (int i = 0; i <= j; i++) s = fma(L[j][i], z[i], s);
                    sh[tl][j][kl] = s;
                }
            }
        }
        __syncthreads();
        // ---- phase 2
        if (act2) {
            const int tn = (a.H - t0) < MJB_NZ_TT ? (a.H - t0) : MJB_NZ_TT;
            double* o = a.out + k2 * a.out_sk + (long long)t0 * a.out_st + j2 * a.out_sj;
            for (int tl = 0; tl < tn; tl++) {
                const int t = t0 + tl;
                const double s = sh[tl][j2][kl2];
                const double e = t >= 2 ? fma(a.beta2, e2, fma(a.beta1, e1, a.beta0 * s)) : s;
                e2 = e1; e1 = e;
                *o = zero_seq ? -a.neg_mean[t * D + j2] : e;
                o += a.out_st;
            }
        }
        __syncthreads();
    }
    }
}

}  // namespace mjb

#ifndef MJB_HOST_EMU
extern "C" int mjb_generate_noise(const mjb_noise_args* a, void* stream) {
    MJB_REQUIRE(a && a->cov && a->out, "mjb_generate_noise: null pointer");
    MJB_REQUIRE(a->K >= 0 && a->H >= 1, "mjb_generate_noise: bad shape K=%d H=%d", a->K, a->H);
    MJB_REQUIRE(a->d >= 1 && a->d <= MJB_MAXD, "mjb_generate_noise: d=%d not in 1..%d", a->d, MJB_MAXD);
    MJB_REQUIRE(!a->zero_last || a->neg_mean, "mjb_generate_noise: zero_last needs the mean sequence");
    MJB_REQUIRE(a->cov_stride == 0 || (a->particles_per_cov >= 32 && a->particles_per_cov % 32 == 0 && a->k_offset % 32 == 0),
                "mjb_generate_noise: per-instance covariances need particles_per_cov (and k_offset) in multiples of 32");
    if (a->K == 0) return MJB_OK;
    // persistent blocks striding over the 32-particle tiles: 4 resident blocks per SM's worth at most
    static int sms = 0;
    if (sms == 0 && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0) != cudaSuccess) sms = 148;
    const long long ntiles = ((long long)a->K + MJB_NZ_KP - 1) / MJB_NZ_KP;
    const int block = MJB_NZ_THREADS, grid = (int)(ntiles < 4ll * sms ? ntiles : 4ll * sms);
    cudaStream_t s = (cudaStream_t)stream;
    switch (a->d) {
#define MJB_CASE(D) case D: mjb::noise_kernel<D><<<grid, block, 0, s>>>(*a); break;
        MJB_CASE(1) MJB_CASE(2) MJB_CASE(3) MJB_CASE(4) MJB_CASE(5) MJB_CASE(6) MJB_CASE(7) MJB_CASE(8)
#undef MJB_CASE
    }
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}
#endif  // MJB_HOST_EMU
