// K2: action-noise sampling.  Counter-based Philox4x32-10 Gaussians, covariance factor transform
// and the reference's autoregressive filter along the horizon, one thread per particle, written
// in whatever layout the caller's strides describe (particle-minor = coalesced).
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

template <int D>
__global__ void __launch_bounds__(128) noise_kernel(mjb_noise_args a) {
    __shared__ double L[MJB_MAXD][MJB_MAXD];
    if (threadIdx.x == 0) noise_chol<D>(a.cov, L);
    __syncthreads();
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.K) return;
    const unsigned long long gk = (unsigned long long)(a.k_offset + k);
    Philox ph{(unsigned)a.seed, (unsigned)(a.seed >> 32)};
    // counter = (global particle, step counter, t, stream id << 8 | pair index); key = seed
    const unsigned step_ctr = a.step_ptr ? (unsigned)(*a.step_ptr) : (unsigned)a.offset;
    const unsigned tag_hi = (unsigned)(a.offset >> 32) << 8;
    double e1[D], e2[D];   // filtered history t-1, t-2
#pragma unroll
    for (int j = 0; j < D; j++) { e1[j] = 0.0; e2[j] = 0.0; }
    const bool zero_seq = a.zero_last && (long long)gk == a.K_global - 1;
    for (int t = 0; t < a.H; t++) {
        double z[(D + 3) / 4 * 4];
        noise_normals<D>(ph, gk, step_ctr, tag_hi, t, z);
        double e[D];
        noise_shape<D>(L, z, t, a.beta0, a.beta1, a.beta2, e1, e2, e);
#pragma unroll
        for (int j = 0; j < D; j++) {
            e2[j] = e1[j]; e1[j] = e[j];
            const double o = zero_seq ? -a.neg_mean[t * D + j] : e[j];
            a.out[k * a.out_sk + t * a.out_st + j * a.out_sj] = o;
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
    if (a->K == 0) return MJB_OK;
    const int block = 128, grid = (a->K + block - 1) / block;
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
