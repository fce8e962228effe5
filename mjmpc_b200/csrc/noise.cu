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

namespace mjb {

#define MJB_MAXD 8

struct Philox {
    unsigned k0, k1;
    __device__ __forceinline__ void operator()(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned (&o)[4]) const {
        unsigned a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; r++) {
            const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            const unsigned n0 = hi1 ^ c1 ^ a, n2 = hi0 ^ c3 ^ b;
            c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
    }
};

// Four standard normals from the four 32-bit words of one Philox call.  The Gaussian variates are
// formed in FP32 (Box-Muller on 32-/24-bit uniforms with the SFU log / sincos: 24-bit resolution,
// |z| <= 6.7 sigma) and promoted to FP64; the covariance transform and the filter below run in FP64.
// Exploration noise does not need 53-bit variates, and the kernel is instruction-bound on the Philox
// rounds, so every generated word is used.
__device__ __forceinline__ void normal_pair32(unsigned a, unsigned b, double& z0, double& z1) {
    const float u1 = ((float)a + 0.5f) * 2.3283064365386963e-10f;    // (0,1], 2^-32
    const float u2 = (float)(b >> 8) * 5.9604644775390625e-08f;      // [0,1), 2^-24: exact in FP32
    const float rad = sqrtf(-2.0f * logf(fminf(u1, 1.0f)));
    float s, c;
    sincospif(2.0f * u2, &s, &c);
    z0 = (double)(rad * c); z1 = (double)(rad * s);
}

template <int D>
__global__ void __launch_bounds__(128) noise_kernel(mjb_noise_args a) {
    __shared__ double L[MJB_MAXD][MJB_MAXD];
    if (threadIdx.x == 0) {
        // lower Cholesky factor of cov (positive semi-definite tolerated: a non-positive pivot zeroes its column)
        for (int j = 0; j < D; j++) {
            double s = a.cov[j * D + j];
            for (int k = 0; k < j; k++) s -= L[j][k] * L[j][k];
            const double piv = s > 0.0 ? sqrt(s) : 0.0;
            L[j][j] = piv;
            for (int i = j + 1; i < D; i++) {
                double t = a.cov[i * D + j];
                for (int k = 0; k < j; k++) t -= L[i][k] * L[j][k];
                L[i][j] = piv > 0.0 ? t / piv : 0.0;
            }
            for (int i = 0; i < j; i++) L[i][j] = 0.0;
        }
    }
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
#pragma unroll
        for (int p = 0; p < (D + 3) / 4; p++) {
            unsigned r[4];
            ph((unsigned)gk, step_ctr, (unsigned)t, tag_hi | (unsigned)p, r);
            normal_pair32(r[0], r[1], z[4 * p], z[4 * p + 1]);
            normal_pair32(r[2], r[3], z[4 * p + 2], z[4 * p + 3]);
        }
        double e[D];
#pragma unroll
        for (int j = 0; j < D; j++) {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i <= j; i++) s += L[j][i] * z[i];
            e[j] = s;
        }
        if (t >= 2) {
#pragma unroll
            for (int j = 0; j < D; j++) e[j] = a.beta0 * e[j] + a.beta1 * e1[j] + a.beta2 * e2[j];
        }
#pragma unroll
        for (int j = 0; j < D; j++) {
            e2[j] = e1[j]; e1[j] = e[j];
            const double o = zero_seq ? -a.neg_mean[t * D + j] : e[j];
            a.out[k * a.out_sk + t * a.out_st + j * a.out_sj] = o;
        }
    }
}

}  // namespace mjb

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
