// Counter-based Gaussian noise shared by the standalone noise kernel (K2, noise.cu) and the fused
// in-rollout generation (rollout_reacher.cu): both must produce bit-identical samples for the same
// (seed, step, global particle, t), so the generator lives here once.
#pragma once

namespace mjb {

#ifndef MJB_MAXD
#define MJB_MAXD 8
#endif

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

// Two standard normals from two 32-bit words of a Philox call (Box-Muller).  The Gaussian variates are formed
// in FP32 on the special-function unit -- lg2.approx for the radius, sqrt.approx, sin/cos.approx on an angle
// shifted into [-pi, pi) where their absolute error is 2^-21 -- and promoted to FP64; the covariance transform
// and the filter below run in FP64.  24-bit angle, 32-bit radius uniform: |z| <= 6.7 sigma, absolute error of a
// variate ~1e-6 sigma.  Exploration noise does not need 53-bit variates, and the kernel is instruction-bound:
// with the libm-accurate logf / sincospif / sqrtf this function was 330 of the 590 instructions per
// (particle, step); every generated Philox word is used.
__device__ __forceinline__ void normal_pair32(unsigned a, unsigned b, double& z0, double& z1) {
    const float u1 = ((float)a + 0.5f) * 2.3283064365386963e-10f;    // (0,1], 2^-32
    const float u2 = (float)(b >> 8) * 5.9604644775390625e-08f;      // [0,1), 2^-24: exact in FP32
    // -2 ln u1; lg2.approx has an ABSOLUTE error of ~2^-22, so just below u1 = 1 its sign can be wrong:
    // clamp, or the square root of a tiny negative number would put a NaN into one sample in ~10^7
    const float r2 = fmaxf(-1.3862943611198906f * __log2f(fminf(u1, 1.0f)), 0.0f);
    float rad;
#if defined(__CUDA_ARCH__)
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(r2));
#else
    rad = sqrtf(r2);                   // host emulation of the kernels (tests/hostcheck/kernel_emu.cpp)
#endif
    // cos(2 pi u2) = -cos(2 pi u2 - pi): the shifted angle is uniform on [-pi, pi)
    const float ang = fmaf(u2, 6.2831853071795865f, -3.1415926535897932f);
    z0 = (double)(-rad * __cosf(ang)); z1 = (double)(-rad * __sinf(ang));
}


// z[0..D) standard normals of (particle gk, step counter, horizon index t): ceil(D/4) Philox calls
template <int D>
__device__ __forceinline__ void noise_normals(const Philox& ph, unsigned long long gk, unsigned step_ctr, unsigned tag_hi,
                                              int t, double (&z)[(D + 3) / 4 * 4]) {
#pragma unroll
    for (int p = 0; p < (D + 3) / 4; p++) {
        unsigned r[4];
        ph((unsigned)gk, step_ctr, (unsigned)t, tag_hi | (unsigned)p, r);
        normal_pair32(r[0], r[1], z[4 * p], z[4 * p + 1]);
        normal_pair32(r[2], r[3], z[4 * p + 2], z[4 * p + 3]);
    }
}

// eps_t = filter(L z): covariance factor transform, then the reference's autoregressive recursion
// eps[i] = b0 eps[i] + b1 eps[i-1] + b2 eps[i-2] for i >= 2 (control_utils.py:31-33) on the already
// filtered history (h1, h2).  Explicit FMAs so both kernels round identically.
template <int D>
__device__ __forceinline__ void noise_shape(const double (*L)[MJB_MAXD], const double (&z)[(D + 3) / 4 * 4], int t,
                                            double b0, double b1, double b2, const double (&h1)[D],
                                            const double (&h2)[D], double (&e)[D]) {
#pragma unroll
    for (int j = 0; j < D; j++) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i <= j; i++) s = fma(L[j][i], z[i], s);
        e[j] = t >= 2 ? fma(b2, h2[j], fma(b1, h1[j], b0 * s)) : s;
    }
}

// lower Cholesky factor of a (D,D) covariance (positive semi-definite tolerated: a non-positive pivot
// zeroes its column); one thread
template <int D>
__device__ inline void noise_chol(const double* __restrict__ cov, double (*L)[MJB_MAXD]) {
    for (int j = 0; j < D; j++) {
        double s = cov[j * D + j];
        for (int k = 0; k < j; k++) s -= L[j][k] * L[j][k];
        const double piv = s > 0.0 ? sqrt(s) : 0.0;
        L[j][j] = piv;
        for (int i = j + 1; i < D; i++) {
            double t = cov[i * D + j];
            for (int k = 0; k < j; k++) t -= L[i][k] * L[j][k];
            L[i][j] = piv > 0.0 ? t / piv : 0.0;
        }
        for (int i = 0; i < j; i++) L[i][j] = 0.0;
    }
}

}  // namespace mjb
