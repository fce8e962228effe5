// Register-pressure probe for a role-split rollout (two cooperating warps per 32 particles): the two roles
// compiled as separate kernels so that ptxas reports what each needs.  Not product code, never launched.
//   nvcc -O3 -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -DOCC=8 -Xptxas -v -c tools/role_probe.cu
// Result (CUDA 12.9): role B (mass matrix + every factor/solve) 142 registers unconstrained, 126 with 0 spills at
// 8 blocks/SM; role A (bias forces, limit rows, active-set logic) 168 unconstrained, 128 with 136 B of spills.
#include "../mjmpc_b200/csrc/chain_dynamics.cuh"
#include "../mjmpc_b200/csrc/common.h"
namespace mjb {
__constant__ double c_params[CH_NDEV];
struct ConstParams { __device__ __forceinline__ double operator[](int i) const { return c_params[i]; } };
#define PB 32      // particles per block (columns of the scratch)
struct PairScratch {
    double* base;
    __device__ __forceinline__ double ld(int slot) const { return base[slot * PB]; }
    __device__ __forceinline__ void st(int slot, double v) { base[slot * PB] = v; }
};
enum { SX_RHS = SC_NSLOT, SX_DADD = SC_NSLOT + 7, SX_X = SC_NSLOT + 14, SX_FLAG = SC_NSLOT + 21, SX_N = SC_NSLOT + 22 };

#ifndef OCC
#define OCC 8
#endif
// role A: bias forces, limit rows, active-set logic; factor/solves are done by the partner (x read from scratch)
template <class T>
__global__ void __launch_bounds__(64, OCC) roleA(const double* st, int nsub, double* out) {
    __shared__ double smem[SX_N * PB];
    const int lane = threadIdx.x & 31;
    PairScratch sc{smem + lane};
    ConstParams prm;
    double q[7], v[7], sn[7], cs[7];
#pragma unroll
    for (int j = 0; j < 7; j++) { q[j] = st[j] + 1e-3 * threadIdx.x; v[j] = st[7 + j]; }
#pragma unroll 1
    for (int s = 0; s < nsub; s++) {
#pragma unroll
        for (int j = 0; j < 7; j++) sincos_joint(q[j], sn[j], cs[j]);
        double f[7];
        chain_mass_bias<T, 1>(prm, sc, sn, cs, v, f);
#pragma unroll
        for (int j = 0; j < 7; j++) f[j] = sc.ld(SC_U + j) - prm[CH_DAMPING + j] * v[j] - f[j];
        Rows R;
        const bool any = make_rows<T>(prm, sc, q, v, sn, cs, R);
        double dadd[7], radd[7], x[7];
        unsigned act = 0;
        int phase = any ? 0 : 1;
#pragma unroll
        for (int j = 0; j < 7; j++) {
            const double bm = R.bs[j] * sc.ld(sc_m(j, j));
            const bool on = any && sflip(f[j], R.sgn(j)) < bm;
            act |= on ? (1u << j) : 0u;
            dadd[j] = any ? (on ? R.D[j] : 0.0) : prm[CH_HDAMP + j];
            radd[j] = sflip((on ? R.D[j] : 0.0) * R.bs[j], R.sgn(j));
        }
        int again;
#pragma unroll 1
        do {
            MJB_OPAQUE(phase);
#pragma unroll
            for (int j = 0; j < 7; j++) { sc.st(SX_RHS + j, f[j] + radd[j]); sc.st(SX_DADD + j, dadd[j]); }
            __syncthreads();          // partner factors and solves
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 7; j++) x[j] = sc.ld(SX_X + j);
            again = 0;
            if (phase == 0) {
                unsigned actx = 0;
#pragma unroll
                for (int j = 0; j < 7; j++) actx |= (sflip(x[j], R.sgn(j)) < R.bs[j]) ? (1u << j) : 0u;
                const bool ok = actx == act;
#pragma unroll
                for (int j = 0; j < 7; j++) {
                    const double De = ((actx >> j) & 1u) ? R.D[j] : 0.0;
                    dadd[j] = ok ? prm[CH_HDAMP + j] : De;
                    radd[j] = sflip(De * (R.bs[j] - (ok ? sflip(x[j], R.sgn(j)) : 0.0)), R.sgn(j));
                }
                act = actx;
                phase = ok ? 1 : 0;
                again = 1;
            }
            MJB_OPAQUE(again);
        } while (again);
        const double h = prm[CS_TIMESTEP];
#pragma unroll
        for (int j = 0; j < 7; j++) { v[j] += h * x[j]; q[j] += h * v[j]; }
    }
#pragma unroll
    for (int j = 0; j < 7; j++) out[(blockIdx.x * 64 + threadIdx.x) * 14 + j] = q[j] + v[j];
}

// role B: mass matrix, then every factor/solve the partner asks for
template <class T>
__global__ void __launch_bounds__(64, OCC) roleB(const double* st, int nsub, double* out) {
    __shared__ double smem[SX_N * PB];
    const int lane = threadIdx.x & 31;
    PairScratch sc{smem + lane};
    ConstParams prm;
    double q[7], v[7], sn[7], cs[7];
#pragma unroll
    for (int j = 0; j < 7; j++) { q[j] = st[j] + 1e-3 * threadIdx.x; v[j] = st[7 + j]; }
#pragma unroll 1
    for (int s = 0; s < nsub; s++) {
#pragma unroll
        for (int j = 0; j < 7; j++) sincos_joint(q[j], sn[j], cs[j]);
        double dummy[7];
        chain_mass_bias<T, 2>(prm, sc, sn, cs, v, dummy);
        double x[7];
        int again;
#pragma unroll 1
        do {
            __syncthreads();
            double H[7][7], dinv[7];
#pragma unroll
            for (int i = 0; i < 7; i++) {
#pragma unroll
                for (int j = 0; j < i; j++) H[i][j] = sc.ld(sc_m(i, j));
                H[i][i] = sc.ld(sc_m(i, i)) + sc.ld(SX_DADD + i);
                x[i] = sc.ld(SX_RHS + i);
            }
            ldl7(H, dinv);
            ldl7_solve(H, dinv, x);
#pragma unroll
            for (int j = 0; j < 7; j++) sc.st(SX_X + j, x[j]);
            __syncthreads();
            again = (int)sc.ld(SX_FLAG);
            MJB_OPAQUE(again);
        } while (again);
        const double h = prm[CS_TIMESTEP];
#pragma unroll
        for (int j = 0; j < 7; j++) { v[j] += h * x[j]; q[j] += h * v[j]; }
    }
#pragma unroll
    for (int j = 0; j < 7; j++) out[(blockIdx.x * 64 + threadIdx.x) * 14 + j] = q[j] + v[j];
}
template __global__ void roleA<SawyerTraits>(const double*, int, double*);
template __global__ void roleB<SawyerTraits>(const double*, int, double*);
}
