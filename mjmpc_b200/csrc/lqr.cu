// K10: batched rollout of the reference's linear-quadratic toy env with the step cost fused in, one thread per
// particle.  Replaces LQREnv.step (mjmpc/envs/basic/lqr.py:31-35: cost = x'Qx + u'Ru on the PRE-step state,
// x <- Ax + Bu, no clipping, never done) under GymEnvWrapper.rollout (mjmpc/envs/gym_env_wrapper.py:125-153).
// A, B, Q, R are staged in shared memory once per block; the state lives in registers (compile-time indexed
// arrays bounded by MJB_LQR_MAXN / MJB_LQR_MAXD, runtime sizes guard the loops).  Sums run in index order;
// numpy's BLAS may order them differently, so parity is stated at 1e-12 relative, not bit-exact.
#include "common.h"

namespace mjb {
__global__ void __launch_bounds__(128) rollout_lqr_kernel(mjb_lqr_args a) {
    __shared__ double sA[MJB_LQR_MAXN * MJB_LQR_MAXN], sB[MJB_LQR_MAXN * MJB_LQR_MAXD], sQ[MJB_LQR_MAXN * MJB_LQR_MAXN],
        sR[MJB_LQR_MAXD * MJB_LQR_MAXD];
    const int n = a.n, d = a.d;
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) { sA[i] = a.A[i]; sQ[i] = a.Q[i]; }
    for (int i = threadIdx.x; i < n * d; i += blockDim.x) sB[i] = a.B[i];
    for (int i = threadIdx.x; i < d * d; i += blockDim.x) sR[i] = a.R[i];
    __syncthreads();
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.K) return;
    const int ctrl = (int)(k / a.particles_per_ctrl);
    double x[MJB_LQR_MAXN], u[MJB_LQR_MAXD];
#pragma unroll
    for (int i = 0; i < MJB_LQR_MAXN; i++) x[i] = i < n ? a.state[(long long)ctrl * n + i] : 0.0;
    const double* mean = a.mean + (long long)ctrl * a.H * d;
    for (int t = 0; t < a.H; t++) {
#pragma unroll
        for (int j = 0; j < MJB_LQR_MAXD; j++) {
            u[j] = 0.0;
            if (j < d) {
                u[j] = mean[t * d + j];
                if (a.noise) u[j] += a.noise[k * a.noise_sk + t * a.noise_st + j * a.noise_sj];
                if (a.actions) a.actions[k * a.act_sk + t * a.act_st + j * a.act_sj] = u[j];
            }
        }
        // cost = (x'Q) x + (u'R) u, row vector first like state.T.dot(Q).dot(state)
        double cost = 0.0;
#pragma unroll
        for (int j = 0; j < MJB_LQR_MAXN; j++) {
            if (j < n) {
                double xq = 0.0;
#pragma unroll
                for (int i = 0; i < MJB_LQR_MAXN; i++) if (i < n) xq += x[i] * sQ[i * n + j];
                cost += xq * x[j];
            }
        }
        double cu = 0.0;
#pragma unroll
        for (int j = 0; j < MJB_LQR_MAXD; j++) {
            if (j < d) {
                double ur = 0.0;
#pragma unroll
                for (int i = 0; i < MJB_LQR_MAXD; i++) if (i < d) ur += u[i] * sR[i * d + j];
                cu += ur * u[j];
            }
        }
        a.costs[k * a.costs_sk + t * a.costs_st] = cost + cu;
        double xn[MJB_LQR_MAXN];
#pragma unroll
        for (int i = 0; i < MJB_LQR_MAXN; i++) {
            double ax = 0.0, bu = 0.0;
            if (i < n) {
#pragma unroll
                for (int j = 0; j < MJB_LQR_MAXN; j++) if (j < n) ax += sA[i * n + j] * x[j];
#pragma unroll
                for (int j = 0; j < MJB_LQR_MAXD; j++) if (j < d) bu += sB[i * d + j] * u[j];
            }
            xn[i] = ax + bu;
        }
#pragma unroll
        for (int i = 0; i < MJB_LQR_MAXN; i++) {
            x[i] = xn[i];
            if (a.states_out && i < n) a.states_out[(k * a.H + t) * n + i] = x[i];
        }
    }
}
}  // namespace mjb

#ifndef MJB_HOST_EMU
extern "C" int mjb_rollout_lqr(const mjb_lqr_args* a, void* stream) {
    MJB_REQUIRE(a && a->A && a->B && a->Q && a->R && a->state && a->mean && a->costs, "mjb_rollout_lqr: null pointer");
    MJB_REQUIRE(a->K >= 1 && a->H >= 1, "mjb_rollout_lqr: K and H must be positive");
    MJB_REQUIRE(a->n >= 1 && a->n <= MJB_LQR_MAXN && a->d >= 1 && a->d <= MJB_LQR_MAXD,
                "mjb_rollout_lqr: d_state=%d / d_action=%d not in 1..%d / 1..%d", a->n, a->d, MJB_LQR_MAXN, MJB_LQR_MAXD);
    MJB_REQUIRE(a->particles_per_ctrl >= 1 && a->K % a->particles_per_ctrl == 0,
                "Number of particles must be divisible by number of controllers");
    mjb::rollout_lqr_kernel<<<(a->K + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*a);
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}
#endif  // MJB_HOST_EMU
