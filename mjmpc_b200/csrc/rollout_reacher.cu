// K1: batched reacher_7dof rollout with the step cost fused in.  One thread per particle, the
// whole horizon in one launch, state in registers.  Replaces the reference's per-particle
// env-copy loop: GymEnvWrapper.rollout (mjmpc/envs/gym_env_wrapper.py:125-153) ->
// Reacher7DOFEnv.step (mjmpc/envs/basic/reacher_env.py:29-39) -> MuJoCo mj_step x frame_skip,
// fanned out by SubprocVecEnv.rollout (mjmpc/envs/vec_env/subproc_vec_env.py:161-186).
#include "chain_dynamics.cuh"
#include "common.h"

namespace mjb {

struct GlobalParams {
    const double* __restrict__ p;
    __device__ __forceinline__ double operator[](int i) const { return __ldg(p + i); }
};

// 64-thread blocks, 4 per SM: 256 resident particles per SM at up to 255 registers each; the
// small block keeps the second wave of a K=65536 launch evenly spread over the 148 SMs.
template <class T>
__global__ void __launch_bounds__(64, 4) rollout_reacher_kernel(const double* __restrict__ params, int n_inst,
                                                                mjb_rollout_args a) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.K) return;
    const int ctrl = k / a.particles_per_ctrl;
    const int inst = (k / a.particles_per_model) % n_inst;
    GlobalParams prm{params + (size_t)inst * CH_NDEV};
    const double* __restrict__ st = a.state + (size_t)ctrl * MJB_STATE_DIM;
    const double* __restrict__ mean = a.mean + (size_t)ctrl * a.H * 7;
    double q[7], v[7], u[7], sn[7], cs[7];
#pragma unroll
    for (int j = 0; j < 7; j++) { q[j] = __ldg(st + j); v[j] = __ldg(st + 7 + j); }
    const V3 target = {__ldg(st + 14), __ldg(st + 15), __ldg(st + 16)};
    const V3 hand_local = {prm[CS_HAND], prm[CS_HAND + 1], prm[CS_HAND + 2]};
    const int fs = (int)prm[CS_FRAME_SKIP];
    int nc = 0;
    for (int t = 0; t < a.H; t++) {
#pragma unroll
        for (int j = 0; j < 7; j++) {
            double x = __ldg(mean + t * 7 + j);
            if (a.noise) x += __ldg(a.noise + k * a.noise_sk + t * a.noise_st + j * a.noise_sj);
            u[j] = x;
            if (a.actions) a.actions[k * a.act_sk + t * a.act_st + j * a.act_sj] = x;
        }
        V3 hand = {0.0, 0.0, 0.0};
        for (int s = 0; s < fs; s++) {
#pragma unroll
            for (int j = 0; j < 7; j++) sincos(q[j], &sn[j], &cs[j]);
            // data.site_xpos read after mj_step is the one computed by the LAST forward pass, i.e. at
            // the state before the last substep's integration (reacher_env.py:31-35).
            if (s == fs - 1) hand = chain_point_world<T>(prm, sn, cs, hand_local);
            nc += chain_substep<T>(prm, q, v, sn, cs, u) ? 1 : 0;
        }
        a.costs[k * a.costs_sk + t * a.costs_st] = reach_cost(hand, target);
        if (a.qv_traj) {
            double* o = a.qv_traj + ((size_t)k * a.H + t) * 14;
#pragma unroll
            for (int j = 0; j < 7; j++) { o[j] = q[j]; o[7 + j] = v[j]; }
        }
        if (a.next_obs) {
            double* o = a.next_obs + ((size_t)k * a.H + t) * MJB_OBS_DIM;
#pragma unroll
            for (int j = 0; j < 7; j++) { o[j] = q[j]; o[7 + j] = v[j]; }
            o[14] = hand.x; o[15] = hand.y; o[16] = hand.z;
            o[17] = hand.x - target.x; o[18] = hand.y - target.y; o[19] = hand.z - target.z;
        }
    }
    if (a.ncon) a.ncon[k] = nc;
}

}  // namespace mjb

extern "C" int mjb_rollout_reacher(const mjb_model* m, const mjb_rollout_args* a, void* stream) {
    MJB_REQUIRE(m && a, "mjb_rollout_reacher: null handle");
    MJB_REQUIRE(a->K >= 1 && a->H >= 1, "mjb_rollout_reacher: K and H must be positive (K=%d H=%d)", a->K, a->H);
    MJB_REQUIRE(a->particles_per_ctrl >= 1 && a->K % a->particles_per_ctrl == 0,
                "Number of particles must be divisible by number of controllers");
    MJB_REQUIRE(a->particles_per_model >= 1 && a->K % a->particles_per_model == 0,
                "Number of particles must be divisible by number of cpus");  // subproc_vec_env.py:162
    MJB_REQUIRE(a->state && a->mean && a->costs, "mjb_rollout_reacher: state, mean and costs are required");
    MJB_CUDA(cudaSetDevice(m->device));
    const int block = 64;
    const int grid = (a->K + block - 1) / block;
    cudaStream_t s = (cudaStream_t)stream;
    if (m->fits_sawyer)
        mjb::rollout_reacher_kernel<mjb::SawyerTraits><<<grid, block, 0, s>>>(m->d_params, m->n_instances, *a);
    else
        mjb::rollout_reacher_kernel<mjb::DenseTraits><<<grid, block, 0, s>>>(m->d_params, m->n_instances, *a);
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}
