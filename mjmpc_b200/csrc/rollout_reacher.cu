// K1: batched reacher_7dof rollout with the step cost fused in.  One thread per particle, the
// whole horizon in one launch, state in registers.  Replaces the reference's per-particle
// env-copy loop: GymEnvWrapper.rollout (mjmpc/envs/gym_env_wrapper.py:125-153) ->
// Reacher7DOFEnv.step (mjmpc/envs/basic/reacher_env.py:29-39) -> MuJoCo mj_step x frame_skip,
// fanned out by SubprocVecEnv.rollout (mjmpc/envs/vec_env/subproc_vec_env.py:161-186).
#include <stdlib.h>
#include <mutex>
#include <type_traits>
#include "chain_dynamics.cuh"
#include "philox_noise.cuh"
#include "common.h"

namespace mjb {

struct GlobalParams {
    const double* p;
    __device__ __forceinline__ double operator[](int i) const { return __ldg(p + i); }
};

// Single-model launches (the common case) read the model from constant memory: every parameter
// index is a compile-time constant after unrolling, so it becomes a c[bank][offset] operand of the
// FP64 instruction itself -- no load, no register.  Per-worker models (dynamics randomisation) read
// their instance's block from global memory instead.
__constant__ double c_params[CH_NDEV];
struct ConstParams {
    __device__ __forceinline__ double operator[](int i) const { return c_params[i]; }
};

#define MJB_ROLLOUT_BLOCK 64
#ifndef MJB_OCC
#define MJB_OCC 4          // resident blocks per SM the register allocation is bounded for
#endif
// one shared-memory column per scratch slot: slot * BLOCK + thread (no bank conflicts)
struct SmemScratch {
    double* base;
    // asynchronous 8-byte global -> shared copy into a slot (LDGSTS: no register, no stall until waited for)
#if defined(__CUDA_ARCH__)
    __device__ __forceinline__ void fetch(int slot, const double* g) const {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(base + slot * MJB_ROLLOUT_BLOCK);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(g) : "memory");
    }
    __device__ __forceinline__ static void fetch_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#else       // host pass of nvcc and the host emulation of the kernels (tests/hostcheck/kernel_emu.cpp): a plain copy
    void fetch(int slot, const double* g) const { base[slot * MJB_ROLLOUT_BLOCK] = *g; }
    static void fetch_wait() {}
#endif
    __device__ __forceinline__ double ld(int slot) const { return base[slot * MJB_ROLLOUT_BLOCK]; }
    __device__ __forceinline__ void st(int slot, double v) { base[slot * MJB_ROLLOUT_BLOCK] = v; }
};

// 64-thread blocks, 4 per SM: 256 resident particles per SM at up to 255 registers each; the
// small block keeps the second wave of a K=65536 launch evenly spread over the 148 SMs.
// EXTRA: also write the per-step state trajectory / observations / constraint counters (tests, adaptors
// that hand observations back); the production instantiation carries none of that code.
// FUSED: the action noise is generated in the kernel (Philox + covariance factor + AR filter, the code of
// K2) on the integer / FP32 pipes the FP64-bound rollout leaves idle; the noise tensor never touches HBM.
template <class T, class P, bool EXTRA, bool FUSED>
__global__ void __launch_bounds__(MJB_ROLLOUT_BLOCK, MJB_OCC) rollout_reacher_kernel(const double* __restrict__ params, int n_inst,
                                                                mjb_rollout_args a) {
    __shared__ double smem[(FUSED ? SC_NSLOT_FUSED : SC_NSLOT) * MJB_ROLLOUT_BLOCK];
    __shared__ double Lsh[FUSED ? MJB_MAXD : 1][MJB_MAXD];
    if constexpr (FUSED) {
        if (threadIdx.x == 0) noise_chol<7>(a.noise_cov, Lsh);
        __syncthreads();
    }
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.K) return;
    SmemScratch sc{smem + threadIdx.x};
    const int ctrl = k / a.particles_per_ctrl;
    P prm;
    if constexpr (std::is_same<P, GlobalParams>::value) {
        const int inst = (k / a.particles_per_model) % n_inst;
        prm.p = params + (size_t)inst * CH_NDEV;
    }
    const double* __restrict__ st = a.state + (size_t)ctrl * MJB_STATE_DIM;
    // closed_loop (EXTRA instantiations only): `mean` holds the (d_obs + 1, 7) weights of a linear policy
    const bool closed = EXTRA && a.closed_loop;
    const double* __restrict__ mean = a.mean + (size_t)ctrl * (closed ? (MJB_OBS_DIM + 1) * 7 : a.H * 7);
    double q[7], v[7], sn[7], cs[7];
#pragma unroll
    for (int j = 0; j < 7; j++) { q[j] = __ldg(st + j); v[j] = __ldg(st + 7 + j); }
    const V3 target = {__ldg(st + 14), __ldg(st + 15), __ldg(st + 16)};
    const V3 hand_local = {prm[CS_HAND], prm[CS_HAND + 1], prm[CS_HAND + 2]};
    const int fs = (int)prm[CS_FRAME_SKIP];
    int nc = 0;
    // hand position of the current observation: fresh kinematics at the set state (set_env_state ends with
    // sim.forward(), reacher_env.py:88-99), afterwards the stale one the last step's cost used
    V3 hand_prev = {0.0, 0.0, 0.0};
    if (closed) {
#pragma unroll
        for (int j = 0; j < 7; j++) sincos_joint(q[j], sn[j], cs[j]);
        hand_prev = chain_point_world<T>(prm, sn, cs, hand_local);
    }
    const double* __restrict__ np = a.noise ? a.noise + k * a.noise_sk : nullptr;
    double* __restrict__ ap = a.actions ? a.actions + k * a.act_sk : nullptr;
    double* __restrict__ cp = a.costs + k * a.costs_sk;
    // The noise row of env step t+1 is copied global -> shared asynchronously while step t is simulated,
    // so its HBM latency (the only long-latency load of the loop) is never waited for.
    // fused noise: Philox keyed like mjb_generate_noise
    const unsigned long long gk = FUSED ? (unsigned long long)(a.noise_k_offset + k) : 0ull;
    const Philox ph{(unsigned)a.noise_seed, (unsigned)(a.noise_seed >> 32)};
    const unsigned step_ctr = FUSED ? (a.noise_step_ptr ? (unsigned)(*a.noise_step_ptr) : 0u) + (unsigned)a.noise_offset : 0u;
    const unsigned tag_hi = (unsigned)(a.noise_offset >> 32) << 8;
    const bool zero_seq = FUSED && a.noise_zero_last && (long long)gk == a.noise_K_global - 1;
    if constexpr (FUSED) {
#pragma unroll
        for (int j = 0; j < 7; j++) { sc.st(SC_NZ + j, 0.0); sc.st(SC_E2 + j, 0.0); }
    }
#ifndef MJB_NO_PREFETCH
    if (!FUSED && np) {
        const double* nj = np;
#pragma unroll
        for (int j = 0; j < 7; j++) { sc.fetch(SC_NZ + j, nj); nj += a.noise_sj; }
        np += a.noise_st;
    }
#endif
    for (int t = 0; t < a.H; t++) {
        {
            double* aj = ap;
            double x[7], ub[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            if (closed) {
                // u = mean' [obs; 1] (gym_env_wrapper.py:135-136), obs = (qpos, qvel, hand, hand - target)
                double ob[MJB_OBS_DIM];
#pragma unroll
                for (int j = 0; j < 7; j++) { ob[j] = q[j]; ob[7 + j] = v[j]; }
                ob[14] = hand_prev.x; ob[15] = hand_prev.y; ob[16] = hand_prev.z;
                ob[17] = hand_prev.x - target.x; ob[18] = hand_prev.y - target.y; ob[19] = hand_prev.z - target.z;
#pragma unroll
                for (int j = 0; j < 7; j++) {
                    double s = __ldg(mean + MJB_OBS_DIM * 7 + j);
#pragma unroll
                    for (int i = 0; i < MJB_OBS_DIM; i++) s = fma(__ldg(mean + i * 7 + j), ob[i], s);
                    ub[j] = s;
                }
            }
            if constexpr (FUSED) {
                double z[8], e[7], h1[7], h2[7];
                noise_normals<7>(ph, gk, step_ctr, tag_hi, t, z);
#pragma unroll
                for (int j = 0; j < 7; j++) { h1[j] = sc.ld(SC_NZ + j); h2[j] = sc.ld(SC_E2 + j); }
                noise_shape<7>(Lsh, z, t, a.noise_beta0, a.noise_beta1, a.noise_beta2, h1, h2, e);
#pragma unroll
                for (int j = 0; j < 7; j++) {
                    sc.st(SC_E2 + j, h1[j]); sc.st(SC_NZ + j, e[j]);
                    const double m = __ldg(mean + t * 7 + j);
                    x[j] = m + (zero_seq ? -m : e[j]);
                }
            } else {
#ifndef MJB_NO_PREFETCH
            if (np) SmemScratch::fetch_wait();
#pragma unroll
            for (int j = 0; j < 7; j++) {
                x[j] = closed ? ub[j] : __ldg(mean + t * 7 + j);
                if (np) x[j] += sc.ld(SC_NZ + j);
            }
            if (np && t + 1 < a.H) {
                const double* nj = np;
#pragma unroll
                for (int j = 0; j < 7; j++) { sc.fetch(SC_NZ + j, nj); nj += a.noise_sj; }
                np += a.noise_st;
            }
#else
            {
                const double* nj = np;
#pragma unroll
                for (int j = 0; j < 7; j++) {
                    x[j] = closed ? ub[j] : __ldg(mean + t * 7 + j);
                    if (np) { x[j] += __ldg(nj); nj += a.noise_sj; }
                }
                if (np) np += a.noise_st;
            }
#endif
            }
#pragma unroll
            for (int j = 0; j < 7; j++) {
                sc.st(SC_U + j, actuator_torque(prm, j, x[j]));   // held over the frame_skip substeps
                if (ap) { *aj = x[j]; aj += a.act_sj; }
            }
            if (ap) ap += a.act_st;
        }
        V3 hand = {0.0, 0.0, 0.0};
        for (int s = 0; s < fs; s++) {
#pragma unroll
            for (int j = 0; j < 7; j++) sincos_joint(q[j], sn[j], cs[j]);
            // data.site_xpos read after mj_step is the one computed by the LAST forward pass, i.e. at
            // the state before the last substep's integration (reacher_env.py:31-35).
            if (s == fs - 1) hand = chain_point_world<T>(prm, sn, cs, hand_local);
            const bool any = chain_substep<T>(prm, sc, q, v, sn, cs);
            if (EXTRA) nc += any ? 1 : 0;
        }
        *cp = reach_cost(hand, target);
        cp += a.costs_st;
        if (EXTRA) hand_prev = hand;
        if (EXTRA && a.qv_traj) {
            double* o = a.qv_traj + ((size_t)k * a.H + t) * 14;
#pragma unroll
            for (int j = 0; j < 7; j++) { o[j] = q[j]; o[7 + j] = v[j]; }
        }
        if (EXTRA && a.next_obs) {
            double* o = a.next_obs + ((size_t)k * a.H + t) * MJB_OBS_DIM;
#pragma unroll
            for (int j = 0; j < 7; j++) { o[j] = q[j]; o[7 + j] = v[j]; }
            o[14] = hand.x; o[15] = hand.y; o[16] = hand.z;
            o[17] = hand.x - target.x; o[18] = hand.y - target.y; o[19] = hand.z - target.z;
        }
    }
    if (EXTRA && a.ncon) a.ncon[k] = nc;
}

}  // namespace mjb

#include "rollout_reacher_split.cuh"

#ifndef MJB_HOST_EMU
namespace mjb {
// owner of the per-device constant bank (see common.h)
static std::mutex g_bank_mutex;
static const mjb_model* g_bank_owner[64] = {nullptr};
static unsigned long long g_bank_serial[64] = {0};

static int bank_upload(const mjb_model* m, cudaStream_t s) {
    MJB_CUDA(cudaSetDevice(m->device));
    MJB_CUDA(cudaMemcpyToSymbolAsync(c_params, m->h_params, sizeof(double) * CH_NDEV, 0, cudaMemcpyHostToDevice, s));
    g_bank_serial[m->device & 63] = m->serial;
    return MJB_OK;
}
// true: this launch may read the model from the constant bank
static int bank_claim(const mjb_model* m, int K, cudaStream_t s, bool* use_const) {
    *use_const = false;
    if (m->n_instances != 1) return MJB_OK;
    std::lock_guard<std::mutex> lock(g_bank_mutex);
    const int dev = m->device & 63;
    if (g_bank_owner[dev] == nullptr && K >= 64) {
        g_bank_owner[dev] = m;
        g_bank_serial[dev] = 0;
    }
    if (g_bank_owner[dev] != m) return MJB_OK;
    if (g_bank_serial[dev] != m->serial) {
        const int rc = bank_upload(m, s);       // first launch of the owner (stream-ordered before the kernel)
        if (rc != MJB_OK) return rc;
    }
    *use_const = true;
    return MJB_OK;
}
int const_bank_on_update(mjb_model* m) {
    std::lock_guard<std::mutex> lock(g_bank_mutex);
    const int dev = m->device & 63;
    if (g_bank_owner[dev] != m) return MJB_OK;
    // kernels (or graph replays) of the owner may be in flight on any stream: drain, then replace the constants
    MJB_CUDA(cudaSetDevice(m->device));
    MJB_CUDA(cudaDeviceSynchronize());
    const int rc = bank_upload(m, nullptr);
    if (rc != MJB_OK) return rc;
    MJB_CUDA(cudaStreamSynchronize(nullptr));
    return MJB_OK;
}
void const_bank_release(const mjb_model* m) {
    std::lock_guard<std::mutex> lock(g_bank_mutex);
    const int dev = m->device & 63;
    if (g_bank_owner[dev] == m) { g_bank_owner[dev] = nullptr; g_bank_serial[dev] = 0; }
}
// particles per launch up to which the role-split kernel is used (4 warps per 32 particles; measured cross-over,
// DESIGN 4.2); MJB_SPLIT_MAX_K overrides it (0 = never)
static int g_split_max_k = -1;
static int split_max_k() {
    if (g_split_max_k < 0) {
        const char* e = getenv("MJB_SPLIT_MAX_K");
        g_split_max_k = e ? atoi(e) : 8192;
    }
    return g_split_max_k;
}
}  // namespace mjb

#ifdef MJB_SPLIT_TIMING
extern "C" int mjb_split_profile(unsigned long long* out32) {
    MJB_CUDA(cudaMemcpyFromSymbol(out32, mjb::g_split_prof, sizeof(unsigned long long) * 32));
    return MJB_OK;
}
#endif

extern "C" int mjb_rollout_split_max_k(int new_value) {
    const int old = mjb::split_max_k();
    if (new_value >= 0) mjb::g_split_max_k = new_value;
    return old;
}

extern "C" int mjb_rollout_reacher(const mjb_model* m, const mjb_rollout_args* a, void* stream) {
    MJB_REQUIRE(m && a, "mjb_rollout_reacher: null handle");
    MJB_REQUIRE(a->K >= 1 && a->H >= 1, "mjb_rollout_reacher: K and H must be positive (K=%d H=%d)", a->K, a->H);
    MJB_REQUIRE(a->particles_per_ctrl >= 1 && a->K % a->particles_per_ctrl == 0,
                "Number of particles must be divisible by number of controllers");
    MJB_REQUIRE(a->particles_per_model >= 1 && a->K % a->particles_per_model == 0,
                "Number of particles must be divisible by number of cpus");  // subproc_vec_env.py:162
    MJB_REQUIRE(a->state && a->mean && a->costs, "mjb_rollout_reacher: state, mean and costs are required");
    MJB_CUDA(cudaSetDevice(m->device));
    const int block = MJB_ROLLOUT_BLOCK;
    const int grid = (a->K + block - 1) / block;
    cudaStream_t s = (cudaStream_t)stream;
    const bool extra = a->qv_traj || a->next_obs || a->ncon || a->closed_loop;
    const bool fused = a->noise_cov != nullptr;
    MJB_REQUIRE(!(fused && a->closed_loop), "mjb_rollout_reacher: closed-loop rollouts take an explicit noise tensor");
    MJB_REQUIRE(!(fused && a->noise), "mjb_rollout_reacher: pass either a noise tensor or in-kernel noise parameters, not both");
    bool use_const = false;
    {
        const int rc = mjb::bank_claim(m, a->K, s, &use_const);
        if (rc != MJB_OK) return rc;
    }
    // small launches: four warps per 32 particles (rollout_reacher_split.cuh); it has no observation / closed-loop /
    // fused-noise instantiations
    if (a->K <= mjb::split_max_k() && !fused && !a->closed_loop && !a->next_obs && m->uniform_frame_skip) {
        // 32 particles per block while that leaves every SM at most one block, else 64 (two warps per role in lock step)
        int sm_count = 148;
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, m->device);
        const bool wide = (a->K + 31) / 32 > sm_count;
        const int lanes = wide ? 64 : 32;
        const int sgrid = (a->K + lanes - 1) / lanes;
        const size_t sbytes = sizeof(double) * mjb::SX_NSLOT * lanes;
        const bool sx = a->qv_traj || a->ncon;
#define MJB_SLAUNCH4(T, P, E, L)                                                                                          \
    do {                                                                                                                  \
        MJB_CUDA(cudaFuncSetAttribute(mjb::rollout_reacher_split_kernel<mjb::T, mjb::P, E, L>,                              \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * mjb::SX_NSLOT * L))); \
        mjb::rollout_reacher_split_kernel<mjb::T, mjb::P, E, L><<<sgrid, 4 * L, sbytes, s>>>(m->d_params, m->n_instances, *a); \
    } while (0)
#define MJB_SLAUNCH3(T, P, E) do { if (wide) MJB_SLAUNCH4(T, P, E, 64); else MJB_SLAUNCH4(T, P, E, 32); } while (0)
#define MJB_SLAUNCH(T, P) do { if (sx) MJB_SLAUNCH3(T, P, true); else MJB_SLAUNCH3(T, P, false); } while (0)
        if (m->fits_sawyer) {
            if (use_const) MJB_SLAUNCH(SawyerTraits, ConstParams);
            else MJB_SLAUNCH(SawyerTraits, GlobalParams);
        } else {
            MJB_SLAUNCH(DenseTraits, GlobalParams);
        }
#undef MJB_SLAUNCH4
#undef MJB_SLAUNCH3
#undef MJB_SLAUNCH
        MJB_CUDA(cudaGetLastError());
        return MJB_OK;
    }
#define MJB_LAUNCH4(T, P, E, F) mjb::rollout_reacher_kernel<mjb::T, mjb::P, E, F><<<grid, block, 0, s>>>(m->d_params, m->n_instances, *a)
#define MJB_LAUNCH(T, P)                                                    \
    do {                                                                    \
        if (extra) { if (fused) MJB_LAUNCH4(T, P, true, true); else MJB_LAUNCH4(T, P, true, false); }   \
        else { if (fused) MJB_LAUNCH4(T, P, false, true); else MJB_LAUNCH4(T, P, false, false); }       \
    } while (0)
    if (m->fits_sawyer) {
        if (use_const) MJB_LAUNCH(SawyerTraits, ConstParams);
        else MJB_LAUNCH(SawyerTraits, GlobalParams);
    } else {
        // arbitrary offsets / COMs / inertias on the same axis pattern: dense variant, global parameters
        MJB_LAUNCH(DenseTraits, GlobalParams);
    }
#undef MJB_LAUNCH4
#undef MJB_LAUNCH
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}
#endif  // MJB_HOST_EMU
