// K1, role-split instantiation for SMALL particle counts per GPU (<= ~16 k: the sharded regime of N >= 4 GPUs).
//
// With one thread per particle (rollout_reacher_kernel) a launch of K <= 16 k particles leaves every SM
// sub-partition with at most one warp, and that lone warp needs ~2.8 cycles per instruction (ncu at K = 8192,
// profiles/r02_a_k1_8192_ncu.txt: 1 issue + 0.86 fixed-latency dependency + 0.62 instruction fetch + 0.3 other)
// for the ~2400 instructions of a substep: 64 sequential substeps = 0.27 ms however few particles there are.
// The substep is a DAG, not a chain: the bias forces (RNEA), the mass matrix (CRBA) and the constraint rows depend
// only on (q, v); the factorisation of M + hB for the Euler solve does not depend on the constraint solve.  Here
// FOUR WARPS work on the same 32 particles (lane l of every warp = particle l of the block), one role each, on
// the four sub-partitions of an SM, exchanging M, f, the constraint force and qacc through shared memory:
//
//            phase A                     phase B                          phase C            phase D (no barrier after it)
//   role 0   RNEA -> f = u - Bv - c      --                               --                 integrate, advance sin/cos
//   role 1   CRBA -> M                   LDL'(M + hB)                     solve -> qacc      integrate, advance sin/cos
//   role 2   limit / contact rows        Newton on the active set -> fc   --                 integrate, advance sin/cos
//   role 3   hand position -> step cost  next step: u = mean + noise,     --                 integrate, advance sin/cos
//                                        action out, actuator torques
//
// Critical path per substep ~ max(RNEA, CRBA, rows) + Newton + one triangular solve pair + a quarter of the sin/cos
// instead of their sum.  The arithmetic is the monolithic kernel's (same device functions, same order); it replaces
// the reference's per-worker fan-out at small per-worker batch (mjmpc/envs/vec_env/subproc_vec_env.py:161-186 around
// gym_env_wrapper.py:125-153 / reacher_env.py:29-39).
#pragma once

namespace mjb {

// LANES particles per block (32 or 64): every role is LANES / 32 warps that run the same instructions at the same
// time and so share their instruction fetches -- two 32-particle blocks on one SM would be eight independent
// instruction streams (tools/icache_probe.cu: distinct streams beyond the 32 KB L1.5 instruction cache cost 5 - 12
// cycles per instruction at a full grid)
// shared scratch slots beyond the monolithic layout (SC_NSLOT); RNEA's link wrenches get a private region so that
// they cannot collide with the contact Jacobian / M2 the rows warp writes concurrently
enum { SX_F = SC_NSLOT, SX_FC = SX_F + 7, SX_A = SX_FC + 7, SX_PRIV = SX_A + 7, SX_NSLOT = SX_PRIV + 36 };

// dynamic shared memory of the kernel below: nvcc / the block emulator of tests/hostcheck (g++) / the thread-serial
// harness kernel_emu.cpp, which only parses this file
#if defined(__CUDACC__)
#define MJB_DYN_SMEM(name) extern __shared__ double name[]
#elif defined(MJB_HOST_EMU)
#define MJB_DYN_SMEM(name) double* name = nullptr
#else
#define MJB_DYN_SMEM(name) double* name = (double*)emu::g_dyn_smem
#endif

#ifdef MJB_SPLIT_TIMING
// development aid (tools/split_timeline.py): cycles every role spends computing in each phase, block 0 only
__device__ unsigned long long g_split_prof[4][8];
#define MJB_TSTAMP(ph) do { if (blockIdx.x == 0 && lane == 0) { const long long _t = clock64(); prof[ph] += _t - tlast; tlast = _t; } } while (0)
#define MJB_TSYNC(ph) do { __syncthreads(); if (blockIdx.x == 0 && lane == 0) { const long long _t = clock64(); prof[ph] += _t - tlast; tlast = _t; } } while (0)
#else
#define MJB_TSTAMP(ph) ((void)0)
#define MJB_TSYNC(ph) __syncthreads()
#endif

template <int LANES> struct SplitScratch {
    double* base;
#if defined(__CUDA_ARCH__)
    __device__ __forceinline__ void fetch(int slot, const double* g) const {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(base + slot * LANES);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(g) : "memory");
    }
    __device__ __forceinline__ static void fetch_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#else
    void fetch(int slot, const double* g) const { base[slot * LANES] = *g; }
    static void fetch_wait() {}
#endif
    __device__ __forceinline__ double ld(int slot) const { return base[slot * LANES]; }
    __device__ __forceinline__ void st(int slot, double v) { base[slot * LANES] = v; }
};

// EXTRA: also write the state trajectory and the constraint counters (tests)
template <class T, class P, bool EXTRA, int LANES>
__global__ void __launch_bounds__(4 * LANES, 64 / LANES) rollout_reacher_split_kernel(const double* __restrict__ params, int n_inst,
                                                                                     mjb_rollout_args a) {
    MJB_DYN_SMEM(smem);                        // SX_NSLOT * LANES doubles
    const int lane = threadIdx.x % LANES, role = threadIdx.x / LANES;
    // a particle beyond K still walks through every barrier (on particle K-1's data) and writes nothing
    const int kk = blockIdx.x * LANES + lane;
    const bool live = kk < a.K;
    const int k = live ? kk : a.K - 1;
    SplitScratch<LANES> sc{smem + lane};
    SplitScratch<LANES> scp{smem + SX_PRIV * LANES + lane};
    const int ctrl = k / a.particles_per_ctrl;
    P prm;
    if constexpr (std::is_same<P, GlobalParams>::value) {
        const int inst = (k / a.particles_per_model) % n_inst;
        prm.p = params + (size_t)inst * CH_NDEV;
    }
    const double* __restrict__ st = a.state + (size_t)ctrl * MJB_STATE_DIM;
    const double* __restrict__ mean = a.mean + (size_t)ctrl * a.H * 7;
    double q[7], v[7], sn[7], cs[7];
#pragma unroll
    for (int j = 0; j < 7; j++) { q[j] = __ldg(st + j); v[j] = __ldg(st + 7 + j); }
#pragma unroll
    for (int j = 0; j < 7; j++) sincos_joint(q[j], sn[j], cs[j]);
    const int fs = (int)prm[CS_FRAME_SKIP];
    const double h = prm[CS_TIMESTEP];
    // role 3 state
    const V3 target = {__ldg(st + 14), __ldg(st + 15), __ldg(st + 16)};
    const V3 hand_local = {prm[CS_HAND], prm[CS_HAND + 1], prm[CS_HAND + 2]};
    const double* __restrict__ np = a.noise ? a.noise + k * a.noise_sk : nullptr;
    double* __restrict__ ap = a.actions ? a.actions + k * a.act_sk : nullptr;
    double* __restrict__ cp = a.costs + k * a.costs_sk;
    int nc = 0;
    // controls of env step t: u = mean[t] + noise[k, t] -> action out (unclipped), actuator torques -> SC_U; the noise
    // row of step t + 1 is then fetched asynchronously (role 3 only)
    auto step_controls = [&](int t) {
        if (np) SplitScratch<LANES>::fetch_wait();
        double* aj = ap;
#pragma unroll
        for (int j = 0; j < 7; j++) {
            double x = __ldg(mean + t * 7 + j);
            if (np) x += sc.ld(SC_NZ + j);
            sc.st(SC_U + j, actuator_torque(prm, j, x));
            if (ap) { if (live) *aj = x; aj += a.act_sj; }
        }
        if (ap) ap += a.act_st;
        if (np && t + 1 < a.H) {
            const double* nj = np;
#pragma unroll
            for (int j = 0; j < 7; j++) { sc.fetch(SC_NZ + j, nj); nj += a.noise_sj; }
            np += a.noise_st;
        }
    };
    if (role == 3) {
        if (np) {
            const double* nj = np;
#pragma unroll
            for (int j = 0; j < 7; j++) { sc.fetch(SC_NZ + j, nj); nj += a.noise_sj; }
            np += a.noise_st;
        }
        step_controls(0);
    }
    __syncthreads();
#ifdef MJB_SPLIT_TIMING
    long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
#endif
    double Hf[7][7], dinv[7];          // role 1: LDL' of M + hB, alive from phase B to phase C
    for (int t = 0; t < a.H; t++) {
        for (int s = 0; s < fs; s++) {
            const bool last = s == fs - 1;
            // ---------------------------------------------------------------- phase A
            bool any = false;
            Rows R;
            R.Dc = 0.0; R.arefc = 0.0; R.below = 0u;       // (only role 2 fills and reads the rows)
            if (role == 0) {
                double f[7];
                chain_mass_bias<T, 1>(prm, scp, sn, cs, v, f);
#pragma unroll
                for (int j = 0; j < 7; j++) sc.st(SX_F + j, sc.ld(SC_U + j) - prm[CH_DAMPING + j] * v[j] - f[j]);
            } else if (role == 1) {
                double dummy[7];
                chain_mass_bias<T, 2>(prm, sc, sn, cs, v, dummy);
            } else if (role == 2) {
                any = make_rows<T>(prm, sc, q, v, sn, cs, R);
                if (EXTRA) nc += any ? 1 : 0;
            } else if (last) {
                // data.site_xpos after mj_step is the one of the LAST forward pass: the state before the last
                // substep's integration (reacher_env.py:31-35)
                const V3 hand = chain_point_world<T>(prm, sn, cs, hand_local);
                if (live) *cp = reach_cost(hand, target);
                cp += a.costs_st;
            }
            MJB_TSTAMP(0);
            MJB_TSYNC(1);
            // ---------------------------------------------------------------- phase B
            if (role == 1 || role == 2) {
                // roles 1 and 2 run through ONE copy of the load / factor code: role 1 leaves it with the factor of
                // M + hB (kept in Hf / dinv until phase C), role 2 goes on with the Newton solve
                double fc[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                if (role == 1 || any) {
                    double f[7];
#pragma unroll
                    for (int j = 0; j < 7; j++) f[j] = role == 2 ? sc.ld(SX_F + j) : 0.0;
                    constraint_force<T>(prm, sc, f, R, fc, Hf, dinv, role == 1);
                }
                if (role == 2) {
#pragma unroll
                    for (int j = 0; j < 7; j++) sc.st(SX_FC + j, fc[j]);
                }
            } else if (role == 3 && last && t + 1 < a.H) {
                step_controls(t + 1);      // role 0 read SC_U in phase A; the next read is after two barriers
            }
            MJB_TSTAMP(2);
            MJB_TSYNC(3);
            // ---------------------------------------------------------------- phase C
            if (role == 1) {
                double x[7];
#pragma unroll
                for (int j = 0; j < 7; j++) x[j] = sc.ld(SX_F + j) + sc.ld(SX_FC + j);
                ldl7_solve(Hf, dinv, x);
#pragma unroll
                for (int j = 0; j < 7; j++) sc.st(SX_A + j, x[j]);
            }
            MJB_TSTAMP(4);
            MJB_TSYNC(5);
            // ---------------------------------------------------------------- phase D
            // every warp integrates its own copy of (q, v) and advances its own sin / cos with the same
            // instructions: identical values in all four roles, no exchange, no fourth barrier
            bool big = false;
#pragma unroll
            for (int j = 0; j < 7; j++) {
                v[j] = fma(h, sc.ld(SX_A + j), v[j]);
                const double dq = h * v[j];
                q[j] = fma(h, v[j], q[j]);
                big |= fabs(dq) > 0.25;
                sincos_advance(dq, sn[j], cs[j]);
            }
            if (big) {                                  // a joint faster than 25 rad/s: exact sin / cos
                double io[21];
#pragma unroll
                for (int j = 0; j < 7; j++) io[j] = q[j];
                sincos_all(io);
#pragma unroll
                for (int j = 0; j < 7; j++) { sn[j] = io[7 + j]; cs[j] = io[14 + j]; }
            }
            MJB_TSTAMP(6);
        }
        if (EXTRA && role == 3 && live && a.qv_traj) {
            double* o = a.qv_traj + ((size_t)k * a.H + t) * 14;
#pragma unroll
            for (int j = 0; j < 7; j++) { o[j] = q[j]; o[7 + j] = v[j]; }
        }
    }
    if (EXTRA && role == 2 && live && a.ncon) a.ncon[k] = nc;
#ifdef MJB_SPLIT_TIMING
    if (blockIdx.x == 0 && lane == 0)
        for (int i = 0; i < 8; i++) g_split_prof[role][i] = (unsigned long long)prof[i];
#endif
}

}  // namespace mjb
