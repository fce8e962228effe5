// K3-K8: everything the controllers do with a finished rollout -- discounted cost-to-go, the
// exponential-utility (softmax) reductions of MPPI / DMD-MPC / PFMPC, CEM elite selection and
// moments, RandomShooting argmin, PFMPC systematic resampling and the hot-start shifts.
// Reference sites are cited per entry point in include/mjmpc_b200.h.  All reductions are
// two-stage and ordered (per-block partials, then a fixed-order sum), so results are
// deterministic and independent of scheduling; shards combine in rank order.
#include <math.h>
#include "common.h"

#define MJB_MAXD 8
#define MJB_MAXH 128
#define MJB_CHUNK 2048        // particles per block in the (chunk, t) reductions
#define MJB_RB 256            // threads per reduction block

namespace mjb {

struct GammaSeq { double g[MJB_MAXH]; int raw; };

static int load_gamma(GammaSeq& G, const double* host, int H) {
    if (H > MJB_MAXH) return set_error(MJB_EINVAL, "horizon %d exceeds the supported maximum %d", H, MJB_MAXH);
    G.raw = 0;
    for (int t = 0; t < H; t++) { G.g[t] = host[t]; if (host[t] == 0.0) G.raw = 1; }
    return MJB_OK;
}

// Order-preserving 64-bit key of a double.  -0.0 is folded onto +0.0 first: numpy compares them equal, so a tie
// between the two must go to the lower index like any other tie (np.argmin / the stable sort order), not to the
// one whose sign bit happens to be set (costs = -rewards turns a reward of 0.0 into a cost of -0.0).
__device__ __forceinline__ unsigned long long enc_key(double x) {
    x = (x == 0.0) ? 0.0 : x;
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dec_key(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)b);
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------ cost_to_go
// control_utils.py:37-46 with numpy's operation order: multiply, sequential reverse add, divide.
__global__ void cost_to_go_kernel(const double* __restrict__ c, long long sk, long long st, GammaSeq G, int K, int H,
                                  double* __restrict__ out, long long osk, long long ost) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    double S = 0.0;
    for (int t = H - 1; t >= 0; t--) {
        const double x = c[k * sk + t * st];
        if (G.raw) { out[k * osk + t * ost] = x; continue; }
        S = __dadd_rn(S, __dmul_rn(G.g[t], x));
        out[k * osk + t * ost] = __ddiv_rn(S, G.g[t]);
    }
}

// ------------------------------------------------------------------------------ softmax phase 1
// scratch layout (doubles): [0] "blocks done" counter of the fused update (zero once, resets itself) | [MAXH, MAXH + H*d) u_n |
// block partials H*nchunks*NACC | block minima T*nb (nb = blocks of traj_cost_kernel)
__host__ __device__ inline long long sm_off_partials(int H, int d) { return MJB_MAXH + (long long)H * d; }
__host__ __device__ inline int sm_nb(int K) { return (K + MJB_RB - 1) / MJB_RB; }
__host__ __device__ inline int sm_nchunks(int K) { return (K + MJB_CHUNK - 1) / MJB_CHUNK; }
__host__ __device__ inline long long sm_off_bmin(int K, int H, int d, int NACC) {
    return sm_off_partials(H, d) + (long long)H * sm_nchunks(K) * NACC;
}
// only with the control cost (MPPI alpha == 0): u_n = mean @ inv(cov)
__global__ void softmax_prep_kernel(mjb_softmax_args a, int T) {
    if (!a.control_cost) return;
    // u_n = mean @ inv(cov)   (mppi.py:106); inverse by Gauss-Jordan with partial pivoting on a d x 2d tableau
    __shared__ double A[MJB_MAXD][2 * MJB_MAXD];
    const int d = a.d;
    if (threadIdx.x == 0) {
        for (int i = 0; i < d; i++)
            for (int j = 0; j < d; j++) { A[i][j] = a.cov[i * d + j]; A[i][d + j] = i == j ? 1.0 : 0.0; }
        for (int c = 0; c < d; c++) {
            int p = c;
            for (int r = c + 1; r < d; r++) if (fabs(A[r][c]) > fabs(A[p][c])) p = r;
            if (p != c) for (int j = 0; j < 2 * d; j++) { const double t = A[c][j]; A[c][j] = A[p][j]; A[p][j] = t; }
            const double inv = 1.0 / A[c][c];
            for (int j = 0; j < 2 * d; j++) A[c][j] *= inv;
            for (int r = 0; r < d; r++) if (r != c) {
                const double f = A[r][c];
                for (int j = 0; j < 2 * d; j++) A[r][j] -= f * A[c][j];
            }
        }
    }
    __syncthreads();
    double* un = a.scratch + MJB_MAXH;
    for (int i = threadIdx.x; i < a.H * d; i += blockDim.x) {
        const int t = i / d, j = i % d;
        double s = 0.0;
        for (int l = 0; l < d; l++) s += a.mean[t * d + l] * A[l][d + j];
        un[i] = s;
    }
}

// per-particle trajectory cost (mppi.py:84-97 / gaussian_dmd.py:94-104) and its minimum over particles.
// TD = true: MPPIQ's TD(lambda) return (mppiq.py:92-126) instead of the discounted cost-to-go; W is then
// the (H-1)-entry weight sequence cumprod([1, gamma*td_lam, ...]) and G is unused.
// T1: one weight row (not time-based): only the t = 0 entry of the cost-to-go is used, so the per-step divisions
// and stores of the other rows are skipped (the accumulation S is the same sequence of operations).
// Minimum: warp shuffle, then one value per block and row in bmin[row * nb + block] -- no atomics, nothing to reset.
template <int D, bool TD, bool T1>
__global__ void __launch_bounds__(MJB_RB) traj_cost_kernel(mjb_softmax_args a, GammaSeq G, GammaSeq W, int T, double* __restrict__ bmin) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const double* __restrict__ un = a.scratch + MJB_MAXH;
    const bool live = k < a.K;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double S = 0.0, Sc = 0.0, tot0 = INFINITY, qnext = 0.0;
    double keep[MJB_MAXH / 32];          // lane l keeps the warp minimum of rows l, l + 32, ...
#pragma unroll
    for (int i = 0; i < MJB_MAXH / 32; i++) keep[i] = INFINITY;
    // costs are read PF steps at a time ahead of the (sequential, reference-ordered) accumulation: the kernel is one
    // load latency per batch long (32 steps in flight when only the t = 0 row is needed: one batch at H = 32)
    constexpr int PF = (T1 && !TD) ? 32 : 8;
    for (int tb = a.H - 1; tb >= 0; tb -= PF) {
        double cbuf[PF], qbuf[TD ? PF : 1];
#pragma unroll
        for (int u = 0; u < PF; u++) {
            const int t = tb - u;
            cbuf[u] = (live && t >= 0) ? a.costs[k * a.costs_sk + t * a.costs_st] : 0.0;
            if (TD) qbuf[u] = (live && t >= 0 && a.qvals) ? a.qvals[k * a.q_sk + t * a.q_st] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < PF; u++) {
            const int t = tb - u;
            if (t < 0) break;
            const bool need = !T1 || t == 0;       // this row's trajectory cost is used
            double tot = INFINITY;
            if (live) {
                const double c = cbuf[u];
                double cc = 0.0;
                if (a.control_cost) {
#pragma unroll
                    for (int j = 0; j < D; j++) {
                        const double m = a.mean[t * D + j];
                        const double dl = a.actions[k * a.act_sk + t * a.act_st + j * a.act_sj] - m;
                        cc += 0.5 * un[t * D + j] * (m + 2.0 * dl);
                    }
                }
                if constexpr (TD) {
                    // mppiq.py:96-97 total per-step cost, then calculate_returns (:104-126), reverse order
                    const double ct = a.control_cost ? __dadd_rn(c, __dmul_rn(a.lam, cc)) : c;
                    const double q = a.qvals ? qbuf[u] : (t == a.H - 1 ? ct : 0.0);
                    if (t == a.H - 1) tot = q;
                    else {
                        const double td = __dsub_rn(__dadd_rn(ct, __dmul_rn(a.td_gamma, qnext)), q);
                        double ctg = 0.0;
                        if (W.raw) ctg = td;
                        else { S = __dadd_rn(S, __dmul_rn(W.g[t], td)); if (need) ctg = __ddiv_rn(S, W.g[t]); }
                        tot = __dadd_rn(q, __dmul_rn(a.td_lam, ctg));
                    }
                    qnext = q;
                } else {
                    double ctg = 0.0;
                    if (G.raw) ctg = c;
                    else { S = __dadd_rn(S, __dmul_rn(G.g[t], c)); if (need) ctg = __ddiv_rn(S, G.g[t]); }
                    double ccg = 0.0;
                    if (a.control_cost) {
                        if (G.raw) ccg = cc;
                        else { Sc = __dadd_rn(Sc, __dmul_rn(G.g[t], cc)); if (need) ccg = __ddiv_rn(Sc, G.g[t]); }
                    }
                    tot = ctg + a.lam * ccg;
                }
                if (!T1) a.total[(long long)t * a.K + k] = tot;
                tot0 = tot;
            }
            if (!T1) {
                double mn = tot;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
#pragma unroll
                for (int i = 0; i < MJB_MAXH / 32; i++) if ((t >> 5) == i && lane == (t & 31)) keep[i] = mn;
            }
        }
    }
    __shared__ double wmin[MJB_RB / 32][MJB_MAXH];
    if (T1) {
        if (live) a.total[k] = tot0;
        double mn = tot0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        if (lane == 0) wmin[wid][0] = mn;
    } else {
#pragma unroll
        for (int i = 0; i < MJB_MAXH / 32; i++) if (lane + 32 * i < T) wmin[wid][lane + 32 * i] = keep[i];
    }
    __syncthreads();
    for (int tr = threadIdx.x; tr < T; tr += blockDim.x) {
        double m = wmin[0][tr];
#pragma unroll
        for (int w2 = 1; w2 < MJB_RB / 32; w2++) m = fmin(m, wmin[w2][tr]);
        bmin[(long long)tr * gridDim.x + blockIdx.x] = m;
    }
}

// minimum of one row of block minima, by all threads of a block (every thread returns it)
__device__ double row_min(const double* __restrict__ bmin, int nb, int tr) {
    __shared__ double rm[MJB_RB / 32];
    double m = INFINITY;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) m = fmin(m, bmin[(long long)tr * nb + i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
    __syncthreads();                       // rm may still be read by a previous call
    if ((threadIdx.x & 31) == 0) rm[threadIdx.x >> 5] = m;
    __syncthreads();
    m = rm[0];
    for (int w2 = 1; w2 < (int)(blockDim.x >> 5); w2++) m = fmin(m, rm[w2]);
    return m;
}

// Generic weighted reduction over particles for one (chunk, t) tile.
//   WMODE 0: w = exp(ninv*total - ninv*m), m = row minimum from the block minima   1: w = flag (0/1)   2: w = 1
//   CMODE 0: none  1: diag  2: full lower triangle of  w (delta-mu)(delta-mu)',  delta = a - mean[t]
// block partial -> out[(t*nchunks + chunk)*NACC + c],  NACC = 1 + D + ncov.  `stage`: >= 8 * NACC doubles of shared memory.
template <int D, int WMODE, int CMODE>
__device__ __forceinline__ void reduce_tile(
    int K, int H, const double* __restrict__ total, int T, const double* __restrict__ bmin, int nb, double ninv,
    const unsigned char* __restrict__ flags, const double* __restrict__ actions, long long sk, long long st, long long sj,
    const double* __restrict__ mean, const double* __restrict__ mu, double* __restrict__ out, double* stage) {
    constexpr int NC = CMODE == 0 ? 0 : (CMODE == 1 ? D : D * (D + 1) / 2);
    constexpr int NACC = 1 + D + NC;
    const int t = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
    double acc[NACC];
#pragma unroll
    for (int c = 0; c < NACC; c++) acc[c] = 0.0;
    double xmax = 0.0;
    const double* trow = nullptr;
    if (WMODE == 0) { const int tr = T > 1 ? t : 0; xmax = ninv * row_min(bmin, nb, tr); trow = total + (long long)tr * K; }
    double mrow[D], murow[D];
#pragma unroll
    for (int j = 0; j < D; j++) { mrow[j] = (CMODE != 0) ? mean[t * D + j] : 0.0; murow[j] = (CMODE != 0 && mu) ? mu[j] : 0.0; }
    const long long k0 = (long long)chunk * MJB_CHUNK;
    // 4 particles per trip with every load issued before the first use: 32 independent loads in flight
    // per thread (the one-particle-per-trip form sat on the load latency at a quarter of HBM speed)
    constexpr int U = 4;
    for (int i0 = threadIdx.x; i0 < MJB_CHUNK; i0 += MJB_RB * U) {
        double av[U][D], tw[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const long long k = k0 + i0 + u * MJB_RB;
            ok[u] = k < K;
            const long long kk = ok[u] ? k : k0;
            if (WMODE == 0) tw[u] = trow[kk];
            else if (WMODE == 1) tw[u] = flags[kk] ? 1.0 : 0.0;
            else tw[u] = 1.0;
#pragma unroll
            for (int j = 0; j < D; j++) av[u][j] = actions[kk * sk + t * st + j * sj];
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            double w = WMODE == 0 ? exp(ninv * tw[u] - xmax) : tw[u];
            w = ok[u] ? w : 0.0;
            acc[0] += w;
            double dl[D];
#pragma unroll
            for (int j = 0; j < D; j++) {
                acc[1 + j] += w * av[u][j];
                dl[j] = av[u][j] - mrow[j] - murow[j];
            }
            if (CMODE == 1) {
#pragma unroll
                for (int j = 0; j < D; j++) acc[1 + D + j] += w * (dl[j] * dl[j]);
            } else if (CMODE == 2) {
                int c = 1 + D;
#pragma unroll
                for (int i2 = 0; i2 < D; i2++)
#pragma unroll
                    for (int j = 0; j <= i2; j++) acc[c++] += w * (dl[i2] * dl[j]);
            }
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < NACC; c++) {
        const double v = warp_sum(acc[c]);
        if (lane == 0) stage[wid * NACC + c] = v;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < NACC; c += MJB_RB) {
        double s = 0.0;
#pragma unroll
        for (int w2 = 0; w2 < MJB_RB / 32; w2++) s += stage[w2 * NACC + c];
        out[((long long)t * nchunks + chunk) * NACC + c] = s;
    }
}

template <int D, int WMODE, int CMODE>
__global__ void __launch_bounds__(MJB_RB) weighted_reduce_kernel(
    int K, int H, const double* __restrict__ total, int T, const double* __restrict__ bmin, int nb, double ninv,
    const unsigned char* __restrict__ flags, const double* __restrict__ actions, long long sk, long long st, long long sj,
    const double* __restrict__ mean, const double* __restrict__ mu, double* __restrict__ out) {
    __shared__ double stage[(MJB_RB / 32) * (1 + MJB_MAXD + MJB_MAXD * (MJB_MAXD + 1) / 2)];
    reduce_tile<D, WMODE, CMODE>(K, H, total, T, bmin, nb, ninv, flags, actions, sk, st, sj, mean, mu, out, stage);
}

// loads of values another block / another GPU wrote during this kernel: from L2, never from this SM's L1
__device__ __forceinline__ double ld_cg(const double* p) {
#if defined(__CUDA_ARCH__)
    return __ldcg(p);
#else
    return *p;
#endif
}

// fixed-order sum over chunks: part[t*NACC + c] = sum_chunk blockpart[(t*nchunks+chunk)*NACC + c]
// (eight loads issued before the first add: the one-load-per-add form paid one L2 round trip per chunk)
__device__ __forceinline__ double chunk_sum_one(const double* __restrict__ p, int nchunks, int NACC) {
    double s = 0.0;
    int ch = 0;
    for (; ch + 32 <= nchunks; ch += 32) {          // K = 65536: all 32 chunks of a row in flight at once
        double v[32];
#pragma unroll
        for (int u = 0; u < 32; u++) v[u] = ld_cg(p + (long long)(ch + u) * NACC);
#pragma unroll
        for (int u = 0; u < 32; u++) s += v[u];
    }
    for (; ch + 8 <= nchunks; ch += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = ld_cg(p + (long long)(ch + u) * NACC);
#pragma unroll
        for (int u = 0; u < 8; u++) s += v[u];
    }
    for (; ch < nchunks; ch++) s += ld_cg(p + (long long)ch * NACC);
    return s;
}
__global__ void chunk_sum_kernel(const double* __restrict__ bp, int nchunks, int NACC, double* __restrict__ part) {
    const int t = blockIdx.x;
    for (int c = threadIdx.x; c < NACC; c += blockDim.x)
        part[t * NACC + c] = chunk_sum_one(bp + (long long)t * nchunks * NACC + c, nchunks, NACC);
}
// phase-1 epilogue of the step-by-step path, one block per row t: the row's chunk sums and (t < T) its minimum
__global__ void __launch_bounds__(MJB_RB) softmax_finalize_kernel(const double* __restrict__ bp, int nchunks, int NACC, int T,
                                                                  const double* __restrict__ bmin, int nb, double* __restrict__ part) {
    const int t = blockIdx.x;
    for (int c = threadIdx.x; c < NACC; c += blockDim.x)
        part[T + t * NACC + c] = chunk_sum_one(bp + (long long)t * nchunks * NACC + c, nchunks, NACC);
    if (t < T) {
        const double m = row_min(bmin, nb, t);
        if (threadIdx.x == 0) part[t] = m;
    }
}

// ------------------------------------------------------------------------------ softmax phase 2
struct CombineSmem {
    double mstar[MJB_MAXH];
    double comb[MJB_MAXH][1 + MJB_MAXD + MJB_MAXD * (MJB_MAXD + 1) / 2];
};
__device__ void softmax_combine_body(const mjb_combine_args& a, CombineSmem& sm) {
    const int H = a.H, d = a.d, T = a.time_based ? H : 1;
    const int nc = a.cov_mode == MJB_COV_NONE ? 0 : (a.cov_mode == MJB_COV_DIAG ? d : d * (d + 1) / 2);
    const int NACC = 1 + d + nc, P = T + H * NACC;
    const double ninv = -1.0 / a.lam;
    for (int tr = threadIdx.x; tr < T; tr += blockDim.x) {
        double m = INFINITY;
        for (int r = 0; r < a.n_shards; r++) m = fmin(m, ld_cg(a.partials + (long long)r * P + tr));
        sm.mstar[tr] = m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < H * NACC; i += blockDim.x) {
        const int t = i / NACC, tr = T > 1 ? t : 0;
        double s = 0.0;
        for (int r = 0; r < a.n_shards; r++) {
            const double* p = a.partials + (long long)r * P;
            const double scale = exp(ninv * ld_cg(p + tr) - ninv * sm.mstar[tr]);   // <= 1: shard minimum vs global minimum
            s += ld_cg(p + T + i) * scale;
        }
        sm.comb[t][i % NACC] = s;
    }
    __syncthreads();
    // a.mean == NULL: statistics only (the _calc_val pass leaves the distribution untouched)
    for (int i = threadIdx.x; a.mean && i < H * d; i += blockDim.x) {
        const int t = i / d, j = i % d;
        a.mean[i] = (1.0 - a.step_size) * a.mean[i] + a.step_size * (sm.comb[t][1 + j] / sm.comb[t][0]);
    }
    if (a.mean && a.cov_mode != MJB_COV_NONE) {
        for (int i = threadIdx.x; i < d * d; i += blockDim.x) {
            const int r = i / d, c = i % d;
            double upd = 0.0;
            if (a.cov_mode == MJB_COV_DIAG) {
                if (r == c) { for (int t = 0; t < H; t++) upd += sm.comb[t][1 + d + r] / sm.comb[t][0]; upd /= H; }
            } else {
                const int hi = r > c ? r : c, lo = r > c ? c : r;
                const int idx = 1 + d + hi * (hi + 1) / 2 + lo;
                for (int t = 0; t < H; t++) upd += sm.comb[t][idx] / sm.comb[t][0];
                upd /= H;
            }
            a.cov[i] = (1.0 - a.step_size) * a.cov[i] + a.step_size * upd;
        }
    }
    if (a.stats) {
        if (threadIdx.x == 0) {
            a.stats[0] = -a.lam * (ninv * sm.mstar[0] + log(sm.comb[0][0] / (double)a.K_global));
            a.stats[1] = sm.mstar[0];
        }
        for (int tr = threadIdx.x; tr < T; tr += blockDim.x) a.stats[2 + tr] = sm.comb[tr][0];
        // minima for the other rows of a time-based weighting live after the normalisers
        for (int tr = threadIdx.x; tr < T; tr += blockDim.x) a.stats[2 + T + tr] = sm.mstar[tr];
    }
}

__global__ void softmax_combine_kernel(mjb_combine_args a) {
    __shared__ CombineSmem sm;
    softmax_combine_body(a, sm);
}

// Exchange of the per-rank partial vectors over NVLink peer memory, by one block per GPU.  Every rank owns a
// symmetric buffer [2][world][P] doubles + [2][world] sequence flags, mapped into all peers.  The block
//   1. stores its partial vector straight into slot [parity][rank] of EVERY peer's buffer (P2P stores),
//   2. publishes its sequence number in every peer's flag word (release at system scope after a system fence),
//   3. spins (acquire at system scope) until all `world` flags of its own buffer carry this sequence number,
// and returns the `world` partial vectors in rank order; the caller reads them with L1-bypassing loads (ld_cg).
// Two parity halves: a rank can be at most one exchange ahead of the slowest peer.
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
#if defined(__CUDA_ARCH__)
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#else
    *(volatile unsigned long long*)p = v;
#endif
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
#if defined(__CUDA_ARCH__)
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
#else
    return *(const volatile unsigned long long*)p;
#endif
}
__device__ const double* exchange_partials(const double* __restrict__ local, int P, void* const* __restrict__ peers, int rank,
                                           unsigned long long seq, int world) {
    const int parity = (int)(seq & 1ull);
    for (int r = 0; r < world; r++) {
        double* dst = (double*)peers[r] + ((long long)parity * world + rank) * P;
        for (int i = threadIdx.x; i < P; i += blockDim.x) dst[i] = ld_cg(local + i);
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world)
        st_release_sys((unsigned long long*)((double*)peers[threadIdx.x] + 2ll * world * P) + parity * world + rank, seq);
    double* mine = (double*)peers[rank];
    if ((int)threadIdx.x < world) {
        const unsigned long long* flag = (const unsigned long long*)(mine + 2ll * world * P) + parity * world + threadIdx.x;
        while (ld_acquire_sys(flag) < seq) { __nanosleep(64); }
    }
    __threadfence_system();
    __syncthreads();
    return mine + (long long)parity * world * P;
}

// Fused exchange + combine (one block per GPU): the 2 KB exchange costs one NVLink round trip inside the reduction
// epilogue instead of a separate NCCL launch.
__global__ void __launch_bounds__(256) softmax_exchange_combine_kernel(mjb_combine_args a, const double* __restrict__ local,
                                                                        int P, void* const* __restrict__ peers, int rank,
                                                                        unsigned long long seq) {
    __shared__ CombineSmem sm;
    a.partials = exchange_partials(local, P, peers, rank, seq, a.n_shards);
    softmax_combine_body(a, sm);
}

// ------------------------------------------------------------------------------ the whole update tail in one kernel
// weighted reduction (WMODE 0) whose LAST block to finish also runs everything that used to follow it as separate
// launches: chunk sums + row minima -> this rank's partial vector, [peer exchange,] combine, next action, shift of
// the mean sequence, cov += beta I.  Same arithmetic in the same order as the separate kernels (bit-identical).
struct TailArgs {
    mjb_combine_args c;
    double* partials;                    // this rank's partial vector (P doubles), written here
    const double* bp;                    // block partials of the reduction
    const double* bmin;
    int nb, nch, NACC, T, P;
    unsigned int* counter;               // blocks done; zero before the first launch, reset by the last block
    void* const* peers; int rank; unsigned long long seq;      // c.n_shards > 1
    double* action_out;                  // (d,) <- mean[0] after the update, before the shift; or NULL
    int shift, base_action;
    double cov_shift_beta;
};

__device__ void shift_mean_body(double* mean, int H, int d, int base, const double* rnd, double* buf) {
    // single block: read everything, sync, write (rows overlap)
    for (int i = threadIdx.x; i < H * d; i += blockDim.x) buf[i] = mean[i];
    __syncthreads();
    for (int i = threadIdx.x; i < H * d; i += blockDim.x) {
        const int t = i / d, j = i % d;
        double v;
        if (t < H - 1) v = buf[(t + 1) * d + j];
        else if (base == MJB_BASE_NULL) v = 0.0;
        else if (base == MJB_BASE_REPEAT) v = H >= 2 ? buf[(H - 1) * d + j] : buf[j];   // mean[-2] after the shift
        else v = rnd[j];
        mean[i] = v;
    }
}

template <int D, int CMODE>
__global__ void __launch_bounds__(MJB_RB) softmax_reduce_tail_kernel(
    int K, int H, const double* __restrict__ total, int T, const double* __restrict__ bmin, int nb, double ninv,
    const double* __restrict__ actions, long long sk, long long st, long long sj, const double* __restrict__ mean,
    double* __restrict__ out, TailArgs ta) {
    __shared__ CombineSmem sm;           // the staging area of the reduction aliases it (dead before the tail starts)
    __shared__ bool is_last;
    reduce_tile<D, 0, CMODE>(K, H, total, T, bmin, nb, ninv, nullptr, actions, sk, st, sj, mean, nullptr, out, &sm.comb[0][0]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(ta.counter, 1u) == gridDim.x * gridDim.y - 1u;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // this rank's partial vector (the work of softmax_finalize_kernel)
    for (int i = threadIdx.x; i < H * ta.NACC; i += blockDim.x) {
        const int t = i / ta.NACC, c = i % ta.NACC;
        ta.partials[ta.T + i] = chunk_sum_one(ta.bp + (long long)t * ta.nch * ta.NACC + c, ta.nch, ta.NACC);
    }
    for (int tr = 0; tr < ta.T; tr++) {
        const double m = row_min(ta.bmin, ta.nb, tr);
        if (threadIdx.x == 0) ta.partials[tr] = m;
    }
    __threadfence();
    __syncthreads();
    mjb_combine_args c = ta.c;
    if (c.n_shards > 1) c.partials = exchange_partials(ta.partials, ta.P, ta.peers, ta.rank, ta.seq, c.n_shards);
    else c.partials = ta.partials;
    softmax_combine_body(c, sm);
    __syncthreads();
    if (ta.action_out) for (int j = threadIdx.x; j < c.d; j += blockDim.x) ta.action_out[j] = c.mean[j];
    __syncthreads();
    if (ta.shift) {
        shift_mean_body(c.mean, H, c.d, ta.base_action, nullptr, &sm.comb[0][0]);
        if (ta.cov_shift_beta != 0.0 && (int)threadIdx.x < c.d) c.cov[threadIdx.x * c.d + threadIdx.x] += ta.cov_shift_beta * 1.0;
    }
    if (threadIdx.x == 0) *ta.counter = 0u;
}

__global__ void softmax_weights_kernel(const double* __restrict__ total, int K, const double* __restrict__ stats, int T,
                                       int t, double ninv, double* __restrict__ w) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const double m = stats[2 + T + t], S = stats[2 + t];
    w[k] = exp(ninv * total[(long long)t * K + k] - ninv * m) / S;
}

// ------------------------------------------------------------------------------ batched small MPPI
// one block per controller; sh[0..K) holds the trajectory costs, then the un-normalised weights
template <int D>
__global__ void __launch_bounds__(256) mppi_batched_kernel(mjb_mppi_batched_args a, GammaSeq G) {
    extern __shared__ double sh[];
    __shared__ double un[MJB_MAXH * MJB_MAXD];
    __shared__ double Ainv[MJB_MAXD][2 * MJB_MAXD];
    __shared__ double red[256];
    const int c = blockIdx.x, K = a.K, H = a.H;
    double* mean = a.mean + (long long)c * H * D;
    const long long kbase = (long long)c * K;
    if (a.control_cost) {
        if (threadIdx.x == 0) {
            for (int i = 0; i < D; i++)
                for (int j = 0; j < D; j++) { Ainv[i][j] = a.cov[i * D + j]; Ainv[i][D + j] = i == j ? 1.0 : 0.0; }
            for (int col = 0; col < D; col++) {
                int p = col;
                for (int r = col + 1; r < D; r++) if (fabs(Ainv[r][col]) > fabs(Ainv[p][col])) p = r;
                if (p != col) for (int j = 0; j < 2 * D; j++) { const double t = Ainv[col][j]; Ainv[col][j] = Ainv[p][j]; Ainv[p][j] = t; }
                const double inv = 1.0 / Ainv[col][col];
                for (int j = 0; j < 2 * D; j++) Ainv[col][j] *= inv;
                for (int r = 0; r < D; r++) if (r != col) {
                    const double f = Ainv[r][col];
                    for (int j = 0; j < 2 * D; j++) Ainv[r][j] -= f * Ainv[col][j];
                }
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < H * D; i += blockDim.x) {
            const int t = i / D, j = i % D;
            double s = 0.0;
            for (int l = 0; l < D; l++) s += mean[t * D + l] * Ainv[l][D + j];
            un[i] = s;
        }
        __syncthreads();
    }
    // trajectory cost per particle (reference order), block minimum
    double mn = INFINITY;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const long long gk = kbase + k;
        double S = 0.0, Sc = 0.0, ctg = 0.0, ccg = 0.0;
        for (int t = H - 1; t >= 0; t--) {
            const double cst = a.costs[gk * a.costs_sk + t * a.costs_st];
            if (G.raw) ctg = cst;
            else { S = __dadd_rn(S, __dmul_rn(G.g[t], cst)); ctg = __ddiv_rn(S, G.g[t]); }
            if (a.control_cost) {
                double cc = 0.0;
                for (int j = 0; j < D; j++) {
                    const double m = mean[t * D + j];
                    const double dl = a.actions[gk * a.act_sk + t * a.act_st + j * a.act_sj] - m;
                    cc += 0.5 * un[t * D + j] * (m + 2.0 * dl);
                }
                if (G.raw) ccg = cc;
                else { Sc = __dadd_rn(Sc, __dmul_rn(G.g[t], cc)); ccg = __ddiv_rn(Sc, G.g[t]); }
            }
        }
        const double tot = ctg + a.lam * ccg;
        sh[k] = tot;
        mn = fmin(mn, tot);
    }
    red[threadIdx.x] = mn;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) { if (threadIdx.x < s) red[threadIdx.x] = fmin(red[threadIdx.x], red[threadIdx.x + s]); __syncthreads(); }
    const double ninv = -1.0 / a.lam, xmax = ninv * red[0];
    __syncthreads();
    double ws = 0.0;
    for (int k = threadIdx.x; k < K; k += blockDim.x) { const double w = exp(ninv * sh[k] - xmax); sh[k] = w; ws += w; }
    red[threadIdx.x] = ws;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) { if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s]; __syncthreads(); }
    const double Ssum = red[0];
    // weighted mean: one (t,j) entry per thread, particles in index order
    for (int i = threadIdx.x; i < H * D; i += blockDim.x) {
        const int t = i / D, j = i % D;
        double s = 0.0;
        for (int k = 0; k < K; k++) s += sh[k] * a.actions[(kbase + k) * a.act_sk + t * a.act_st + j * a.act_sj];
        mean[i] = (1.0 - a.step_size) * mean[i] + a.step_size * (s / Ssum);
    }
    if (a.value && threadIdx.x == 0) a.value[c] = -a.lam * (xmax + log(Ssum / (double)K));
}

__global__ void shift_mean_batched_kernel(double* mean, int n, int H, int d, int base, const double* rnd) {
    // grid.x = controller; rows move up by one, so walk forward in t
    double* m = mean + (long long)blockIdx.x * H * d;
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
        const double last = m[(H - 1) * d + j];
        for (int t = 0; t < H - 1; t++) m[t * d + j] = m[(t + 1) * d + j];
        double v;
        if (base == MJB_BASE_NULL) v = 0.0;
        else if (base == MJB_BASE_REPEAT) v = last;
        else v = rnd[(long long)blockIdx.x * d + j];
        m[(H - 1) * d + j] = v;
    }
}

// ------------------------------------------------------------------------------ elite selection
// 8-pass MSB radix select on order-preserving keys, one block; ties at the threshold resolved by index.
__global__ void __launch_bounds__(1024) select_elites_kernel(const double* __restrict__ v, long long K, long long E,
                                                            unsigned char* __restrict__ flags, long long* __restrict__ ids) {
    __shared__ unsigned hist[256];
    __shared__ unsigned long long s_prefix, s_want;
    __shared__ long long s_scan[1024];
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) { s_prefix = 0; s_want = (unsigned long long)E; }
    __syncthreads();
    for (int pass = 0; pass < 8; pass++) {
        const int shift = 56 - 8 * pass;
        for (int b = tid; b < 256; b += nt) hist[b] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        const unsigned long long himask = pass == 0 ? 0ull : (~0ull << (shift + 8));
        for (long long i = tid; i < K; i += nt) {
            const unsigned long long key = enc_key(v[i]);
            if ((key & himask) == prefix) {
                // warp-aggregate: costs share their exponent byte, so whole warps hit one bin
                const unsigned b = (unsigned)(key >> shift) & 255u;
                const unsigned peers = __match_any_sync(__activemask(), b);
                if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&hist[b], (unsigned)__popc(peers));
            }
        }
        __syncthreads();
        if (tid == 0) {
            unsigned long long want = s_want, cum = 0;
            int b = 0;
            for (; b < 256; b++) { if (cum + hist[b] >= want) break; cum += hist[b]; }
            s_want = want - cum;
            s_prefix = prefix | ((unsigned long long)b << shift);
        }
        __syncthreads();
    }
    const unsigned long long thr = s_prefix;      // key of the E-th smallest value
    const long long need_eq = (long long)s_want;  // how many keys equal to it are elite (lowest indices first)
    // ordered pass over contiguous index ranges
    const long long per = (K + nt - 1) / nt, lo = (long long)tid * per, hi = lo + per < K ? lo + per : K;
    long long neq = 0;
    for (long long i = lo; i < hi; i++) neq += enc_key(v[i]) == thr;
    s_scan[tid] = neq;
    __syncthreads();
    if (tid == 0) { long long c = 0; for (int i = 0; i < nt; i++) { const long long x = s_scan[i]; s_scan[i] = c; c += x; } }
    __syncthreads();
    long long eq_before = s_scan[tid], nel = 0;
    __syncthreads();
    for (long long i = lo; i < hi; i++) {
        const unsigned long long key = enc_key(v[i]);
        bool e = key < thr;
        if (key == thr) { e = eq_before < need_eq; eq_before++; }
        flags[i] = e ? 1 : 0;
        nel += e;
    }
    if (ids) {
        s_scan[tid] = nel;
        __syncthreads();
        if (tid == 0) { long long c = 0; for (int i = 0; i < nt; i++) { const long long x = s_scan[i]; s_scan[i] = c; c += x; } }
        __syncthreads();
        long long o = s_scan[tid];
        for (long long i = lo; i < hi; i++) if (flags[i]) ids[o++] = i;
    }
}

#if defined(__CUDACC__)
// The same selection by a thread-block CLUSTER of 8 CTAs (sm_100a): every CTA keeps its eighth of the keys in shared
// memory (up to 16 384 keys = 128 KB per CTA), so the eight radix passes never touch L2 again; the per-pass
// histograms of the eight CTAs are summed through distributed shared memory behind cluster barriers, every CTA
// picks the bin redundantly.  One launch, ~10 us at K = 65 536 against ~250 us for the one-block form above (eight
// passes of 64 dependent L2 loads per thread).  Ties at the threshold go to the LOWER index, ids ascending.
#define MJB_SEL_NC 8
#define MJB_SEL_MAXKEYS 16384
__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <class T> __device__ __forceinline__ T ld_dsmem(const T* p, unsigned rank);
template <> __device__ __forceinline__ unsigned ld_dsmem<unsigned>(const unsigned* p, unsigned rank) {
    unsigned a = (unsigned)__cvta_generic_to_shared(p), ra, v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(ra) : "memory");
    return v;
}
template <> __device__ __forceinline__ long long ld_dsmem<long long>(const long long* p, unsigned rank) {
    unsigned a = (unsigned)__cvta_generic_to_shared(p), ra;
    long long v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    asm volatile("ld.shared::cluster.s64 %0, [%1];" : "=l"(v) : "r"(ra) : "memory");
    return v;
}
// exclusive prefix of one value per thread over the block (1024 threads); returns the prefix, *total = block sum
// (in every thread)
__device__ long long block_excl_scan(long long x, long long* total) {
    __shared__ long long wsum[32], s_total;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    long long inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const long long y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    __syncthreads();
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        long long w = wsum[lane], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const long long y = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += y; }
        wsum[lane] = winc - w;                   // exclusive prefix of the warp sums
        if (lane == 31) s_total = winc;
    }
    __syncthreads();
    *total = s_total;
    return wsum[wid] + inc - x;
}
__global__ void __cluster_dims__(MJB_SEL_NC, 1, 1) __launch_bounds__(1024)
select_elites_cluster_kernel(const double* __restrict__ v, long long K, long long E, unsigned char* __restrict__ flags,
                             long long* __restrict__ ids) {
    extern __shared__ unsigned long long keys[];
    __shared__ unsigned hist[256], tot[256];
    __shared__ unsigned long long s_prefix, s_want;
    __shared__ long long s_cnt[2];               // this CTA's keys equal to the threshold / elite keys
    const int tid = threadIdx.x;
    const unsigned rank = cluster_rank();
    const long long per = (K + MJB_SEL_NC - 1) / MJB_SEL_NC;
    const long long lo = (long long)rank * per;
    const int n = (int)(lo >= K ? 0 : (lo + per <= K ? per : K - lo));
    for (int i = tid; i < n; i += 1024) keys[i] = enc_key(v[lo + i]);
    if (tid == 0) { s_prefix = 0; s_want = (unsigned long long)E; }
    __syncthreads();
    for (int pass = 0; pass < 8; pass++) {
        const int shift = 56 - 8 * pass;
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        const unsigned long long himask = pass == 0 ? 0ull : (~0ull << (shift + 8));
        for (int i0 = 0; i0 < n; i0 += 1024) {
            const int i = i0 + tid;
            const bool in = i < n && ((keys[i] & himask) == prefix);
            // warp-aggregate: costs share their exponent byte, so whole warps hit one bin
            const unsigned b = in ? (unsigned)(keys[i] >> shift) & 255u : 256u;
            const unsigned peers = __match_any_sync(0xffffffffu, b);
            if (in && (tid & 31) == (__ffs(peers) - 1)) atomicAdd(&hist[b], (unsigned)__popc(peers));
        }
        __syncthreads();
        cluster_sync_all();                      // every CTA's histogram of this pass is complete
        if (tid < 256) {
            unsigned sum = 0;
#pragma unroll
            for (unsigned r = 0; r < MJB_SEL_NC; r++) sum += ld_dsmem(&hist[tid], r);
            tot[tid] = sum;
        }
        __syncthreads();
        if (tid < 32) {
            // bin holding the `want`-th smallest key: lane l owns bins 8l .. 8l+7
            unsigned long long part = 0;
#pragma unroll
            for (int b = 0; b < 8; b++) part += tot[8 * tid + b];
            unsigned long long inc = part;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned long long y = __shfl_up_sync(0xffffffffu, inc, o); if (tid >= o) inc += y; }
            const unsigned long long want = s_want;
            const unsigned hit = __ballot_sync(0xffffffffu, inc >= want);
            const int owner = __ffs(hit) - 1;
            __syncwarp();                            // every lane has read s_want before the owner overwrites it
            if (tid == owner) {
                unsigned long long cum = inc - part;
                int b = 0;
                for (; b < 8; b++) { if (cum + tot[8 * tid + b] >= want) break; cum += tot[8 * tid + b]; }
                s_want = want - cum;
                s_prefix = prefix | ((unsigned long long)(8 * tid + b) << shift);
            }
        }
        __syncthreads();
        cluster_sync_all();                      // nobody still reads my histogram when the next pass clears it
    }
    const unsigned long long thr = s_prefix;     // key of the E-th smallest value
    const long long need_eq = (long long)s_want; // how many keys equal to it are elite (lowest indices first)
    // ordered pass: thread t owns the contiguous keys [t*c, (t+1)*c) of this CTA
    const int c = (n + 1023) / 1024, a0 = tid * c < n ? tid * c : n, a1 = a0 + c < n ? a0 + c : n;
    long long neq = 0, nlt = 0;
    for (int i = a0; i < a1; i++) { neq += keys[i] == thr; nlt += keys[i] < thr; }
    long long tot_eq, tot_lt;
    const long long eq_before_cta = block_excl_scan(neq, &tot_eq);
    const long long lt_before_cta = block_excl_scan(nlt, &tot_lt);
    if (tid == 0) { s_cnt[0] = tot_eq; s_cnt[1] = tot_lt; }
    __syncthreads();
    cluster_sync_all();
    long long eq_base = 0, lt_base = 0;
    for (unsigned r = 0; r < rank; r++) { eq_base += ld_dsmem(&s_cnt[0], r); lt_base += ld_dsmem(&s_cnt[1], r); }
    long long eq_before = eq_base + eq_before_cta;
    // ids position of this thread's first elite: elites before it = keys below the threshold + admitted ties
    long long o = (lt_base + lt_before_cta) + (eq_before < need_eq ? eq_before : need_eq);
    for (int i = a0; i < a1; i++) {
        const unsigned long long key = keys[i];
        bool e = key < thr;
        if (key == thr) { e = eq_before < need_eq; eq_before++; }
        flags[lo + i] = e ? 1 : 0;
        if (e && ids) ids[o++] = lo + i;
    }
    cluster_sync_all();                          // no CTA exits while a peer may still read its shared memory
}
#endif  // __CUDACC__

__global__ void __launch_bounds__(1024) argmin_kernel(const double* __restrict__ v, long long K, long long* out_idx,
                                                     double* out_val) {
    __shared__ unsigned long long sk[1024];
    __shared__ long long si[1024];
    unsigned long long bk = ~0ull;
    long long bi = -1;
    for (long long i = threadIdx.x; i < K; i += blockDim.x) {
        const unsigned long long key = enc_key(v[i]);
        if (key < bk) { bk = key; bi = i; }          // strided ascending i: first occurrence wins within a thread
    }
    sk[threadIdx.x] = bk; si[threadIdx.x] = bi;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const unsigned long long ok = sk[threadIdx.x + s];
            const long long oi = si[threadIdx.x + s];
            if (oi >= 0 && (ok < sk[threadIdx.x] || (ok == sk[threadIdx.x] && (si[threadIdx.x] < 0 || oi < si[threadIdx.x])))) {
                sk[threadIdx.x] = ok; si[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out_idx[0] = si[0]; if (out_val) out_val[0] = si[0] >= 0 ? v[si[0]] : INFINITY; }
}

// pass-1 / pass-2 combine of elite moments (cem.py:71-86)
__global__ void elite_combine_kernel(mjb_elite_combine_args a) {
    const int H = a.H, d = a.d, P1 = 1 + H * d + d, P2 = d * (d + 1) / 2;
    __shared__ double n_s, mu_s[MJB_MAXD];
    if (threadIdx.x == 0) {
        double n = 0.0;
        for (int r = 0; r < a.n_shards; r++) n += a.partial1[(long long)r * P1];
        n_s = n;
    }
    __syncthreads();
    const double n = n_s;
    if (threadIdx.x < d) {
        double s = 0.0;
        for (int r = 0; r < a.n_shards; r++) s += a.partial1[(long long)r * P1 + 1 + H * d + threadIdx.x];
        mu_s[threadIdx.x] = s / (n * H);
        a.mu[threadIdx.x] = mu_s[threadIdx.x];
    }
    __syncthreads();
    if (!a.partial2) return;
    for (int i = threadIdx.x; i < d * d; i += blockDim.x) {
        const int r = i / d, c = i % d;
        const int hi = r > c ? r : c, lo = r > c ? c : r;
        double s = 0.0;
        for (int q = 0; q < a.n_shards; q++) s += a.partial2[(long long)q * P2 + hi * (hi + 1) / 2 + lo];
        double upd;
        if (a.full_cov) upd = s / (n * H - 1.0);           // np.cov, ddof = 1
        else upd = r == c ? s / (n * H) : 0.0;             // np.diag(np.var), ddof = 0
        a.cov[i] = (1.0 - a.step_size) * a.cov[i] + a.step_size * upd;
    }
    for (int i = threadIdx.x; i < H * d; i += blockDim.x) {
        double s = 0.0;
        for (int q = 0; q < a.n_shards; q++) s += a.partial1[(long long)q * P1 + 1 + i];
        a.mean[i] = (1.0 - a.step_size) * a.mean[i] + a.step_size * (s / n);
    }
}

// partial1 = [ n | sum_a (H,d) | sum_delta (d) ] from the per-t reduced block (NACC = 1 + d)
__global__ void elite_pack1_kernel(const double* __restrict__ part, const double* __restrict__ mean, int H, int d,
                                   double* __restrict__ out) {
    const int NACC = 1 + d;
    if (threadIdx.x == 0) out[0] = part[0];
    for (int i = threadIdx.x; i < H * d; i += blockDim.x) out[1 + i] = part[(i / d) * NACC + 1 + i % d];
    __syncthreads();
    if (threadIdx.x < d) {
        double s = 0.0;
        for (int t = 0; t < H; t++) s += part[t * NACC + 1 + threadIdx.x] - part[t * NACC] * mean[t * d + threadIdx.x];
        out[1 + H * d + threadIdx.x] = s;
    }
}
// partial2 = sum over t of the lower-triangular second moments (NACC = 1 + d + d(d+1)/2)
__global__ void elite_pack2_kernel(const double* __restrict__ part, int H, int d, double* __restrict__ out) {
    const int P2 = d * (d + 1) / 2, NACC = 1 + d + P2;
    for (int c = threadIdx.x; c < P2; c += blockDim.x) {
        double s = 0.0;
        for (int t = 0; t < H; t++) s += part[t * NACC + 1 + d + c];
        out[c] = s;
    }
}

__global__ void blend_best_kernel(const double* __restrict__ actions, long long sk, long long st, long long sj,
                                  const long long* best, long long k_offset, int K, int H, int d, double step,
                                  double* mean) {
    const long long b = best[0] - k_offset;
    if (b < 0 || b >= K) return;
    for (int i = threadIdx.x; i < H * d; i += blockDim.x)
        mean[i] = (1.0 - step) * mean[i] + step * actions[b * sk + (i / d) * st + (i % d) * sj];
}

// ------------------------------------------------------------------------------ batched RandomShooting / CEM instances
// one block of 256 threads per instance; dynamic shared memory: K cost-to-go values + K flags
__global__ void __launch_bounds__(256) instances_batched_kernel(mjb_instances_args a, GammaSeq G) {
    extern __shared__ double sh[];
    __shared__ unsigned hist[256];
    __shared__ unsigned long long s_prefix, s_want;
    __shared__ long long s_scan[256];
    __shared__ double red[256], mu[MJB_MAXD];
    __shared__ long long redi[256];
    const int c = blockIdx.x, K = a.K, H = a.H, D = a.d, tid = threadIdx.x, nt = blockDim.x;
    unsigned char* flag = (unsigned char*)(sh + K);
    double* mean = a.mean + (long long)c * H * D;
    const long long kbase = (long long)c * K;
    // cost-to-go at t = 0 in the reference's operation order
    double vsum = 0.0;
    for (int k = tid; k < K; k += nt) {
        double S = 0.0, ctg = 0.0;
        for (int t = H - 1; t >= 0; t--) {
            const double cst = a.costs[(kbase + k) * a.costs_sk + t * a.costs_st];
            if (G.raw) ctg = cst;
            else { S = __dadd_rn(S, __dmul_rn(G.g[t], cst)); ctg = __ddiv_rn(S, G.g[t]); }
        }
        sh[k] = ctg;
        vsum += ctg;
    }
    red[tid] = vsum;
    __syncthreads();
    for (int s2 = nt / 2; s2 > 0; s2 >>= 1) { if (tid < s2) red[tid] += red[tid + s2]; __syncthreads(); }
    if (a.value && tid == 0) a.value[c] = red[0] / (double)K;
    __syncthreads();
    if (a.mode == MJB_INST_RS) {
        // argmin, first occurrence
        unsigned long long bk = ~0ull; long long bi = -1;
        for (int k = tid; k < K; k += nt) { const unsigned long long key = enc_key(sh[k]); if (key < bk) { bk = key; bi = k; } }
        ((unsigned long long*)red)[tid] = bk; redi[tid] = bi;
        __syncthreads();
        for (int s2 = nt / 2; s2 > 0; s2 >>= 1) {
            if (tid < s2) {
                const unsigned long long ok = ((unsigned long long*)red)[tid + s2]; const long long oi = redi[tid + s2];
                const unsigned long long mk = ((unsigned long long*)red)[tid];
                if (oi >= 0 && (ok < mk || (ok == mk && (redi[tid] < 0 || oi < redi[tid])))) { ((unsigned long long*)red)[tid] = ok; redi[tid] = oi; }
            }
            __syncthreads();
        }
        const long long best = redi[0];
        if (a.ids && tid == 0) a.ids[c] = best;
        if (a.apply)
            for (int i = tid; i < H * D; i += nt)
                mean[i] = (1.0 - a.step_size) * mean[i] + a.step_size * a.actions[(kbase + best) * a.act_sk + (i / D) * a.act_st + (i % D) * a.act_sj];
        return;
    }
    // ---- CEM: radix select of the num_elite smallest keys, ties to the lower index
    if (tid == 0) { s_prefix = 0; s_want = (unsigned long long)a.num_elite; }
    __syncthreads();
    for (int pass = 0; pass < 8; pass++) {
        const int shift = 56 - 8 * pass;
        hist[tid] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix, himask = pass == 0 ? 0ull : (~0ull << (shift + 8));
        for (int k = tid; k < K; k += nt) {
            const unsigned long long key = enc_key(sh[k]);
            if ((key & himask) == prefix) atomicAdd(&hist[(unsigned)(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned long long want = s_want, cum = 0;
            int b = 0;
            for (; b < 256; b++) { if (cum + hist[b] >= want) break; cum += hist[b]; }
            s_want = want - cum;
            s_prefix = prefix | ((unsigned long long)b << shift);
        }
        __syncthreads();
    }
    const unsigned long long thr = s_prefix;
    const long long need_eq = (long long)s_want;
    const int per = (K + nt - 1) / nt, lo = tid * per < K ? tid * per : K, hi = lo + per < K ? lo + per : K;
    long long neq = 0;
    for (int k = lo; k < hi; k++) neq += enc_key(sh[k]) == thr;
    s_scan[tid] = neq;
    __syncthreads();
    if (tid == 0) { long long cc = 0; for (int i = 0; i < nt; i++) { const long long x = s_scan[i]; s_scan[i] = cc; cc += x; } }
    __syncthreads();
    long long eq_before = s_scan[tid], nel = 0;
    __syncthreads();
    for (int k = lo; k < hi; k++) {
        const unsigned long long key = enc_key(sh[k]);
        bool e = key < thr;
        if (key == thr) { e = eq_before < need_eq; eq_before++; }
        flag[k] = e ? 1 : 0;
        nel += e;
    }
    s_scan[tid] = nel;
    __syncthreads();
    if (tid == 0) { long long cc = 0; for (int i = 0; i < nt; i++) { const long long x = s_scan[i]; s_scan[i] = cc; cc += x; } }
    __syncthreads();
    if (a.ids) { long long o = s_scan[tid]; for (int k = lo; k < hi; k++) if (flag[k]) a.ids[(long long)c * a.num_elite + o++] = k; }
    if (!a.apply) return;
    const double E = (double)a.num_elite, n = E * (double)H;
    // pooled mean of the elite deltas (delta = action - old mean), per dimension, particles in index order
    if (tid < D) {
        double s2 = 0.0;
        for (int k = 0; k < K; k++) if (flag[k])
            for (int t = 0; t < H; t++) s2 += a.actions[(kbase + k) * a.act_sk + t * a.act_st + tid * a.act_sj] - mean[t * D + tid];
        mu[tid] = s2 / n;
    }
    __syncthreads();
    // covariance update: np.var (ddof 0) on the diagonal, or np.cov (ddof 1) in full
    double* cov = a.cov + (long long)c * D * D;
    for (int i = tid; i < D * D; i += nt) {
        const int r = i / D, q = i % D;
        double upd = 0.0;
        if (a.mode == MJB_INST_CEM_FULL || r == q) {
            double s2 = 0.0;
            for (int k = 0; k < K; k++) if (flag[k])
                for (int t = 0; t < H; t++) {
                    const double dr = a.actions[(kbase + k) * a.act_sk + t * a.act_st + r * a.act_sj] - mean[t * D + r] - mu[r];
                    const double dq = a.actions[(kbase + k) * a.act_sk + t * a.act_st + q * a.act_sj] - mean[t * D + q] - mu[q];
                    s2 += dr * dq;
                }
            upd = a.mode == MJB_INST_CEM_FULL ? s2 / (n - 1.0) : s2 / n;
        }
        cov[i] = (1.0 - a.step_size) * cov[i] + a.step_size * upd;
    }
    __syncthreads();                                     // every delta above used the OLD mean
    for (int i = tid; i < H * D; i += nt) {
        const int t = i / D, j = i % D;
        double s2 = 0.0;
        for (int k = 0; k < K; k++) if (flag[k]) s2 += a.actions[(kbase + k) * a.act_sk + t * a.act_st + j * a.act_sj];
        mean[i] = (1.0 - a.step_size) * mean[i] + a.step_size * (s2 / E);
    }
}
__global__ void cov_add_diag_batched_kernel(double* cov, int n, int d, double beta, const double* v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * d) { const int c = i / d, j = i % d; cov[(long long)c * d * d + j * d + j] += beta * (v ? v[j] : 1.0); }
}

// ------------------------------------------------------------------------------ PFMPC resampling
__global__ void __launch_bounds__(256) seq_cumsum_kernel(const double* __restrict__ w, long long M, double* __restrict__ cs) {
    __shared__ double tile[2048];
    __shared__ double carry;
    if (threadIdx.x == 0) carry = 0.0;
    for (long long base = 0; base < M; base += 2048) {
        const int n = (int)(M - base < 2048 ? M - base : 2048);
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) tile[i] = w[base + i];
        __syncthreads();
        if (threadIdx.x == 0) {
            double c = carry;
            for (int i = 0; i < n; i++) { c = __dadd_rn(c, tile[i]); tile[i] = c; }   // reference order: c += w[i]
            carry = c;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) cs[base + i] = tile[i];
    }
}
__global__ void resample_search_kernel(const double* __restrict__ cs, long long M, double r, long long* __restrict__ idx) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const double u = __dadd_rn(r, __ddiv_rn((double)m, (double)M));       // r + m*1.0/M*1.0
    long long res;
    if (!(0.0 < u)) res = M - 1;                                          // loop never runs: act_seq[-1]
    else {
        long long lo = 0, hi = M;                                         // first i with cs[i] >= u
        while (lo < hi) { const long long mid = (lo + hi) >> 1; if (cs[mid] >= u) hi = mid; else lo = mid + 1; }
        res = lo < M ? lo : M - 1;
    }
    idx[m] = res;
}
// Parallel form of the same resampling, exact by certification.  The reference compares a SEQUENTIALLY accumulated
// FP64 prefix sum with u_m = r + m/M; a parallel scan rounds differently, but both sums of M non-negative terms stay
// within gamma_M * S (S = total weight) of the exact prefix, so they differ by at most eps = 4 M 2^-53 S.  An index
// found in the parallel prefix is therefore the reference's index whenever u_m is more than eps away from the two
// neighbouring prefix values; only if some u_m is closer (probability ~1e-6 per particle) the flag goes up and the
// two sequential kernels below re-do the whole resampling in the reference's order.  Bit-exact always, ~15 us
// instead of ~300 (65 536 dependent additions on one thread) in all but one call in a few dozen.
// aux: [0] flag (as unsigned), [1] eps
__global__ void __launch_bounds__(1024) par_cumsum_kernel(const double* __restrict__ w, long long M, double* __restrict__ cs,
                                                          double* __restrict__ aux) {
    __shared__ double wsum[32], s_total;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long c = (M + 1023) / 1024, a0 = tid * c < M ? tid * c : M, a1 = a0 + c < M ? a0 + c : M;
    double loc = 0.0;
    for (long long i = a0; i < a1; i++) loc += w[i];
    double inc = loc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const double x = wsum[lane];
        double winc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double y = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += y; }
        wsum[lane] = winc - x;
        if (lane == 31) s_total = winc;
    }
    __syncthreads();
    double run = wsum[wid] + (inc - loc);
    for (long long i = a0; i < a1; i++) { run += w[i]; cs[i] = run; }
    if (tid == 0) {
        *(unsigned int*)aux = 0u;
        aux[1] = 4.0 * (double)M * 1.1102230246251565e-16 * s_total;
    }
}
__global__ void resample_search_certified_kernel(const double* __restrict__ cs, long long M, double r,
                                                 long long* __restrict__ idx, double* __restrict__ aux) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const double eps = aux[1];
    const double u = __dadd_rn(r, __ddiv_rn((double)m, (double)M));       // r + m*1.0/M*1.0
    long long res;
    bool sure = true;
    if (!(0.0 < u)) res = M - 1;                                          // loop never runs: act_seq[-1]
    else {
        long long lo = 0, hi = M;                                         // first i with cs[i] >= u
        while (lo < hi) { const long long mid = (lo + hi) >> 1; if (cs[mid] >= u) hi = mid; else lo = mid + 1; }
        if (lo < M) sure = (cs[lo] - u > eps) && (lo == 0 || u - cs[lo - 1] > eps);
        else sure = u - cs[M - 1] > eps;
        res = lo < M ? lo : M - 1;
    }
    idx[m] = res;
    if (!sure) atomicOr((unsigned int*)aux, 1u);
}
__global__ void __launch_bounds__(256) seq_cumsum_if_kernel(const double* __restrict__ w, long long M, double* __restrict__ cs,
                                                            const double* __restrict__ aux) {
    if (*(const unsigned int*)aux == 0u) return;
    __shared__ double tile[2048];
    __shared__ double carry;
    if (threadIdx.x == 0) carry = 0.0;
    for (long long base = 0; base < M; base += 2048) {
        const int n = (int)(M - base < 2048 ? M - base : 2048);
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) tile[i] = w[base + i];
        __syncthreads();
        if (threadIdx.x == 0) {
            double c = carry;
            for (int i = 0; i < n; i++) { c = __dadd_rn(c, tile[i]); tile[i] = c; }   // reference order: c += w[i]
            carry = c;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) cs[base + i] = tile[i];
    }
}
__global__ void resample_search_if_kernel(const double* __restrict__ cs, long long M, double r, long long* __restrict__ idx,
                                          const double* __restrict__ aux) {
    if (*(const unsigned int*)aux == 0u) return;
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const double u = __dadd_rn(r, __ddiv_rn((double)m, (double)M));
    long long res;
    if (!(0.0 < u)) res = M - 1;
    else {
        long long lo = 0, hi = M;
        while (lo < hi) { const long long mid = (lo + hi) >> 1; if (cs[mid] >= u) hi = mid; else lo = mid + 1; }
        res = lo < M ? lo : M - 1;
    }
    idx[m] = res;
}
__global__ void gather_kernel(const double* __restrict__ in, long long isk, long long ist, long long isj,
                              const long long* __restrict__ idx, int K, int H, int d, double* __restrict__ out,
                              long long osk, long long ost, long long osj) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int tj = blockIdx.y;
    if (k >= K) return;
    const int t = tj / d, j = tj % d;
    out[k * osk + t * ost + j * osj] = in[idx[k] * isk + t * ist + j * isj];
}

__global__ void sub_mean_kernel(const double* __restrict__ x, long long sk, long long st, long long sj,
                                const double* __restrict__ mean, int K, int H, int d, double* __restrict__ out,
                                long long osk, long long ost, long long osj) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int tj = blockIdx.y;
    if (k >= K) return;
    const int t = tj / d, j = tj % d;
    out[k * osk + t * ost + j * osj] = x[k * sk + t * st + j * sj] - mean[tj];
}

// ------------------------------------------------------------------------------ batched small PFMPC
// One block per independent PFMPC instance (BASELINE config 5): cost-to-go, softmax weights, the reference's
// sequential cumulative sum, the systematic-resampling searches, the gather into the new particle set and
// its mean -- particle_filter_controller.py:92-113,159-174 per instance.  sh[0..K): trajectory costs ->
// weights -> cumulative sums; ish[0..K): resampled indices.
__global__ void __launch_bounds__(256) pf_batched_kernel(mjb_pf_batched_args a, GammaSeq G) {
    extern __shared__ double sh[];
    __shared__ double red[256];
    const int c = blockIdx.x, K = a.K, H = a.H, d = a.d;
    int* ish = (int*)(sh + K);
    const long long kbase = (long long)c * K;
    double mn = INFINITY;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const long long gk = kbase + k;
        double S = 0.0, ctg = 0.0;
        for (int t = H - 1; t >= 0; t--) {
            const double cst = a.costs[gk * a.costs_sk + t * a.costs_st];
            if (G.raw) ctg = cst;
            else { S = __dadd_rn(S, __dmul_rn(G.g[t], cst)); ctg = __ddiv_rn(S, G.g[t]); }
        }
        sh[k] = ctg;
        mn = fmin(mn, ctg);
    }
    red[threadIdx.x] = mn;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) { if (threadIdx.x < s) red[threadIdx.x] = fmin(red[threadIdx.x], red[threadIdx.x + s]); __syncthreads(); }
    const double ninv = -1.0 / a.lam, xmax = ninv * red[0];
    __syncthreads();
    double ws = 0.0;
    for (int k = threadIdx.x; k < K; k += blockDim.x) { const double w = exp(ninv * sh[k] - xmax); sh[k] = w; ws += w; }
    red[threadIdx.x] = ws;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) { if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s]; __syncthreads(); }
    const double Ssum = red[0];
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const double w = sh[k] / Ssum;
        sh[k] = w;
        if (a.weights) a.weights[kbase + k] = w;
    }
    __syncthreads();
    if (threadIdx.x == 0) {                       // reference order: c += w[i]
        double cs = 0.0;
        for (int i = 0; i < K; i++) { cs = __dadd_rn(cs, sh[i]); sh[i] = cs; }
    }
    __syncthreads();
    const double r = a.r[c];
    for (int m = threadIdx.x; m < K; m += blockDim.x) {
        const double u = __dadd_rn(r, __ddiv_rn((double)m, (double)K));    // r + m*1.0/M*1.0
        int res;
        if (!(0.0 < u)) res = K - 1;                                       // loop never runs: act_seq[-1]
        else {
            int lo = 0, hi = K;                                            // first i with cs[i] >= u
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (sh[mid] >= u) hi = mid; else lo = mid + 1; }
            res = lo < K ? lo : K - 1;
        }
        ish[m] = res;
        if (a.idx) a.idx[kbase + m] = res;
    }
    __syncthreads();
    // gather + mean of the resampled set: one (t,j) column per thread, particles in index order
    for (int i = threadIdx.x; i < H * d; i += blockDim.x) {
        const int t = i / d, j = i % d;
        double s = 0.0;
        for (int m = 0; m < K; m++) {
            const double v = a.samples[(kbase + ish[m]) * a.s_sk + t * a.s_st + j * a.s_sj];
            a.out[(kbase + m) * a.o_sk + t * a.o_st + j * a.o_sj] = v;
            s += v;
        }
        a.mean[(long long)c * H * d + i] = s / (double)K;
    }
}

__global__ void sub_mean_batched_kernel(const double* __restrict__ x, long long sk, long long st, long long sj,
                                        const double* __restrict__ mean, int K, int ppc, int H, int d,
                                        double* __restrict__ out, long long osk, long long ost, long long osj) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int tj = blockIdx.y;
    if (k >= K) return;
    const int t = tj / d, j = tj % d;
    out[k * osk + t * ost + j * osj] = x[k * sk + t * st + j * sj] - mean[(k / ppc) * (long long)H * d + tj];
}

// ------------------------------------------------------------------------------ shifts
__global__ void shift_mean_kernel(double* mean, int H, int d, int base, const double* rnd) {
    extern __shared__ double buf[];
    shift_mean_body(mean, H, d, base, rnd, buf);
}
__global__ void cov_add_diag_kernel(double* cov, int d, double beta, const double* v) {
    if (threadIdx.x < d) cov[threadIdx.x * d + threadIdx.x] += beta * (v ? v[threadIdx.x] : 1.0);
}
__global__ void pf_shift_kernel(double* s, long long sk, long long st, long long sj, const double* dl, long long dk,
                                long long dt, long long dj, int K, int H, int d, int base, const double* rnd) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    for (int j = 0; j < d; j++) {
        double prev = 0.0;
        for (int t = 0; t < H; t++) {
            // particle_filter_controller.py:133-139: roll left (last column keeps its value), then add noise
            const double src = t < H - 1 ? s[k * sk + (t + 1) * st + j * sj] : s[k * sk + t * st + j * sj];
            double v = src + dl[k * dk + t * dt + j * dj];
            if (t == H - 1) {
                if (base == MJB_BASE_NULL) v = 0.0;
                else if (base == MJB_BASE_REPEAT) v = H >= 2 ? prev : v;
                else v = rnd[j];
            }
            s[k * sk + t * st + j * sj] = v;
            prev = v;
        }
    }
}

template <int D>
static int launch_reduce(int wmode, int cmode, dim3 grid, cudaStream_t s, int K, int H, const double* total, int T,
                         const double* bmin, int nb, double ninv, const unsigned char* flags, const double* actions,
                         long long sk, long long st, long long sj, const double* mean, const double* mu, double* out) {
#define MJB_L(W, C) weighted_reduce_kernel<D, W, C><<<grid, MJB_RB, 0, s>>>(K, H, total, T, bmin, nb, ninv, flags, actions, sk, st, sj, mean, mu, out)
    if (wmode == 0 && cmode == 0) MJB_L(0, 0);
    else if (wmode == 0 && cmode == 1) MJB_L(0, 1);
    else if (wmode == 0 && cmode == 2) MJB_L(0, 2);
    else if (wmode == 1 && cmode == 0) MJB_L(1, 0);
    else if (wmode == 1 && cmode == 2) MJB_L(1, 2);
    else if (wmode == 2 && cmode == 0) MJB_L(2, 0);
    else return set_error(MJB_EINVAL, "unsupported reduction mode %d/%d", wmode, cmode);
#undef MJB_L
    return MJB_OK;
}
static int dispatch_reduce(int d, int wmode, int cmode, dim3 grid, cudaStream_t s, int K, int H, const double* total, int T,
                           const double* bmin, int nb, double ninv, const unsigned char* flags, const double* actions,
                           long long sk, long long st, long long sj, const double* mean, const double* mu, double* out) {
    switch (d) {
#define MJB_CASE(D) case D: return launch_reduce<D>(wmode, cmode, grid, s, K, H, total, T, bmin, nb, ninv, flags, actions, sk, st, sj, mean, mu, out);
        MJB_CASE(1) MJB_CASE(2) MJB_CASE(3) MJB_CASE(4) MJB_CASE(5) MJB_CASE(6) MJB_CASE(7) MJB_CASE(8)
#undef MJB_CASE
    }
    return set_error(MJB_EINVAL, "d_action=%d not in 1..%d", d, MJB_MAXD);
}
static int ncov_of(int d, int cov_mode) { return cov_mode == MJB_COV_NONE ? 0 : (cov_mode == MJB_COV_DIAG ? d : d * (d + 1) / 2); }
static int nchunks_of(int K) { return (K + MJB_CHUNK - 1) / MJB_CHUNK; }

}  // namespace mjb

using namespace mjb;

extern "C" int mjb_cost_to_go(const double* costs, long long sk, long long st, const double* gamma_seq_host, int K, int H,
                              double* out, long long osk, long long ost, void* stream) {
    MJB_REQUIRE(costs && out && gamma_seq_host, "mjb_cost_to_go: null pointer");
    MJB_REQUIRE(K >= 0 && H >= 1, "mjb_cost_to_go: bad shape");
    GammaSeq G;
    int rc = load_gamma(G, gamma_seq_host, H);
    if (rc) return rc;
    if (K == 0) return MJB_OK;
    cost_to_go_kernel<<<(K + 255) / 256, 256, 0, (cudaStream_t)stream>>>(costs, sk, st, G, K, H, out, osk, ost);
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}

extern "C" long long mjb_softmax_scratch_doubles(int K, int H, int d, int cov_mode) {
    const int NACC = 1 + d + ncov_of(d, cov_mode);
    return sm_off_bmin(K, H, d, NACC) + (long long)H * sm_nb(K) + 64;
}
extern "C" int mjb_softmax_partial_doubles(int H, int d, int time_based, int cov_mode) {
    return (time_based ? H : 1) + H * (1 + d + ncov_of(d, cov_mode));
}

namespace mjb {
static int softmax_check(const mjb_softmax_args* a, const char* who) {
    MJB_REQUIRE(a && a->costs && a->actions && a->mean && a->gamma_seq && a->total && a->scratch && a->partials,
                "%s: null pointer", who);
    MJB_REQUIRE(a->K >= 1 && a->H >= 1, "%s: bad shape K=%d H=%d", who, a->K, a->H);
    MJB_REQUIRE(a->d >= 1 && a->d <= MJB_MAXD, "d_action=%d not in 1..%d", a->d, MJB_MAXD);
    MJB_REQUIRE(a->lam > 0.0, "lam must be positive");
    MJB_REQUIRE(!a->control_cost || a->cov, "control cost needs cov");
    MJB_REQUIRE(!(a->time_based && a->cov_mode != MJB_COV_NONE), "time-based weights have no covariance update");
    if (a->cov_mode < 0 || a->cov_mode > 2)
        return set_error(MJB_EINVAL, "Unidentified covariance type in update_distribution");   // gaussian_dmd.py:85
    MJB_REQUIRE(a->returns == MJB_RETURNS_CTG || a->returns == MJB_RETURNS_TD_LAMBDA, "%s: unknown returns mode %d", who, a->returns);
    const bool td = a->returns == MJB_RETURNS_TD_LAMBDA;
    MJB_REQUIRE(!td || a->H == 1 || a->td_weight_seq, "TD(lambda) returns need td_weight_seq");
    MJB_REQUIRE(!(td && a->cov_mode != MJB_COV_NONE), "TD(lambda) returns have no covariance update");
    return MJB_OK;
}
// [u_n when the control cost is on,] trajectory costs + block minima
static int launch_traj_cost(const mjb_softmax_args* a, cudaStream_t s) {
    const bool td = a->returns == MJB_RETURNS_TD_LAMBDA;
    GammaSeq G, W;
    int rc = load_gamma(G, a->gamma_seq, a->H);
    if (rc) return rc;
    W.raw = 0;
    if (td && a->H > 1) { rc = load_gamma(W, a->td_weight_seq, a->H - 1); if (rc) return rc; }
    const int T = a->time_based ? a->H : 1;
    const int NACC = 1 + a->d + ncov_of(a->d, a->cov_mode);
    if (a->control_cost) softmax_prep_kernel<<<1, 128, 0, s>>>(*a, T);
    double* bmin = a->scratch + sm_off_bmin(a->K, a->H, a->d, NACC);
    const int tgrid = sm_nb(a->K);
    switch (a->d) {
#define MJB_TC(D, TDV, T1V) traj_cost_kernel<D, TDV, T1V><<<tgrid, MJB_RB, 0, s>>>(*a, G, W, T, bmin)
#define MJB_CASE(D) case D: if (td) { if (T == 1) MJB_TC(D, true, true); else MJB_TC(D, true, false); } \
                            else { if (T == 1) MJB_TC(D, false, true); else MJB_TC(D, false, false); } break;
        MJB_CASE(1) MJB_CASE(2) MJB_CASE(3) MJB_CASE(4) MJB_CASE(5) MJB_CASE(6) MJB_CASE(7) MJB_CASE(8)
#undef MJB_CASE
#undef MJB_TC
    }
    return MJB_OK;
}
}  // namespace mjb

extern "C" int mjb_softmax_partials(const mjb_softmax_args* a, void* stream) {
    int rc = softmax_check(a, "mjb_softmax_partials");
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int T = a->time_based ? a->H : 1;
    const int NACC = 1 + a->d + ncov_of(a->d, a->cov_mode);
    const int nch = nchunks_of(a->K), nb = sm_nb(a->K);
    rc = launch_traj_cost(a, s);
    if (rc) return rc;
    double* bp = a->scratch + sm_off_partials(a->H, a->d);
    const double* bmin = a->scratch + sm_off_bmin(a->K, a->H, a->d, NACC);
    rc = dispatch_reduce(a->d, 0, a->cov_mode, dim3(nch, a->H), s, a->K, a->H, a->total, T, bmin, nb, -1.0 / a->lam, nullptr,
                         a->actions, a->act_sk, a->act_st, a->act_sj, a->mean, nullptr, bp);
    if (rc) return rc;
    softmax_finalize_kernel<<<a->H, MJB_RB, 0, s>>>(bp, nch, NACC, T, bmin, nb, a->partials);
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}

// The whole softmax update -- phase 1, [peer exchange,] phase 2, next action, hot-start shift -- in TWO launches
// (three with the control cost): trajectory costs, then the weighted reduction whose last block runs the tail.
// Same results as mjb_softmax_partials + mjb_softmax_combine / mjb_softmax_exchange_combine + mjb_shift_mean +
// mjb_cov_add_diag, bit for bit.  a->scratch[0] is the "blocks done" counter: zero it once after allocating.
extern "C" int mjb_softmax_update_fused(const mjb_softmax_args* a, const mjb_combine_args* c, void* const* peer_bufs_dev,
                                        int rank, unsigned long long seq, double* action_out, int shift, int base_action,
                                        double cov_shift_beta, void* stream) {
    int rc = softmax_check(a, "mjb_softmax_update_fused");
    if (rc) return rc;
    MJB_REQUIRE(c && c->mean, "mjb_softmax_update_fused: the combine must apply the update (mean is NULL)");
    MJB_REQUIRE(c->H == a->H && c->d == a->d && c->time_based == a->time_based && c->cov_mode == a->cov_mode,
                "mjb_softmax_update_fused: softmax / combine shapes disagree");
    MJB_REQUIRE(c->n_shards >= 1 && c->n_shards <= MJB_RB && (c->n_shards == 1 || peer_bufs_dev),
                "mjb_softmax_update_fused: %d shards need the peer-memory exchange buffers", c->n_shards);
    MJB_REQUIRE(c->cov_mode == MJB_COV_NONE || c->cov, "covariance update needs cov");
    MJB_REQUIRE(base_action == MJB_BASE_NULL || base_action == MJB_BASE_REPEAT || !shift,
                "mjb_softmax_update_fused: base_action 'random' needs a host-drawn row; use the separate entry points");
    MJB_REQUIRE(cov_shift_beta == 0.0 || c->cov, "mjb_softmax_update_fused: cov_shift_beta needs the covariance");
    cudaStream_t s = (cudaStream_t)stream;
    const int T = a->time_based ? a->H : 1;
    const int NACC = 1 + a->d + ncov_of(a->d, a->cov_mode);
    const int nch = nchunks_of(a->K), nb = sm_nb(a->K);
    rc = launch_traj_cost(a, s);
    if (rc) return rc;
    double* bp = a->scratch + sm_off_partials(a->H, a->d);
    const double* bmin = a->scratch + sm_off_bmin(a->K, a->H, a->d, NACC);
    TailArgs ta;
    ta.c = *c;
    ta.partials = a->partials; ta.bp = bp; ta.bmin = bmin;
    ta.nb = nb; ta.nch = nch; ta.NACC = NACC; ta.T = T; ta.P = T + a->H * NACC;
    ta.counter = (unsigned int*)a->scratch;
    ta.peers = peer_bufs_dev; ta.rank = rank; ta.seq = seq;
    ta.action_out = action_out; ta.shift = shift; ta.base_action = base_action; ta.cov_shift_beta = cov_shift_beta;
    const dim3 grid(nch, a->H);
    const double ninv = -1.0 / a->lam;
    switch (a->d) {
#define MJB_RT(D, C) softmax_reduce_tail_kernel<D, C><<<grid, MJB_RB, 0, s>>>(a->K, a->H, a->total, T, bmin, nb, ninv, a->actions, a->act_sk, a->act_st, a->act_sj, a->mean, bp, ta)
#define MJB_CASE(D) case D: if (a->cov_mode == 0) MJB_RT(D, 0); else if (a->cov_mode == 1) MJB_RT(D, 1); else MJB_RT(D, 2); break;
        MJB_CASE(1) MJB_CASE(2) MJB_CASE(3) MJB_CASE(4) MJB_CASE(5) MJB_CASE(6) MJB_CASE(7) MJB_CASE(8)
#undef MJB_CASE
#undef MJB_RT
    }
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}

extern "C" int mjb_softmax_combine(const mjb_combine_args* a, void* stream) {
    MJB_REQUIRE(a && a->partials && (a->mean || a->stats), "mjb_softmax_combine: null pointer");
    MJB_REQUIRE(a->H >= 1 && a->H <= MJB_MAXH && a->d >= 1 && a->d <= MJB_MAXD && a->n_shards >= 1, "mjb_softmax_combine: bad shape");
    MJB_REQUIRE(a->cov_mode == MJB_COV_NONE || a->cov || !a->mean, "covariance update needs cov");
    softmax_combine_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(*a);
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}

extern "C" int mjb_softmax_weights(const double* total, int K, const double* stats, int t, double lam, double* w_out,
                                   void* stream) {
    MJB_REQUIRE(total && stats && w_out && K >= 1 && lam > 0.0, "mjb_softmax_weights: bad argument");
    // T is implied by the caller's stats layout: this entry point serves the non-time-based case (T = 1)
    softmax_weights_kernel<<<(K + 255) / 256, 256, 0, (cudaStream_t)stream>>>(total, K, stats, 1, t, -1.0 / lam, w_out);
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}

extern "C" int mjb_select_elites(const double* ctg0, long long K_global, long long num_elite, unsigned char* flags,
                                 long long* ids, void* scratch, void* stream) {
    (void)scratch;
    MJB_REQUIRE(ctg0 && flags, "mjb_select_elites: null pointer");
    MJB_REQUIRE(K_global >= 1 && num_elite >= 1 && num_elite <= K_global, "num_elite=%lld must be in 1..K=%lld", num_elite, K_global);
#if defined(__CUDACC__)
    if (K_global <= (long long)MJB_SEL_NC * MJB_SEL_MAXKEYS && K_global >= 4096) {
        // keys resident in the shared memory of an 8-CTA cluster (one launch, no L2 re-reads)
        const long long per = (K_global + MJB_SEL_NC - 1) / MJB_SEL_NC;
        const size_t smem = sizeof(unsigned long long) * (size_t)per;
        MJB_CUDA(cudaFuncSetAttribute(select_elites_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(sizeof(unsigned long long) * MJB_SEL_MAXKEYS)));
        select_elites_cluster_kernel<<<MJB_SEL_NC, 1024, smem, (cudaStream_t)stream>>>(ctg0, K_global, num_elite, flags, ids);
        MJB_CUDA(cudaGetLastError());
        return MJB_OK;
    }
#endif
    select_elites_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(ctg0, K_global, num_elite, flags, ids);
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}

extern "C" int mjb_argmin(const double* ctg0, long long K, long long* out_index, double* out_value, void* stream) {
    MJB_REQUIRE(ctg0 && out_index && K >= 1, "mjb_argmin: bad argument");
    argmin_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(ctg0, K, out_index, out_value);
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}

extern "C" long long mjb_elite_scratch_doubles(int K, int H, int d) {
    const long long NACC = 1 + d + d * (d + 1) / 2;
    return (long long)H * nchunks_of(K) * NACC + (long long)H * NACC + 64;
}

static int elite_moments(const mjb_elite_args* a, int pass, void* stream) {
    MJB_REQUIRE(a && a->flags && a->actions && a->mean && a->scratch && a->partial, "mjb_elite_moments: null pointer");
    MJB_REQUIRE(a->K >= 1 && a->H >= 1 && a->H <= MJB_MAXH && a->d >= 1 && a->d <= MJB_MAXD, "mjb_elite_moments: bad shape");
    MJB_REQUIRE(pass == 1 || a->mu, "mjb_elite_moments2 needs the pooled mean");
    cudaStream_t s = (cudaStream_t)stream;
    const int nch = nchunks_of(a->K);
    const int cmode = pass == 1 ? 0 : 2;
    const int NACC = 1 + a->d + (pass == 1 ? 0 : a->d * (a->d + 1) / 2);
    double* bp = a->scratch;
    double* part = a->scratch + (long long)a->H * nch * NACC;
    int rc = dispatch_reduce(a->d, 1, cmode, dim3(nch, a->H), s, a->K, a->H, nullptr, 1, nullptr, 0, 0.0, a->flags, a->actions,
                             a->act_sk, a->act_st, a->act_sj, a->mean, a->mu, bp);
    if (rc) return rc;
    chunk_sum_kernel<<<a->H, 64, 0, s>>>(bp, nch, NACC, part);
    if (pass == 1) elite_pack1_kernel<<<1, 256, 0, s>>>(part, a->mean, a->H, a->d, a->partial);
    else elite_pack2_kernel<<<1, 64, 0, s>>>(part, a->H, a->d, a->partial);
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}
extern "C" int mjb_elite_moments1(const mjb_elite_args* a, void* stream) { return elite_moments(a, 1, stream); }
extern "C" int mjb_elite_moments2(const mjb_elite_args* a, void* stream) { return elite_moments(a, 2, stream); }

extern "C" int mjb_elite_combine(const mjb_elite_combine_args* a, void* stream) {
    MJB_REQUIRE(a && a->partial1 && a->mu, "mjb_elite_combine: null pointer");
    MJB_REQUIRE(!a->partial2 || (a->mean && a->cov), "mjb_elite_combine: mean and cov are required with partial2");
    MJB_REQUIRE(a->H >= 1 && a->d >= 1 && a->d <= MJB_MAXD && a->n_shards >= 1, "mjb_elite_combine: bad shape");
    elite_combine_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(*a);
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}

extern "C" int mjb_blend_best(const double* actions, long long sk, long long st, long long sj, const long long* best_index,
                              long long k_offset, int K, int H, int d, double step_size, double* mean, void* stream) {
    MJB_REQUIRE(actions && best_index && mean, "mjb_blend_best: null pointer");
    blend_best_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(actions, sk, st, sj, best_index, k_offset, K, H, d, step_size, mean);
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}

extern "C" int mjb_resample_indices(const double* weights, long long M, double r, double* cumsum_scratch,
                                    long long* idx_out, void* stream) {
    MJB_REQUIRE(weights && cumsum_scratch && idx_out && M >= 1, "mjb_resample_indices: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    if (M < 4096) {            // small sets: the sequential scan is already short
        seq_cumsum_kernel<<<1, 256, 0, s>>>(weights, M, cumsum_scratch);
        resample_search_kernel<<<(unsigned)((M + 255) / 256), 256, 0, s>>>(cumsum_scratch, M, r, idx_out);
    } else {
        // parallel prefix + certified search; the sequential pair runs only when a sample sits within the rounding
        // distance of a bin edge (scratch: M doubles of prefix + 2 of flag / eps)
        double* aux = cumsum_scratch + M;
        const unsigned g = (unsigned)((M + 255) / 256);
        par_cumsum_kernel<<<1, 1024, 0, s>>>(weights, M, cumsum_scratch, aux);
        resample_search_certified_kernel<<<g, 256, 0, s>>>(cumsum_scratch, M, r, idx_out, aux);
        seq_cumsum_if_kernel<<<1, 256, 0, s>>>(weights, M, cumsum_scratch, aux);
        resample_search_if_kernel<<<g, 256, 0, s>>>(cumsum_scratch, M, r, idx_out, aux);
    }
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}

extern "C" int mjb_gather_particles(const double* in, long long isk, long long ist, long long isj, const long long* idx,
                                    int K, int H, int d, double* out, long long osk, long long ost, long long osj,
                                    void* stream) {
    MJB_REQUIRE(in && idx && out && K >= 1 && H >= 1 && d >= 1, "mjb_gather_particles: bad argument");
    MJB_REQUIRE(in != out, "mjb_gather_particles: in-place gather is not supported");
    gather_kernel<<<dim3((K + 255) / 256, H * d), 256, 0, (cudaStream_t)stream>>>(in, isk, ist, isj, idx, K, H, d, out, osk, ost, osj);
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}

extern "C" int mjb_particle_mean(const double* x, long long sk, long long st, long long sj, int K, int H, int d,
                                 double* scratch, double* out, void* stream);
namespace mjb {
__global__ void particle_mean_pack_kernel(const double* part, int H, int d, int K, double* out) {
    for (int i = threadIdx.x; i < H * d; i += blockDim.x) out[i] = part[(i / d) * (1 + d) + 1 + i % d] / (double)K;
}
}
extern "C" int mjb_particle_mean(const double* x, long long sk, long long st, long long sj, int K, int H, int d,
                                 double* scratch, double* out, void* stream) {
    MJB_REQUIRE(x && scratch && out && K >= 1 && H >= 1 && d >= 1 && d <= MJB_MAXD, "mjb_particle_mean: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    const int nch = nchunks_of(K), NACC = 1 + d;
    double* bp = scratch;
    double* part = scratch + (long long)H * nch * NACC;
    int rc = dispatch_reduce(d, 2, 0, dim3(nch, H), s, K, H, nullptr, 1, nullptr, 0, 0.0, nullptr, x, sk, st, sj, nullptr, nullptr, bp);
    if (rc) return rc;
    chunk_sum_kernel<<<H, 64, 0, s>>>(bp, nch, NACC, part);
    particle_mean_pack_kernel<<<1, 256, 0, s>>>(part, H, d, K, out);
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}

extern "C" int mjb_shift_mean(double* mean, int H, int d, int base_action, const double* random_row, void* stream) {
    MJB_REQUIRE(mean && H >= 1 && d >= 1, "mjb_shift_mean: bad argument");
    if (base_action < 0 || base_action > 2)
        return set_error(MJB_ENOTIMPL, "invalid option for base action during shift");   // olgaussian_mpc.py:129
    MJB_REQUIRE(base_action != MJB_BASE_RANDOM || random_row, "base_action 'random' needs a random row");
    shift_mean_kernel<<<1, 256, sizeof(double) * H * d, (cudaStream_t)stream>>>(mean, H, d, base_action, random_row);
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}

extern "C" int mjb_cov_add_diag(double* cov, int d, double beta, const double* v, void* stream) {
    MJB_REQUIRE(cov && d >= 1 && d <= 32, "mjb_cov_add_diag: bad argument");
    cov_add_diag_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(cov, d, beta, v);
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}

extern "C" int mjb_pf_shift(double* samples, long long sk, long long st, long long sj, const double* delta, long long dk,
                            long long dt, long long dj, int K, int H, int d, int base_action, const double* random_row,
                            void* stream) {
    MJB_REQUIRE(samples && delta && K >= 1 && H >= 1 && d >= 1, "mjb_pf_shift: bad argument");
    if (base_action < 0 || base_action > 2)
        return set_error(MJB_ENOTIMPL, "invalid option for base action during shift");   // particle_filter_controller.py:148
    MJB_REQUIRE(base_action != MJB_BASE_RANDOM || random_row, "base_action 'random' needs a random row");
    pf_shift_kernel<<<(K + 127) / 128, 128, 0, (cudaStream_t)stream>>>(samples, sk, st, sj, delta, dk, dt, dj, K, H, d, base_action, random_row);
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}

extern "C" int mjb_particle_sub_mean(const double* x, long long sk, long long st, long long sj, const double* mean, int K,
                                     int H, int d, double* out, long long osk, long long ost, long long osj, void* stream) {
    MJB_REQUIRE(x && mean && out && K >= 1 && H >= 1 && d >= 1, "mjb_particle_sub_mean: bad argument");
    sub_mean_kernel<<<dim3((K + 255) / 256, H * d), 256, 0, (cudaStream_t)stream>>>(x, sk, st, sj, mean, K, H, d, out, osk, ost, osj);
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}

extern "C" int mjb_mppi_update_batched(const mjb_mppi_batched_args* a, void* stream) {
    MJB_REQUIRE(a && a->costs && a->actions && a->mean && a->gamma_seq, "mjb_mppi_update_batched: null pointer");
    MJB_REQUIRE(a->n_ctrl >= 1 && a->K >= 1 && a->K <= 4096 && a->H >= 1, "mjb_mppi_update_batched: bad shape (K per controller must be <= 4096)");
    MJB_REQUIRE(a->d >= 1 && a->d <= MJB_MAXD, "d_action=%d not in 1..%d", a->d, MJB_MAXD);
    MJB_REQUIRE(a->lam > 0.0, "lam must be positive");
    MJB_REQUIRE(!a->control_cost || a->cov, "control cost needs cov");
    GammaSeq G;
    int rc = load_gamma(G, a->gamma_seq, a->H);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = sizeof(double) * a->K;
    switch (a->d) {
#define MJB_CASE(D) case D: mppi_batched_kernel<D><<<a->n_ctrl, 256, smem, s>>>(*a, G); break;
        MJB_CASE(1) MJB_CASE(2) MJB_CASE(3) MJB_CASE(4) MJB_CASE(5) MJB_CASE(6) MJB_CASE(7) MJB_CASE(8)
#undef MJB_CASE
    }
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}

extern "C" int mjb_pf_update_batched(const mjb_pf_batched_args* a, void* stream) {
    MJB_REQUIRE(a && a->costs && a->samples && a->gamma_seq && a->r && a->out && a->mean, "mjb_pf_update_batched: null pointer");
    MJB_REQUIRE(a->n_ctrl >= 1 && a->K >= 1 && a->K <= 4096 && a->H >= 1, "mjb_pf_update_batched: bad shape (K per controller must be <= 4096)");
    MJB_REQUIRE(a->d >= 1 && a->d <= MJB_MAXD, "d_action=%d not in 1..%d", a->d, MJB_MAXD);
    MJB_REQUIRE(a->lam > 0.0, "lam must be positive");
    MJB_REQUIRE(a->samples != a->out, "mjb_pf_update_batched: in-place resampling is not supported");
    GammaSeq G;
    int rc = load_gamma(G, a->gamma_seq, a->H);
    if (rc) return rc;
    const size_t smem = (sizeof(double) + sizeof(int)) * (size_t)a->K;
    // K = 4096 needs 48 KB of dynamic shared memory on top of the kernel's static 2 KB: above the default limit
    if (smem > 40 * 1024)
        MJB_CUDA(cudaFuncSetAttribute(pf_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pf_batched_kernel<<<a->n_ctrl, 256, smem, (cudaStream_t)stream>>>(*a, G);
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}

extern "C" int mjb_particle_sub_mean_batched(const double* x, long long sk, long long st, long long sj, const double* mean,
                                             int n_ctrl, int K, int H, int d, double* out, long long osk, long long ost,
                                             long long osj, void* stream) {
    MJB_REQUIRE(x && mean && out && n_ctrl >= 1 && K >= 1 && H >= 1 && d >= 1, "mjb_particle_sub_mean_batched: bad argument");
    const long long Kt = (long long)n_ctrl * K;
    sub_mean_batched_kernel<<<dim3((unsigned)((Kt + 255) / 256), H * d), 256, 0, (cudaStream_t)stream>>>(
        x, sk, st, sj, mean, (int)Kt, K, H, d, out, osk, ost, osj);
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}

extern "C" int mjb_shift_mean_batched(double* mean, int n, int H, int d, int base_action, const double* random_rows, void* stream) {
    MJB_REQUIRE(mean && n >= 1 && H >= 1 && d >= 1, "mjb_shift_mean_batched: bad argument");
    if (base_action < 0 || base_action > 2)
        return set_error(MJB_ENOTIMPL, "invalid option for base action during shift");
    MJB_REQUIRE(base_action != MJB_BASE_RANDOM || random_rows, "base_action 'random' needs random rows");
    shift_mean_batched_kernel<<<n, 32, 0, (cudaStream_t)stream>>>(mean, n, H, d, base_action, random_rows);
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}

extern "C" int mjb_softmax_exchange_combine(const mjb_combine_args* a, const double* local_partial, void* const* peer_bufs_dev,
                                            int rank, unsigned long long seq, void* stream) {
    MJB_REQUIRE(a && local_partial && peer_bufs_dev && (a->mean || a->stats), "mjb_softmax_exchange_combine: null pointer");
    MJB_REQUIRE(a->H >= 1 && a->H <= MJB_MAXH && a->d >= 1 && a->d <= MJB_MAXD && a->n_shards >= 1 && a->n_shards <= 64,
                "mjb_softmax_exchange_combine: bad shape");
    MJB_REQUIRE(rank >= 0 && rank < a->n_shards && seq >= 1, "mjb_softmax_exchange_combine: bad rank / sequence number");
    const int P = mjb_softmax_partial_doubles(a->H, a->d, a->time_based, a->cov_mode);
    softmax_exchange_combine_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(*a, local_partial, P, peer_bufs_dev, rank, seq);
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}

extern "C" int mjb_instances_update_batched(const mjb_instances_args* a, void* stream) {
    MJB_REQUIRE(a && a->costs && a->actions && a->gamma_seq, "mjb_instances_update_batched: null pointer");
    MJB_REQUIRE(!a->apply || a->mean, "mjb_instances_update_batched: apply needs the mean");
    MJB_REQUIRE(a->n_ctrl >= 1 && a->K >= 1 && a->K <= 4096 && a->H >= 1, "mjb_instances_update_batched: bad shape (K <= 4096)");
    MJB_REQUIRE(a->d >= 1 && a->d <= MJB_MAXD, "d_action=%d not in 1..%d", a->d, MJB_MAXD);
    MJB_REQUIRE(a->mode == MJB_INST_RS || a->mode == MJB_INST_CEM_DIAG || a->mode == MJB_INST_CEM_FULL, "mjb_instances_update_batched: unknown mode %d", a->mode);
    if (a->mode != MJB_INST_RS) {
        MJB_REQUIRE(a->num_elite >= 1 && a->num_elite <= a->K, "num_elite=%lld must be in 1..K=%d", a->num_elite, a->K);
        MJB_REQUIRE(!a->apply || a->cov, "mjb_instances_update_batched: CEM needs the covariances");
        MJB_REQUIRE(a->mode != MJB_INST_CEM_FULL || a->num_elite * a->H >= 2, "np.cov needs at least two elite samples");
    }
    GammaSeq G;
    int rc = load_gamma(G, a->gamma_seq, a->H);
    if (rc) return rc;
    const size_t smem = sizeof(double) * a->K + (size_t)a->K + 8;
    instances_batched_kernel<<<a->n_ctrl, 256, smem, (cudaStream_t)stream>>>(*a, G);
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}

extern "C" int mjb_cov_add_diag_batched(double* cov, int n, int d, double beta, const double* v, void* stream) {
    MJB_REQUIRE(cov && n >= 1 && d >= 1 && d <= MJB_MAXD, "mjb_cov_add_diag_batched: bad argument");
    cov_add_diag_batched_kernel<<<(n * d + 127) / 128, 128, 0, (cudaStream_t)stream>>>(cov, n, d, beta, v);
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}
