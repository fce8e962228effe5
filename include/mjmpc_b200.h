/*
 * mjmpc_b200 -- C ABI of the B200-native sampling-MPC inner loop.
 *
 * The reference (mohakbhardwaj/mjmpc) is pure Python and has no FFI: its backend boundary is
 * two callables injected on the controller (mjmpc/control/controller.py:152-175) plus the
 * numeric bodies of the controller classes.  Each entry point below names the reference
 * function(s) it replaces.  All pointers are plain host or device pointers (stated per
 * argument), all tensors are FP64, strides are in ELEMENTS, `stream` is a cudaStream_t passed
 * as void* (NULL = default stream).  Every function returns MJB_OK (0) or an error code;
 * mjb_last_error() returns the message of the last failure on the calling thread.
 * Nothing here allocates tensors for the caller or frees caller memory.
 */
#ifndef MJMPC_B200_H
#define MJMPC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define MJB_OK 0
#define MJB_EINVAL 1   /* bad argument (maps to ValueError / AssertionError in the Python shim) */
#define MJB_ECUDA 2    /* CUDA runtime failure */
#define MJB_ENOTIMPL 3 /* unsupported option (maps to NotImplementedError) */

#define MJB_MODEL_NPARAM 167 /* doubles per model instance, layout in mjmpc_b200/csrc/chain_model.h */
#define MJB_STATE_DIM 17     /* qpos(7) qvel(7) target_pos(3) */
#define MJB_OBS_DIM 20       /* reacher_env.py:41-47: qpos, qvel, hand, hand-target */

const char* mjb_last_error(void);
int mjb_version(void);
/* sm count, clock (kHz) and compute capability of `device`; fails if it is not sm_100. */
int mjb_device_info(int device, int* sm_count, int* clock_khz, int* cc_major, int* cc_minor);
/* Measured FP64 FMA throughput of the CUDA cores (TFLOP/s, FMA = 2 FLOP): the roofline
 * denominator of the rollout kernel.  Launches a register-only FMA kernel, best of 4 timed runs. */
int mjb_fp64_peak(int device, int blocks_per_sm, int iters, double* tflops_out, double* ms_out);

/* ---- model: replaces MuJoCo's compiled mjModel for sawyer.xml (reacher_env.py:21) -------- */
typedef struct mjb_model mjb_model;
/* host_params: n_instances x MJB_MODEL_NPARAM (host).  Instances are the per-worker models of
 * SubprocVecEnv.randomize_dynamics (subproc_vec_env.py:304-312). */
int mjb_model_create(const double* host_params, int n_instances, int device, mjb_model** out);
int mjb_model_update(mjb_model* m, int first_instance, int n, const double* host_params, void* stream);
int mjb_model_n_instances(const mjb_model* m);
int mjb_model_destroy(mjb_model* m);

/* ---- rollout: replaces GymEnvWrapper.rollout (gym_env_wrapper.py:89-156) around
 * Reacher7DOFEnv.step (reacher_env.py:29-39) and SubprocVecEnv.rollout (subproc_vec_env.py:128-186).
 * Particle k uses state/mean row k / particles_per_ctrl and model instance
 * (k / particles_per_model) % n_instances.  All tensor pointers are DEVICE pointers. */
typedef struct {
    int K;                       /* particles in this launch (this GPU's shard) */
    int H;                       /* horizon (env steps) */
    int particles_per_ctrl;      /* K for one controller; K / n for n batched controllers */
    int particles_per_model;     /* K / n_workers in the reference's worker layout */
    const double* state;         /* (n_ctrl, MJB_STATE_DIM) */
    const double* mean;          /* (n_ctrl, H, 7) row-major (closed_loop: policy weights, see below) */
    const double* noise;         /* (K, H, 7) by strides, or NULL for the mean sequence only */
    long long noise_sk, noise_st, noise_sj;
    double* costs;               /* (K, H) = -reward */
    long long costs_sk, costs_st;
    double* actions;             /* (K, H, 7) unclipped mean+noise (gym_env_wrapper.py:151), or NULL */
    long long act_sk, act_st, act_sj;
    double* qv_traj;             /* (K, H, 14) row-major qpos,qvel after each env step, or NULL */
    double* next_obs;            /* (K, H, MJB_OBS_DIM) row-major, or NULL */
    int* ncon;                   /* (K,) substeps with >=1 active constraint row, or NULL */
    /* Fused K2: with noise == NULL and noise_cov != NULL the kernel generates the noise itself (same Philox
     * counters, covariance factor and filter as mjb_generate_noise -> bit-identical samples) and the noise
     * tensor never exists in HBM.  Fields mirror mjb_noise_args. */
    const double* noise_cov;     /* (7,7) device covariance, or NULL */
    unsigned long long noise_seed, noise_offset;
    const long long* noise_step_ptr;
    double noise_beta0, noise_beta1, noise_beta2;
    long long noise_k_offset, noise_K_global;
    int noise_zero_last;         /* particle noise_K_global-1 gets noise = -mean (mean row of its controller) */
    /* mode="closed_loop_linear" of GymEnvWrapper.rollout (gym_env_wrapper.py:135-136): with closed_loop != 0,
     * `mean` is (n_ctrl, MJB_OBS_DIM + 1, 7) row-major -- the weights of a linear policy -- and the control of
     * step t is mean' [obs_t; 1] + noise[k,t], obs_t the observation before the step (the first one from fresh
     * kinematics at the set state, reacher_env.py:88-99). */
    int closed_loop;
} mjb_rollout_args;
int mjb_rollout_reacher(const mjb_model* m, const mjb_rollout_args* a, void* stream);
/* Largest per-launch particle count that takes the ROLE-SPLIT instantiation of the rollout kernel (four warps
 * per 32 particles: bias forces | mass matrix + Euler solve | constraint rows + Newton | controls + cost; for
 * the latency-bound small launches of sharded controllers -- the reference's small per-worker batches,
 * subproc_vec_env.py:161-186).  Larger launches, and launches that ask for observations, closed-loop control or
 * in-kernel noise, run one thread per particle.  Returns the previous value; a negative argument only queries.
 * Default 8192, the measured cross-over (or the MJB_SPLIT_MAX_K environment variable); 0 switches the split kernel off. */
int mjb_rollout_split_max_k(int new_value);

/* ---- K9 pendulum rollout: replaces PendulumEnv.step (mjmpc/envs/basic/pendulum.py:33-50) under
 * GymEnvWrapper.rollout.  state: (n_ctrl, 2) theta, thetadot; mean (n_ctrl, H, 1); d_action = 1.
 * states_out (K,H,2) row-major or NULL. */
typedef struct {
    int K, H, particles_per_ctrl;
    const double* state;
    const double* mean;
    const double* noise; long long noise_sk, noise_st;
    double* costs; long long costs_sk, costs_st;
    double* actions; long long act_sk, act_st;
    double* states_out;
} mjb_pendulum_args;
int mjb_rollout_pendulum(const mjb_pendulum_args* a, void* stream);

/* ---- linear-quadratic toy env: replaces LQREnv.step (mjmpc/envs/basic/lqr.py:31-35) under
 * GymEnvWrapper.rollout: cost_t = x'Qx + u'Ru on the pre-step state, x <- Ax + Bu, u = mean[t] + noise[k,t]
 * (no clipping).  A (n,n), B (n,d), Q (n,n), R (d,d) row-major DEVICE; state (n_ctrl, n); mean (n_ctrl, H, d);
 * states_out (K,H,n) row-major (the post-step states) or NULL. */
#define MJB_LQR_MAXN 8
#define MJB_LQR_MAXD 8
typedef struct {
    int K, H, n, d, particles_per_ctrl;
    const double* A; const double* B; const double* Q; const double* R;
    const double* state;
    const double* mean;
    const double* noise; long long noise_sk, noise_st, noise_sj;
    double* costs; long long costs_sk, costs_st;
    double* actions; long long act_sk, act_st, act_sj;
    double* states_out;
} mjb_lqr_args;
int mjb_rollout_lqr(const mjb_lqr_args* a, void* stream);

/* ---- K11 runtime-parameterised TREE rollout: replaces MuJoCo's mj_step under
 * gym.envs.mujoco.MujocoEnv.do_simulation for the reference's hinge / slide models with a forward-progress reward --
 * SwimmerEnv.step (mjmpc/envs/basic/swimmer.py:10-19: 'swimmer.xml', frame_skip 4) and the same shape of step in
 * HalfCheetahEnv.step (half_cheetah.py:10-19) -- inside GymEnvWrapper.rollout (mjmpc/envs/gym_env_wrapper.py:123-153).
 *   reward_t = w_fwd * (q[fwd_dof] after - before) / (frame_skip * timestep) - w_ctrl * |a_t|^2,  cost = -reward
 *   a_t = mean[t] + noise[k, t] (recorded unclipped, wrapper :150; MuJoCo clamps the control to ctrlrange)
 *   observation = (qpos[obs_qpos_start:], qvel)   (swimmer.py:21-24: 2; half_cheetah.py:21-25: 1)
 * The model is the link-per-dof block compiled from MJCF by mjmpc_b200/envs/mjcf_tree.py (layout: csrc/tree_model.h,
 * reported by mjb_tree_layout): link_params (nv, LK_STRIDE) doubles, link_ints (nv, LI_STRIDE) ints, globals
 * (TG_STRIDE) doubles, HOST pointers, copied.  Dynamics: hinge / slide joints with any axis, anchor and body
 * orientation, several joints per body, gravity, joint springs / dampers / armature, mj_passive's inertia-box fluid
 * forces (option density / viscosity), soft joint limits (solreflimit / solimplimit), motors with ctrlrange, mj_Euler,
 * contacts of planar mechanisms (below).  No frictionloss, tendons, free / ball joints (the compiler rejects such models). */
typedef struct mjb_tree_model mjb_tree_model;
mjb_tree_model* mjb_tree_model_create(int nv, int nu, const double* link_params, const int* link_ints,
                                      const double* globals, const double* planar_params, const int* planar_anc,
                                      const double* planar_gravity, int n_contacts, const int* contact_ints,
                                      const double* contact_params, int n_instances, int device);
                                      /* NULL + mjb_last_error on failure.  n_instances > 1: link_params (n, nv, LK_STRIDE) and
                                       * planar_params (n, nv, PK_STRIDE) hold one model per worker -- the perturbed copies of
                                       * randomize_dynamics (subproc_vec_env.py:304-312, gym_env_wrapper.py:367-416): same topology,
                                       * different masses / inertias / dampings */
void mjb_tree_model_destroy(mjb_tree_model* m);
void mjb_tree_layout(int* out60);   /* the 60 layout constants of csrc/tree_model.h + rollout_tree_planar.cuh, in order */
/* PLANAR MECHANISMS (all hinge axes parallel, all slides perpendicular to them -- swimmer.xml and half_cheetah.xml are):
 * with the three planar blocks of mjcf_tree.pack_planar -- planar_params (nv, PK_STRIDE), planar_anc (nv) ancestor bit
 * masks, planar_gravity (2); NULL otherwise -- rollouts of models of 3 to 9 dofs run the planar instantiation (3-vectors
 * in place of spatial 6-vectors, everything in registers).  Same results to rounding; mjb_tree_use_planar(0) forces
 * the general instantiation (tests compare the two), a negative argument only queries; returns the previous value.
 * CONTACTS (planar mechanisms only): n_contacts candidate pairs from mjcf_tree.pack_planar_contacts -- contact_ints
 * (n, 3): kind, link 1 (-1 = world), link 2; contact_params (n, CT_STRIDE) -- world plane against the end spheres of a
 * capsule (half_cheetah.xml's floor) and capsule against capsule (swimmer.xml's non-adjacent links): MuJoCo's
 * mjc_PlaneCapsule / mjc_CapsuleCapsule detection, condim-3 pyramidal friction rows (n +- mu t), R = 2 mu^2 (1 - imp) / imp
 * (1 + mu^2)(invweight_1 + invweight_2), solved with the limit rows: active-set iteration around one factorisation,
 * Newton with an exact line search as the fall-back.  Instantiated for the two shapes of the reference's models (7 dofs
 * in series, 9-dof trees); mjb_tree_model_create refuses contacts on anything else (MJB_ENOTIMPL). */
int mjb_tree_use_planar(int on);
typedef struct {
    int K, H, frame_skip, particles_per_ctrl;
    int particles_per_model;        /* K / number of workers: particle k runs model k / particles_per_model (K for one model) */
    int fwd_dof, obs_qpos_start;
    double w_fwd, w_ctrl;
    const double* state;            /* (n_ctrl, 2 nv): qpos, qvel (swimmer.py:33-49 get / set_env_state) */
    const double* mean;             /* (n_ctrl, H, nu) */
    const double* noise; long long noise_sk, noise_st, noise_sj;     /* (K, H, nu) any strides, or NULL */
    double* costs; long long costs_sk, costs_st;
    double* actions; long long act_sk, act_st, act_sj;               /* or NULL */
    double* states_out;             /* (K, H, 2 nv) row-major, state AFTER each env step, or NULL */
    double* next_obs;               /* (K, H, 2 nv - obs_qpos_start) row-major, or NULL */
    int* nefc;                      /* (K,) constraint rows (limits + 4 per contact) summed over the rollout's substeps, or NULL */
} mjb_tree_rollout_args;
int mjb_rollout_tree(const mjb_tree_model* m, const mjb_tree_rollout_args* a, void* stream);

/* ---- K2 noise: replaces generate_noise (mjmpc/utils/control_utils.py:24-34) and the
 * use_zero_control_seq overwrite (mjmpc/control/olgaussian_mpc.py:110-111).  Philox4x32-10 keyed by
 * `seed`, counters (k_offset + k, low32(offset), t, high32(offset) << 8 | pair): results do not
 * depend on how particles are sharded.  All pointers are device pointers. */
typedef struct {
    int K, H, d;
    long long k_offset;          /* global index of this shard's first particle */
    long long K_global;          /* total particles over all shards (zero_last targets K_global-1) */
    unsigned long long seed;     /* controller seed_val */
    unsigned long long offset;   /* low 32 bits: num_steps-like counter; high 32 bits: stream id */
    const long long* step_ptr;   /* optional DEVICE counter replacing the low 32 bits of offset (lets a captured
                                    CUDA graph advance the noise stream without re-recording), or NULL */
    const double* cov;           /* (d,d) covariance, Cholesky-factorised inside the kernel */
    double beta0, beta1, beta2;  /* filter_coeffs */
    int zero_last;               /* 1: particle K_global-1 gets noise = -mean (zero control sequence) */
    const double* neg_mean;      /* (H,d) mean sequence, needed when zero_last */
    double* out;                 /* (K,H,d) by strides */
    long long out_sk, out_st, out_sj;
    /* batched controller instances with their OWN covariance (CEM with batch_size > 1): particle k draws from
     * cov + ((k_offset + k) / particles_per_cov) * cov_stride.  cov_stride == 0 (default): one covariance for all.
     * particles_per_cov must be a multiple of 32 (one covariance per 32-particle tile). */
    long long cov_stride;
    int particles_per_cov;
} mjb_noise_args;
int mjb_generate_noise(const mjb_noise_args* a, void* stream);

/* ---- K3 cost-to-go: replaces cost_to_go (mjmpc/utils/control_utils.py:37-46), bit-exact:
 * x_t = gamma_t*c_t, reverse sequential sum, divide by gamma_t; returns costs unchanged if any
 * gamma_t == 0.  gamma_seq is a HOST pointer (H,). */
int mjb_cost_to_go(const double* costs, long long sk, long long st, const double* gamma_seq_host, int K, int H,
                   double* out, long long out_sk, long long out_st, void* stream);

/* ---- K4/K6 exponential-utility reductions: MPPI._update_distribution/_exp_util/_control_costs
 * (mjmpc/control/mppi.py:69-111), DMDMPC._update_distribution/_exp_util
 * (mjmpc/control/gaussian_dmd.py:65-104), PFMPC._exp_util (particle_filter_controller.py:104-113).
 * Phase 1 reduces this GPU's particles to a partial vector
 *     [ m[T] | acc[H][NACC] ],  T = time_based ? H : 1,  NACC = 1 + d + ncov,
 *     acc[t] = ( sum_k w, sum_k w a[k,t,0..d), sum_k w delta delta' (diag: d, full: d(d+1)/2) ),
 *     w = exp(-(total_k - m)/lam), m = min_k total_k
 * phase 2 combines the partial vectors of all shards (rank order, deterministic) and writes the
 * smoothed mean / covariance.  With one GPU, n_shards = 1. */
#define MJB_COV_NONE 0
#define MJB_COV_DIAG 1
#define MJB_COV_FULL 2
#define MJB_RETURNS_CTG 0
#define MJB_RETURNS_TD_LAMBDA 1
typedef struct {
    int K, H, d;
    const double* costs; long long costs_sk, costs_st;
    const double* actions; long long act_sk, act_st, act_sj;
    const double* mean;          /* (H,d) device */
    const double* cov;           /* (d,d) device; only read when control_cost */
    const double* gamma_seq;     /* (H,) HOST */
    double lam;
    int control_cost;            /* MPPI alpha == 0 */
    int time_based;              /* MPPI time_based_weights */
    int cov_mode;                /* MJB_COV_* (DMD update_cov) */
    double* total;               /* (T,K) out: trajectory cost the weights are computed from */
    double* scratch;             /* >= mjb_softmax_scratch_doubles(K,H,d,cov_mode) */
    double* partials;            /* out: mjb_softmax_partial_doubles(H,d,time_based,cov_mode) */
    /* MPPIQ (mjmpc/control/mppiq.py:76-126): returns == MJB_RETURNS_TD_LAMBDA replaces the discounted
     * cost-to-go by the TD(lambda) return of calculate_returns (mppiq.py:104-126) over the per-step total
     * cost c_t + lam * control_cost_t (mppiq.py:96-97; `lam` carries MPPIQ's beta):
     *     q_hat_t = q_t + td_lam * cost_to_go(c_t + gamma q_{t+1} - q_t, td_weight_seq)_t  (t < H-1),
     *     q_hat_{H-1} = q_{H-1},   q = qvals, or (0, ..., 0, c_{H-1}) when qvals == NULL (mppiq.py:109-111)
     * in the reference's operation order (bit-exact given bit-identical per-step costs). */
    int returns;                 /* MJB_RETURNS_* */
    double td_lam, td_gamma;
    const double* td_weight_seq; /* (H-1,) HOST: cumprod([1, gamma*td_lam, ...]) (mppiq.py:117-119); H == 1: unused */
    const double* qvals;         /* (K,H) device by strides, or NULL */
    long long q_sk, q_st;
} mjb_softmax_args;
long long mjb_softmax_scratch_doubles(int K, int H, int d, int cov_mode);
int mjb_softmax_partial_doubles(int H, int d, int time_based, int cov_mode);
int mjb_softmax_partials(const mjb_softmax_args* a, void* stream);

typedef struct {
    int H, d, n_shards;
    long long K_global;
    const double* partials;      /* (n_shards, P) device, rank order */
    double lam, step_size;
    int time_based, cov_mode;
    double* mean;                /* (H,d) in/out: (1-step)*mean + step*sum_k w a; NULL = statistics only */
    double* cov;                 /* (d,d) in/out when cov_mode != NONE, else may be NULL */
    double* stats;               /* out (2+2T): [0] = -lam*logsumexp(-total/lam, b=1/K) (the _calc_val value,
                                    mppi.py:113-131, gaussian_dmd.py:126-139), [1] = global min m[0],
                                    [2..2+T) = normaliser S[t] w.r.t. the global min, [2+T..2+2T) = global min m[t] */
} mjb_combine_args;
int mjb_softmax_combine(const mjb_combine_args* a, void* stream);
/* Fused multi-GPU variant of the combine: pushes `local_partial` (P doubles) into every peer's symmetric
 * buffer over NVLink peer memory, waits for all peers' sequence flags, then combines in rank order (one
 * kernel, no NCCL call).  peer_bufs_dev: DEVICE array of n_shards pointers to the peers' buffers, each
 * 2*n_shards*P doubles followed by 2*n_shards 64-bit flags (zero-initialised); a->partials is ignored;
 * seq = 1, 2, 3, ... must advance identically on every rank. */
int mjb_softmax_exchange_combine(const mjb_combine_args* a, const double* local_partial, void* const* peer_bufs_dev,
                                 int rank, unsigned long long seq, void* stream);
/* The whole update of the softmax controllers after a rollout -- phase 1, the peer exchange (n_shards > 1),
 * phase 2, next action, hot-start shift -- in two launches (three with the control cost): the last block of the
 * weighted reduction to finish runs everything that follows it.  Replaces, bit for bit, the sequence
 * mjb_softmax_partials -> mjb_softmax_combine | mjb_softmax_exchange_combine -> (action_out <- mean[0]) ->
 * mjb_shift_mean -> mjb_cov_add_diag, i.e. MPPI._update_distribution (mppi.py:69-97) / DMDMPC._update_distribution
 * (gaussian_dmd.py:65-91) + _get_next_action (olgaussian_mpc.py:69-78) + _shift (:116-129, gaussian_dmd.py:106-113).
 * a->scratch[0] holds the "blocks done" counter: zero the scratch buffer once after allocating it.  action_out
 * (d,) device or NULL; shift: 0 / 1; base_action: MJB_BASE_NULL or MJB_BASE_REPEAT; cov_shift_beta != 0:
 * cov += beta I after the shift. */
int mjb_softmax_update_fused(const mjb_softmax_args* a, const mjb_combine_args* c, void* const* peer_bufs_dev, int rank,
                             unsigned long long seq, double* action_out, int shift, int base_action,
                             double cov_shift_beta, void* stream);
/* normalised weights w_k = exp(-(total_k - m)/lam)/S for one t-row of `total`; m,S from stats. */
int mjb_softmax_weights(const double* total, int K, const double* stats, int t, double lam, double* w_out,
                        void* stream);

/* Batched MPPI update for many small independent controllers (sweeps, dynamics-randomisation batches:
 * BASELINE config 5): controller c owns particles [c*K, (c+1)*K) and mean row c; one thread block per
 * controller does cost-to-go, softmax and the weighted mean of mppi.py:69-111.  No collective. */
typedef struct {
    int n_ctrl, K, H, d;         /* K particles per controller (K*8 bytes of shared memory, K <= 4096) */
    const double* costs; long long costs_sk, costs_st;      /* (n_ctrl*K, H) */
    const double* actions; long long act_sk, act_st, act_sj; /* (n_ctrl*K, H, d) */
    double* mean;                /* (n_ctrl, H, d) in/out */
    const double* cov;           /* (d,d), read when control_cost */
    const double* gamma_seq;     /* (H,) HOST */
    double lam, step_size;
    int control_cost;
    double* value;               /* out (n_ctrl,): -lam*logsumexp(-total/lam, b=1/K), or NULL */
} mjb_mppi_batched_args;
int mjb_mppi_update_batched(const mjb_mppi_batched_args* a, void* stream);

/* Batched PFMPC update for many small independent particle filters (BASELINE config 5): controller c owns
 * particles [c*K, (c+1)*K); one thread block per controller does PFMPC._exp_util
 * (particle_filter_controller.py:104-113), the systematic resampler (:159-174, sequential cumulative sum
 * in the reference's order; idx[m] = first i with c_i >= r_c + m/K) and mean_action = mean of the
 * resampled set (:102).  No collective. */
typedef struct {
    int n_ctrl, K, H, d;         /* K particles per controller (K <= 4096) */
    const double* costs; long long costs_sk, costs_st;       /* (n_ctrl*K, H) */
    const double* samples; long long s_sk, s_st, s_sj;        /* (n_ctrl*K, H, d) current particle sets */
    const double* gamma_seq;     /* (H,) HOST */
    double lam;
    const double* r;             /* (n_ctrl,) device: r_c = random.uniform(0, 1/K) of each controller */
    double* weights;             /* out (n_ctrl*K,) normalised weights, or NULL */
    long long* idx;              /* out (n_ctrl*K,) resampled indices, local to the controller, or NULL */
    double* out; long long o_sk, o_st, o_sj;                  /* out (n_ctrl*K, H, d): out[c,m] = samples[c, idx[c,m]] */
    double* mean;                /* out (n_ctrl, H, d) */
} mjb_pf_batched_args;
int mjb_pf_update_batched(const mjb_pf_batched_args* a, void* stream);
/* out[k] = x[k] - mean[k / K] for n_ctrl stacked particle sets (PFMPC.generate_rollouts, :87) */
int mjb_particle_sub_mean_batched(const double* x, long long sk, long long st, long long sj, const double* mean,
                                  int n_ctrl, int K, int H, int d, double* out, long long out_sk, long long out_st,
                                  long long out_sj, void* stream);

/* ---- K5 elite selection + moments: CEM._update_distribution (mjmpc/control/cem.py:65-86) and
 * RandomShooting._update_distribution (mjmpc/control/random_shooting.py:52-62).
 * mjb_select_elites: the num_elite smallest keys of ctg0 (K_global,), ties broken by LOWER index
 * (np.argsort's tie order is unspecified; this is the contract SURVEY 7-H4 fixes).  Writes
 * flags (K_global,) uint8, and ids (num_elite,) int64 ascending by index (NULL ok). */
int mjb_select_elites(const double* ctg0, long long K_global, long long num_elite, unsigned char* flags,
                      long long* ids, void* scratch /* >= 4096 bytes */, void* stream);
/* first index of the minimum of ctg0 (np.argmin): out_index (1,) int64 device */
int mjb_argmin(const double* ctg0, long long K, long long* out_index, double* out_value, void* stream);
/* Elite moments of this shard: partial = [ n | sum_a (H,d) | sum_delta (d) ]  (pass 1)
 * and [ sum (delta-mu)(delta-mu)' lower-triangular d(d+1)/2 ] (pass 2, mu = pooled mean of delta
 * over all shards, device pointer (d,)). */
typedef struct {
    int K, H, d;
    const unsigned char* flags;  /* (K,) elite flags of THIS shard's particles */
    const double* actions; long long act_sk, act_st, act_sj;
    const double* mean;          /* (H,d) */
    const double* mu;            /* (d,) pooled mean of elite deltas (pass 2 only) */
    double* scratch;             /* >= mjb_elite_scratch_doubles(K,H,d) */
    double* partial;             /* out */
} mjb_elite_args;
long long mjb_elite_scratch_doubles(int K, int H, int d);
int mjb_elite_moments1(const mjb_elite_args* a, void* stream);
int mjb_elite_moments2(const mjb_elite_args* a, void* stream);
typedef struct {
    int H, d, n_shards, full_cov;          /* full_cov: 1 = np.cov (ddof 1), 0 = diag(np.var) (ddof 0) */
    const double* partial1;      /* (n_shards, 1 + H*d + d) */
    const double* partial2;      /* (n_shards, d(d+1)/2) or NULL when only the pooled mean is wanted */
    double step_size;
    double* mu;                  /* out (d,): pooled mean of elite deltas */
    double* mean;                /* (H,d) in/out (only touched when partial2 != NULL) */
    double* cov;                 /* (d,d) in/out (only touched when partial2 != NULL) */
} mjb_elite_combine_args;
int mjb_elite_combine(const mjb_elite_combine_args* a, void* stream);
/* mean <- (1-step)*mean + step*actions[best] (random_shooting.py:61-62); best_index device (1,) */
int mjb_blend_best(const double* actions, long long sk, long long st, long long sj, const long long* best_index,
                   long long k_offset, int K, int H, int d, double step_size, double* mean, void* stream);

/* Batched update of many small independent controller instances (one thread block each, no cross-instance
 * reduction), the RandomShooting / CEM counterpart of mjb_mppi_update_batched: instance c owns particles
 * [c*K, (c+1)*K), mean row c and (CEM) covariance c.
 *   mode MJB_INST_RS:        mean <- (1-step) mean + step actions[argmin ctg0]      (random_shooting.py:52-62)
 *   mode MJB_INST_CEM_DIAG / MJB_INST_CEM_FULL:  elite set = num_elite smallest ctg0 (ties -> lower index);
 *        cov <- (1-step) cov + step (diag(np.var) | np.cov) of the elite deltas pooled over (elite, t), deltas taken
 *        from the mean BEFORE its update; mean <- (1-step) mean + step mean(elite actions)      (cem.py:65-86)
 * ids out: (n_ctrl, num_elite) elite indices within the instance, ascending (RS: (n_ctrl,) the argmin); value out:
 * (n_ctrl,) mean cost-to-go (cem.py:107-113, random_shooting.py:65-69); either may be NULL.  apply == 0: values /
 * ids only, mean and cov untouched.  K <= 4096. */
#define MJB_INST_RS 0
#define MJB_INST_CEM_DIAG 1
#define MJB_INST_CEM_FULL 2
typedef struct {
    int n_ctrl, K, H, d, mode, apply;
    long long num_elite;
    const double* costs; long long costs_sk, costs_st;       /* (n_ctrl*K, H) */
    const double* actions; long long act_sk, act_st, act_sj; /* (n_ctrl*K, H, d) */
    double* mean;                /* (n_ctrl, H, d) in/out */
    double* cov;                 /* (n_ctrl, d, d) in/out (CEM modes) */
    const double* gamma_seq;     /* (H,) HOST */
    double step_size;
    long long* ids;
    double* value;
} mjb_instances_args;
int mjb_instances_update_batched(const mjb_instances_args* a, void* stream);
/* cov[c] += beta * diag(v) for n stacked (d,d) covariances (CEM._shift, cem.py:89-95, batched instances) */
int mjb_cov_add_diag_batched(double* cov, int n, int d, double beta, const double* v, void* stream);

/* ---- K7 systematic resampling: PFMPC._resampling (particle_filter_controller.py:159-174).
 * idx[m] = first i with c_i >= r + m/M (clamped to M-1), c_i the SEQUENTIAL FP64 prefix sum of the reference's
 * accumulation order, so indices are bit-exact.  M >= 4096: a parallel prefix sum whose indices are certified
 * against the rounding distance 4 M 2^-53 sum(w) to the bin edges; the sequential scan only re-runs when a sample
 * cannot be certified.  r = random.uniform(0, 1/M) from the host. */
int mjb_resample_indices(const double* weights, long long M, double r, double* cumsum_scratch /* (M + 2,) */,
                         long long* idx_out, void* stream);
/* out[m] = in[idx[m]] for (K,H,d) tensors by strides */
int mjb_gather_particles(const double* in, long long in_sk, long long in_st, long long in_sj, const long long* idx,
                         int K, int H, int d, double* out, long long out_sk, long long out_st, long long out_sj,
                         void* stream);
/* mean over particles: out (H,d) = mean_k x[k] (PFMPC mean_action, particle_filter_controller.py:102,124) */
int mjb_particle_mean(const double* x, long long sk, long long st, long long sj, int K, int H, int d,
                      double* scratch, double* out, void* stream);

/* out = x - mean[None]: the deviation PFMPC.generate_rollouts hands to rollout_fn
 * (particle_filter_controller.py:87) */
int mjb_particle_sub_mean(const double* x, long long sk, long long st, long long sj, const double* mean, int K, int H,
                          int d, double* out, long long out_sk, long long out_st, long long out_sj, void* stream);

/* ---- K8 shift: OLGaussianMPC._shift (mjmpc/control/olgaussian_mpc.py:116-129), CEM._shift
 * (cem.py:89-95), DMDMPC._shift (gaussian_dmd.py:106-113), PFMPC._shift
 * (particle_filter_controller.py:127-150). */
#define MJB_BASE_NULL 0
#define MJB_BASE_REPEAT 1
#define MJB_BASE_RANDOM 2
/* mean[:-1] = mean[1:]; last row per base_action; `random_row` (d,) device, used for MJB_BASE_RANDOM */
int mjb_shift_mean(double* mean, int H, int d, int base_action, const double* random_row, void* stream);
/* the same for n stacked (H,d) means; random_rows (n,d) */
int mjb_shift_mean_batched(double* mean, int n, int H, int d, int base_action, const double* random_rows, void* stream);
/* cov += beta * diag(v)   (v (d,) device, or NULL for the identity) */
int mjb_cov_add_diag(double* cov, int d, double beta, const double* v, void* stream);
/* samples[:, :-1] = samples[:, 1:]; samples += delta; last column per base_action */
int mjb_pf_shift(double* samples, long long sk, long long st, long long sj, const double* delta, long long dk,
                 long long dt, long long dj, int K, int H, int d, int base_action, const double* random_row,
                 void* stream);

/* ---- one whole MPC step of the softmax controllers (MPPI, DMD-MPC) in ONE host call: the loop of
 * Controller.optimize (mjmpc/control/controller.py:207-257) = n_iters x { generate_rollouts
 * (olgaussian_mpc.py:95-114), _update_distribution (mppi.py:69-97, gaussian_dmd.py:65-91) }, next action
 * = mean[0] (olgaussian_mpc.py:69-78), _shift (:116-129; gaussian_dmd.py:106-113).  Launches exactly the
 * kernels of mjb_generate_noise, mjb_rollout_reacher, mjb_softmax_partials, mjb_softmax_combine (or
 * mjb_softmax_exchange_combine), mjb_shift_mean and mjb_cov_add_diag, in that order on `stream`: results are
 * bit-identical to calling them one by one; the host time between launches goes away (it bounds the sharded
 * multi-GPU step, which cannot replay a CUDA graph).  The argument blocks are the ones of the separate entry
 * points and are read during the call only. */
typedef struct {
    int n_iters;
    const mjb_noise_args* noise;       /* K2, re-run every iteration (same (seed, step): same samples; the zero
                                          control sequence follows the updated mean); NULL = no noise kernel (the
                                          rollout reads a tensor an earlier call's noise_next has filled) */
    const mjb_noise_args* noise_next;  /* optional: the NEXT step's noise (it depends on (seed, step) only) drawn on a
                                          side stream while this step rolls out -- forked after the work already on
                                          `stream`, joined back before the call returns its stream to the caller; must
                                          write a different tensor than the one this step's rollout reads.  NULL = off */
    const mjb_model* model;
    const mjb_rollout_args* rollout;
    const mjb_softmax_args* softmax;
    const mjb_combine_args* combine;   /* must apply the update (mean != NULL); n_shards == 1: partials ==
                                          softmax->partials */
    void* const* peer_bufs_dev;        /* n_shards > 1: buffers of mjb_softmax_exchange_combine */
    int rank;
    unsigned long long seq;            /* exchange sequence number of iteration 0; iteration i uses seq + i */
    double* action_out;                /* (d,) device <- mean[0] after the last update, before the shift; or NULL */
    int shift;                         /* hotstart: shift the mean sequence afterwards */
    int base_action;                   /* MJB_BASE_NULL or MJB_BASE_REPEAT */
    double cov_shift_beta;             /* != 0: cov += beta * I after the shift (DMD-MPC with update_cov) */
} mjb_mpc_step_args;
int mjb_softmax_mpc_step(const mjb_mpc_step_args* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif
