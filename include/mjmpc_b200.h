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
    const double* mean;          /* (n_ctrl, H, 7) row-major */
    const double* noise;         /* (K, H, 7) by strides, or NULL for the mean sequence only */
    long long noise_sk, noise_st, noise_sj;
    double* costs;               /* (K, H) = -reward */
    long long costs_sk, costs_st;
    double* actions;             /* (K, H, 7) unclipped mean+noise (gym_env_wrapper.py:151), or NULL */
    long long act_sk, act_st, act_sj;
    double* qv_traj;             /* (K, H, 14) row-major qpos,qvel after each env step, or NULL */
    double* next_obs;            /* (K, H, MJB_OBS_DIM) row-major, or NULL */
    int* ncon;                   /* (K,) substeps with >=1 active constraint row, or NULL */
} mjb_rollout_args;
int mjb_rollout_reacher(const mjb_model* m, const mjb_rollout_args* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif
