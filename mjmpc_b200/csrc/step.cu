// One whole MPC step of the softmax controllers (MPPI, DMD-MPC) behind ONE host call.
// Replaces the host loop of Controller.optimize (mjmpc/control/controller.py:207-257):
//   n_iters x { generate_rollouts (olgaussian_mpc.py:95-114) -> _update_distribution (mppi.py:69-97,
//   gaussian_dmd.py:65-91) } -> next action = mean[0] (olgaussian_mpc.py:69-78) -> _shift (:116-129,
//   gaussian_dmd.py:106-113).
// It launches exactly the kernels the separate entry points launch, in the same order on the same stream, so
// the results are bit-identical to calling them one by one; what goes away is the host time between the
// launches (interpreter, ctypes marshalling, tensor bookkeeping: ~0.3 ms per step from Python against
// ~0.04 ms here), which is what bounds a step that cannot replay a CUDA graph -- the sharded multi-GPU step,
// whose exchange kernel waits on peers -- once a GPU's share of the particles is small.
#include <mutex>
#include "common.h"

extern "C" int mjb_softmax_mpc_step(const mjb_mpc_step_args* a, void* stream) {
    MJB_REQUIRE(a && a->model && a->rollout && a->softmax && a->combine, "mjb_softmax_mpc_step: null argument block");
    MJB_REQUIRE(a->n_iters >= 1, "mjb_softmax_mpc_step: n_iters must be >= 1");
    MJB_REQUIRE(!a->noise_next || (!a->noise_next->zero_last && a->combine->cov_mode == MJB_COV_NONE),
                "mjb_softmax_mpc_step: the next step's noise depends on this step's result (zero control sequence / adapted covariance)");
    MJB_REQUIRE(a->base_action == MJB_BASE_NULL || a->base_action == MJB_BASE_REPEAT || !a->shift,
                "mjb_softmax_mpc_step: base_action 'random' needs a host-drawn row; use the separate entry points");
    const mjb_combine_args* c = a->combine;
    MJB_REQUIRE(c->mean, "mjb_softmax_mpc_step: the combine must apply the update (mean is NULL)");
    MJB_REQUIRE(c->n_shards >= 1 && (c->n_shards == 1 || a->peer_bufs_dev),
                "mjb_softmax_mpc_step: %d shards need the peer-memory exchange buffers", c->n_shards);
    MJB_REQUIRE(c->n_shards > 1 || c->partials == a->softmax->partials,
                "mjb_softmax_mpc_step: with one shard the combine must read the partial vector phase 1 writes");
    MJB_REQUIRE(c->H == a->rollout->H && a->softmax->H == c->H && a->softmax->K == a->rollout->K && a->softmax->d == c->d,
                "mjb_softmax_mpc_step: rollout / softmax / combine shapes disagree");
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    bool forked = false;
    // per-device side stream and fork / join events, created once under a lock (controllers of several threads --
    // logical ranks in tests -- may step concurrently).  Steps that share a device share them: two controllers
    // forking at the same time only add a false dependency between their side-stream launches.
    static std::mutex side_mutex;
    static cudaStream_t side[64] = {nullptr};
    static cudaEvent_t ev_fork[64] = {nullptr}, ev_join[64] = {nullptr};
    const int dev = a->model->device & 63;
    if (a->noise_next) {
        // The next step's noise needs neither this step's state nor its mean: draw it on a side stream while the
        // FP64-bound rollout leaves the integer / SFU pipes (and, in its second wave, a quarter of the registers)
        // idle.  The fork point is everything already queued on `s` (the previous step's rollout was the last
        // reader of the tensor being overwritten); the kernel itself is launched right AFTER this step's rollout
        // and on a lowest-priority stream, so that the rollout's blocks take the SMs first and the noise blocks
        // fill what they leave; joined before returning.
        MJB_REQUIRE(a->noise_next->out != a->rollout->noise, "mjb_softmax_mpc_step: noise_next would overwrite the tensor this step reads");
        std::lock_guard<std::mutex> lock(side_mutex);
        if (!side[dev]) {
            MJB_CUDA(cudaSetDevice(a->model->device));
            int lo = 0, hi = 0;
            MJB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));       // lo = least priority
            MJB_CUDA(cudaStreamCreateWithPriority(&side[dev], cudaStreamNonBlocking, lo));
            MJB_CUDA(cudaEventCreateWithFlags(&ev_fork[dev], cudaEventDisableTiming));
            MJB_CUDA(cudaEventCreateWithFlags(&ev_join[dev], cudaEventDisableTiming));
        }
        MJB_CUDA(cudaEventRecord(ev_fork[dev], s));
        MJB_CUDA(cudaStreamWaitEvent(side[dev], ev_fork[dev], 0));
    }
    // an error after the fork must not leave side-stream work unjoined
    auto fail = [&](int code) {
        if (forked) cudaStreamWaitEvent(s, ev_join[dev], 0);
        return code;
    };
    for (int it = 0; it < a->n_iters; it++) {
        if (a->noise && (rc = mjb_generate_noise(a->noise, stream)) != MJB_OK) return fail(rc);
        if ((rc = mjb_rollout_reacher(a->model, a->rollout, stream)) != MJB_OK) return fail(rc);
        if (a->noise_next && it == 0) {
            std::lock_guard<std::mutex> lock(side_mutex);
            rc = mjb_generate_noise(a->noise_next, (void*)side[dev]);
            // joined even when the launch failed: the side stream still waits on this step's fork event
            MJB_CUDA(cudaEventRecord(ev_join[dev], side[dev]));
            forked = true;
            if (rc != MJB_OK) return fail(rc);
        }
        // the whole update tail rides in the last block of the weighted reduction; after the last iteration it
        // also hands out the next action and shifts the mean sequence (4 launches per MPPI iteration in all)
        const bool final_it = it == a->n_iters - 1;
        rc = mjb_softmax_update_fused(a->softmax, c, a->peer_bufs_dev, a->rank, a->seq + (unsigned long long)it,
                                      final_it ? a->action_out : nullptr, final_it ? a->shift : 0, a->base_action,
                                      final_it && a->shift ? a->cov_shift_beta : 0.0, stream);
        if (rc != MJB_OK) return fail(rc);
    }
    if (forked) MJB_CUDA(cudaStreamWaitEvent(s, ev_join[dev], 0));
    return MJB_OK;
}
