// K9: batched SimplePendulum-v0 rollout with the step cost fused in, one thread per particle.
// Replaces PendulumEnv.step (mjmpc/envs/basic/pendulum.py:33-50) under GymEnvWrapper.rollout
// (mjmpc/envs/gym_env_wrapper.py:125-153).  Arithmetic follows the reference expression by
// expression (no FMA contraction); sin() is the only operation that may differ from numpy by an ulp.
#include "common.h"

namespace mjb {
__global__ void rollout_pendulum_kernel(mjb_pendulum_args a) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.K) return;
    const int ctrl = (int)(k / a.particles_per_ctrl);
    const double g = 10.0, m = 1.0, l = 1.0, dt = .05, max_speed = 8.0, max_torque = 2.0;
    const double PI = 3.141592653589793;
    double th = a.state[ctrl * 2], thdot = a.state[ctrl * 2 + 1];
    const double* mean = a.mean + (long long)ctrl * a.H;
    const double c1 = -3 * g / (2 * l), c2 = 3. / (m * (l * l));
    for (int t = 0; t < a.H; t++) {
        double act = mean[t];
        if (a.noise) act = __dadd_rn(act, a.noise[k * a.noise_sk + t * a.noise_st]);
        if (a.actions) a.actions[k * a.act_sk + t * a.act_st] = act;
        const double u = fmin(fmax(act, -max_torque), max_torque);
        // angle_normalize: ((x + pi) % (2 pi)) - pi with python's sign-of-divisor modulo
        double x = fmod(__dadd_rn(th, PI), 2 * PI);
        if (x < 0.0) x = __dadd_rn(x, 2 * PI);
        x = __dadd_rn(x, -PI);
        const double cost = __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(.1, __dmul_rn(thdot, thdot))),
                                      __dmul_rn(.001, __dmul_rn(u, u)));
        a.costs[k * a.costs_sk + t * a.costs_st] = cost;
        const double acc = __dadd_rn(__dmul_rn(c1, sin(__dadd_rn(th, PI))), __dmul_rn(c2, u));
        const double nthdot = __dadd_rn(thdot, __dmul_rn(acc, dt));
        th = __dadd_rn(th, __dmul_rn(nthdot, dt));
        thdot = fmin(fmax(nthdot, -max_speed), max_speed);
        if (a.states_out) { a.states_out[(k * a.H + t) * 2] = th; a.states_out[(k * a.H + t) * 2 + 1] = thdot; }
    }
}
}  // namespace mjb

#ifndef MJB_HOST_EMU
extern "C" int mjb_rollout_pendulum(const mjb_pendulum_args* a, void* stream) {
    MJB_REQUIRE(a && a->state && a->mean && a->costs, "mjb_rollout_pendulum: null pointer");
    MJB_REQUIRE(a->K >= 1 && a->H >= 1, "mjb_rollout_pendulum: K and H must be positive");
    MJB_REQUIRE(a->particles_per_ctrl >= 1 && a->K % a->particles_per_ctrl == 0,
                "Number of particles must be divisible by number of controllers");
    mjb::rollout_pendulum_kernel<<<(a->K + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*a);
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}
#endif  // MJB_HOST_EMU
