// TEST HARNESS (CPU): compiles the product's device math (chain_dynamics.cuh) for the host so the
// link-frame formulation can be compared with the oracle without a GPU.  Not a product path.
#include <cstddef>
#define MJB_HOST_STATS 1
#include "../../mjmpc_b200/csrc/chain_dynamics.cuh"

using namespace mjb;
struct PtrParams { const double* p; double operator[](int i) const { return p[i]; } };

// policy_w != NULL: closed-loop linear policy, the same steps as the kernel's `closed` branch (rollout_reacher.cu)
template <class T>
static void run(const double* P, const double* qpos, const double* qvel, const double* target, int K, int H,
                const double* mean, const double* noise, double* costs, double* qv, int* iters,
                const double* policy_w = nullptr, double* actions = nullptr) {
    PtrParams prm{P};
    const int fs = (int)P[CS_FRAME_SKIP];
    const V3 hand_local{P[CS_HAND], P[CS_HAND + 1], P[CS_HAND + 2]};
    for (int k = 0; k < K; k++) {
        double q[7], v[7], sn[7], cs[7];
        HostScratch sc;
        for (int j = 0; j < 7; j++) { q[j] = qpos[j]; v[j] = qvel[j]; }
        V3 hand_prev{0, 0, 0};
        if (policy_w) {
            for (int j = 0; j < 7; j++) sincos_joint(q[j], sn[j], cs[j]);
            hand_prev = chain_point_world<T>(prm, sn, cs, hand_local);
        }
        for (int t = 0; t < H; t++) {
            double ub[7];
            if (policy_w) {
                double ob[20];
                for (int j = 0; j < 7; j++) { ob[j] = q[j]; ob[7 + j] = v[j]; }
                ob[14] = hand_prev.x; ob[15] = hand_prev.y; ob[16] = hand_prev.z;
                ob[17] = hand_prev.x - target[0]; ob[18] = hand_prev.y - target[1]; ob[19] = hand_prev.z - target[2];
                for (int j = 0; j < 7; j++) {
                    double s = policy_w[20 * 7 + j];
                    for (int i = 0; i < 20; i++) s = fma(policy_w[i * 7 + j], ob[i], s);
                    ub[j] = s;
                }
            } else for (int j = 0; j < 7; j++) ub[j] = mean[t * 7 + j];
            for (int j = 0; j < 7; j++) {
                const double x = ub[j] + noise[((size_t)k * H + t) * 7 + j];
                if (actions) actions[((size_t)k * H + t) * 7 + j] = x;
                sc.st(SC_U + j, actuator_torque(prm, j, x));
            }
            V3 hand{0, 0, 0};
            for (int s = 0; s < fs; s++) {
                for (int j = 0; j < 7; j++) sincos_joint(q[j], sn[j], cs[j]);
                if (s == fs - 1) hand = chain_point_world<T>(prm, sn, cs, V3{P[CS_HAND], P[CS_HAND + 1], P[CS_HAND + 2]});
                chain_substep<T>(prm, sc, q, v, sn, cs);
            }
            costs[(size_t)k * H + t] = reach_cost(hand, V3{target[0], target[1], target[2]});
            hand_prev = hand;
            for (int j = 0; j < 7; j++) { qv[((size_t)k * H + t) * 14 + j] = q[j]; qv[((size_t)k * H + t) * 14 + 7 + j] = v[j]; }
        }
    }
    (void)iters;
}

extern "C" void hostcheck_rollout(const double* P167, int dense, const double* qpos, const double* qvel,
                                  const double* target, int K, int H, const double* mean, const double* noise,
                                  double* costs, double* qv) {
    double P[CH_NDEV];
    for (int i = 0; i < CH_NPARAM; i++) P[i] = P167[i];
    mjb_derive_params(P);
    if (dense) run<DenseTraits>(P, qpos, qvel, target, K, H, mean, noise, costs, qv, nullptr);
    else run<SawyerTraits>(P, qpos, qvel, target, K, H, mean, noise, costs, qv, nullptr);
}
extern "C" void hostcheck_rollout_cl(const double* P167, const double* qpos, const double* qvel, const double* target,
                                     int K, int H, const double* policy_w, const double* noise, double* costs,
                                     double* qv, double* actions) {
    double P[CH_NDEV];
    for (int i = 0; i < CH_NPARAM; i++) P[i] = P167[i];
    mjb_derive_params(P);
    run<SawyerTraits>(P, qpos, qvel, target, K, H, nullptr, noise, costs, qv, nullptr, policy_w, actions);
}
extern "C" int hostcheck_fits_sawyer(const double* P167) { return mjb_params_fit_sawyer(P167); }
extern "C" void hostcheck_mass_bias(const double* P167, const double* q, const double* v, double* M, double* bias) {
    double P[CH_NDEV];
    for (int i = 0; i < CH_NPARAM; i++) P[i] = P167[i];
    mjb_derive_params(P);
    PtrParams prm{P};
    double sn[7], cs[7], qd[7], b[7];
    HostScratch sc;
    for (int j = 0; j < 7; j++) { sincos_joint(q[j], sn[j], cs[j]); qd[j] = v[j]; }
    chain_mass_bias<SawyerTraits>(prm, sc, sn, cs, qd, b);
    for (int i = 0; i < 7; i++) { bias[i] = b[i]; for (int j = 0; j <= i; j++) { M[i * 7 + j] = sc.ld(sc_m(i, j)); M[j * 7 + i] = M[i * 7 + j]; } }
}

extern "C" void hostcheck_sincos(double x, double* s, double* c) { mjb::sincos_joint(x, *s, *c); }

// record the factor/solve passes of every following substep into buf (call order: particle, step, substep)
extern "C" void hostcheck_record_trips(int* buf) { mjb::g_trips = buf; mjb::g_ntrips = 0; }
extern "C" void hostcheck_stats(long long* out, int reset) {
    for (int i = 0; i < 6; i++) { out[i] = mjb::g_stats[i]; if (reset) mjb::g_stats[i] = 0; }
}

