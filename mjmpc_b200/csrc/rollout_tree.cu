// K11: batched rollout of a runtime-parameterised kinematic TREE of hinge / slide joints with the step cost
// fused in, one thread per particle.  Replaces MuJoCo's mj_step under
// gym.envs.mujoco.MujocoEnv.do_simulation for the reference's forward-progress environments
// (mjmpc/envs/basic/swimmer.py:10-19, half_cheetah.py:10-19) inside GymEnvWrapper.rollout
// (mjmpc/envs/gym_env_wrapper.py:123-153).
//
// Formulation (deliberately not the oracle's, which keeps MuJoCo's world-frame quantities about the tree's centre
// of mass): ONE LINK PER DOF, everything in link coordinates at the joint anchor.
//   pass 1 (root -> leaves)  link transform from q; velocity and bias acceleration (base acceleration = -gravity);
//                            link wrench f = I a + v x* I v - fluid wrench (mj_passive's inertia-box model:
//                            viscous + quadratic drag in the body's inertial frame)
//   pass 2 (leaves -> root)  bias force tau_i = S_i . f_i, wrench handed to the parent; composite inertias;
//                            column i of M from the composite's momentum under S_i walked down to the root
//   constraints              soft joint limits: rows +-e_j with MuJoCo's impedance / reference acceleration; the convex
//                            problem  min 1/2 (a-a0)'M(a-a0) + sum 1/2 D min(0, J a - aref)^2  is solved exactly:
//                            Newton on the active set with an exact piecewise-linear line search
//                            contacts (planar instantiation): MuJoCo's plane-capsule / capsule-capsule detection on the
//                            link frames, condim-3 pyramidal friction rows, solved with the limit rows
//   mj_Euler                 implicit in joint damping: (M + h B) qacc = f + J'lambda
// The model block (tree_model.h) is staged in shared memory once per block: link parameters are indexed with
// run-time link numbers and every lane reads the same address (broadcast).
// Instantiations: <7, serial chain> with every loop unrolled and the per-link state in registers (the reference's
// swimmer is a serial chain of 7 dofs); <run-time nv <= 12, any tree> with the per-link state in local memory.
#include "common.h"
#include "tree_model.h"
#include <stdlib.h>

struct mjb_tree_model {
    int device, nv, nu, serial, planar, n_inst;
    double* d_lk;   // n_inst x nv x LK_STRIDE (per-worker models of randomize_dynamics share topology and globals)
    int* d_li;      // nv x LI_STRIDE
    double* d_g;    // TG_STRIDE
    double* d_pk;   // planar mechanisms: n_inst x (nv x PK_STRIDE, then the in-plane gravity (2))
    int* d_anc;     // planar mechanisms: ancestor bit masks (nv)
    int ncand;      // planar mechanisms with contacts: candidate pairs (rollout_tree_planar.cuh)
    int* d_cti;     // ncand x CTI_STRIDE
    double* d_ctd;  // ncand x CT_STRIDE
};

namespace mjb {
namespace tree {

#if defined(__CUDACC__)
#define TR_HD __host__ __device__ __forceinline__
#else
#define TR_HD inline
#endif

#define TR_BIG 1e300

TR_HD void cross(double* r, const double* a, const double* b) {
    const double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    r[0] = x; r[1] = y; r[2] = z;
}
TR_HD double dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
TR_HD void mv(double* r, const double* M, const double* v) {          // r = M v   (row-major 3x3)
    const double x = M[0] * v[0] + M[1] * v[1] + M[2] * v[2], y = M[3] * v[0] + M[4] * v[1] + M[5] * v[2];
    const double z = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
    r[0] = x; r[1] = y; r[2] = z;
}
TR_HD void mtv(double* r, const double* M, const double* v) {         // r = M' v
    const double x = M[0] * v[0] + M[3] * v[1] + M[6] * v[2], y = M[1] * v[0] + M[4] * v[1] + M[7] * v[2];
    const double z = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
    r[0] = x; r[1] = y; r[2] = z;
}
TR_HD void symv(double* r, const double* S, const double* v) {        // S = xx yy zz xy xz yz
    const double x = S[0] * v[0] + S[3] * v[1] + S[4] * v[2], y = S[3] * v[0] + S[1] * v[1] + S[5] * v[2];
    const double z = S[4] * v[0] + S[5] * v[1] + S[2] * v[2];
    r[0] = x; r[1] = y; r[2] = z;
}

// sin / cos of a joint angle: two-term Cody-Waite reduction by pi/2 and the fdlibm kernel polynomials (~1 ulp for
// |x| < 1e5 rad; the same routine as the reacher kernel's).  Beyond that an out-of-line call of the library routine:
// its huge-argument path is ~100 instructions per call site and would sit 5-9 times in the unrolled loops.
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#endif
void sincos_far(double x, double* s, double* c) { sincos(x, s, c); }
TR_HD void sincos_lean(double x, double* s, double* c) {
    if (fabs(x) > 1e5) { sincos_far(x, s, c); return; }
    const double kd = rint(x * 6.36619772367581382433e-01);
    double r = fma(kd, -1.57079632673412561417e+00, x);
    r = fma(kd, -6.07710050650619224932e-11, r);
    const double z = r * r;
    const double ps = fma(z, fma(z, fma(z, fma(z, fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08),
                                                2.75573137070700676789e-06), -1.98412698298579493134e-04),
                                 8.33333333332248946124e-03), -1.66666666666666324348e-01);
    const double pc = fma(z, fma(z, fma(z, fma(z, fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09),
                                                -2.75573143513906633035e-07), 2.48015872894767294178e-05),
                                 -1.38888888888741095749e-03), 4.16666666666666019037e-02);
    const double sr = fma(z * r, ps, r);
    const double cr = fma(z * z, pc, fma(z, -0.5, 1.0));
    const int k = (int)kd;
    const double a = (k & 1) ? cr : sr, b = (k & 1) ? sr : cr;
    *s = (k & 2) ? -a : a;
    *c = ((k + 1) & 2) ? -b : b;
}

TR_HD double impedance(const double* si, double dist) {
    if (si[0] == si[1] || si[2] <= 1e-15) return 0.5 * (si[0] + si[1]);
    const double x = fabs(dist / si[2]);
    if (x >= 1.0) return si[1];
    if (x <= 0.0) return si[0];
    double y;
    if (si[4] == 1.0) y = x;
    else if (si[4] == 2.0) y = x <= si[3] ? x * x / si[3] : 1.0 - (1.0 - x) * (1.0 - x) / (1.0 - si[3]);
    else if (x <= si[3]) y = pow(x, si[4]) / pow(si[3], si[4] - 1.0);
    else y = 1.0 - pow(1.0 - x, si[4]) / pow(1.0 - si[3], si[4] - 1.0);
    return si[0] + y * (si[1] - si[0]);
}

#if defined(__CUDACC__)
__host__ __device__ __noinline__
#endif
double impedance_call(const double* si, double dist) { return impedance(si, dist); }     // one out-of-line copy (pow paths)

// In-place Cholesky solve of the dense SPD system A x = b (lower triangle of A used and overwritten).
template <int N> TR_HD void chol_solve(int n, double (*A)[N], double* b) {
    for (int j = 0; j < n; j++) {
        double s = A[j][j];
        for (int k = 0; k < j; k++) s -= A[j][k] * A[j][k];
        const double d = sqrt(s), inv = 1.0 / d;
        A[j][j] = inv;                                      // keeps 1 / L_jj
        for (int i = j + 1; i < n; i++) {
            double t = A[i][j];
            for (int k = 0; k < j; k++) t -= A[i][k] * A[j][k];
            A[i][j] = t * inv;
        }
    }
    for (int i = 0; i < n; i++) {
        double t = b[i];
        for (int k = 0; k < i; k++) t -= A[i][k] * b[k];
        b[i] = t * A[i][i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double t = b[i];
        for (int k = i + 1; k < n; k++) t -= A[k][i] * b[k];
        b[i] = t * A[i][i];
    }
}

// Soft joint limits, the constrained solve and mj_Euler for given M and f = qfrc_smooth (damping included); q, v
// advanced in place.  The general instantiation inlines it; the planar one calls the out-of-line copy below, and only
// for the rare active set that its own in-register iteration does not settle.  Returns the number of limit rows.
template <int N>
TR_HD int limits_solve_integrate_inl(int nv, const double* lk, const int* li, double h, double (*M)[N], double* f, bool damped,
                                     double* q, double* v) {
    // ---- joint-limit rows
    int nr = 0, rdof[N];
    double rs[N], aref[N], D[N];
    for (int i = 0; i < N; i++) {
        if (i >= nv) break;
        if (!li[i * LI_STRIDE + LI_LIMITED]) continue;
        const double* L = lk + i * LK_STRIDE;
        const double dlo = q[i] - L[LK_LO], dhi = L[LK_HI] - q[i];
        if (dlo < 0.0 || dhi < 0.0) {
            const double side = dlo < 0.0 ? 1.0 : -1.0, dist = dlo < 0.0 ? dlo : dhi;
            const double imp = impedance(L + LK_SOLIMP, dist);
            const double Rr = fmax(1e-15, (1.0 - imp) * L[LK_INVW] / imp);
            rdof[nr] = i; rs[nr] = side; D[nr] = 1.0 / Rr;
            aref[nr] = -L[LK_SOLB] * (side * v[i]) - L[LK_SOLK] * imp * dist;
            nr++;
        }
    }

    // ---- one solver loop, one factorisation site.  phase 0: unconstrained acceleration (this is mj_Euler's solve when
    // no row exists); phase 1: Newton on the active set with an exact line search; phase 2: mj_Euler's solve, implicit in
    // joint damping, with the constraint force on the right-hand side.
    double A[N][N], b[N], a0[N], fc[N];
    unsigned act = 0;
    int phase = 0;
    for (int i = 0; i < nv; i++) fc[i] = 0.0;
    for (int iter = 0; iter < 48; iter++) {
        for (int i = 0; i < nv; i++) {
            b[i] = f[i] + (phase == 2 ? fc[i] : 0.0);
            for (int j = 0; j <= i; j++) A[i][j] = M[i][j];
            if (phase == 2 || (phase == 0 && nr == 0)) A[i][i] += h * lk[i * LK_STRIDE + LK_DAMP];
        }
        if (phase == 1)
            for (int rr = 0; rr < nr; rr++)
                if (act >> rr & 1) { A[rdof[rr]][rdof[rr]] += D[rr]; b[rdof[rr]] += rs[rr] * D[rr] * aref[rr]; }
        chol_solve<N>(nv, A, b);
        if (phase == 2 || nr == 0) break;
        bool converged = false;
        if (phase == 0) {
            for (int i = 0; i < nv; i++) a0[i] = b[i];
            for (int rr = 0; rr < nr; rr++) if (rs[rr] * a0[rdof[rr]] - aref[rr] < 0.0) act |= 1u << rr;
            phase = 1;
            converged = act == 0;
        } else {
            unsigned act1 = 0;
            for (int rr = 0; rr < nr; rr++) if (rs[rr] * b[rdof[rr]] - aref[rr] < 0.0) act1 |= 1u << rr;
            if (act1 == act) {
                for (int i = 0; i < nv; i++) a0[i] = b[i];
                converged = true;
            } else {
                // the set changes along the step: exact minimiser of the piecewise quadratic on the ray a0 + t (b - a0)
                double pdir[N], g0 = 0.0, h0 = 0.0;
                for (int i = 0; i < nv; i++) pdir[i] = b[i] - a0[i];
                for (int i = 0; i < nv; i++) {
                    double Mp = 0.0, Ma = -f[i];
                    for (int k = 0; k < nv; k++) { Mp += M[i][k] * pdir[k]; Ma += M[i][k] * a0[k]; }
                    g0 += pdir[i] * Ma; h0 += pdir[i] * Mp;
                }
                double res[N], Jp[N], tcur = 0.0, tstar = 1.0;
                for (int rr = 0; rr < nr; rr++) { res[rr] = rs[rr] * a0[rdof[rr]] - aref[rr]; Jp[rr] = rs[rr] * pdir[rdof[rr]]; }
                for (int seg = 0; seg <= nr; seg++) {
                    double tnext = TR_BIG;
                    for (int rr = 0; rr < nr; rr++)
                        if (Jp[rr] != 0.0) { const double tb = -res[rr] / Jp[rr]; if (tb > tcur && tb < tnext) tnext = tb; }
                    const double tmid = tnext >= TR_BIG ? tcur + 1.0 : 0.5 * (tcur + tnext);
                    double c0 = g0, c1 = h0;
                    for (int rr = 0; rr < nr; rr++)
                        if (res[rr] + tmid * Jp[rr] < 0.0) { c0 += D[rr] * res[rr] * Jp[rr]; c1 += D[rr] * Jp[rr] * Jp[rr]; }
                    const double t = -c0 / c1;
                    if (t <= tnext || tnext >= TR_BIG) { tstar = t < tcur ? tcur : t; break; }
                    tcur = tnext;
                }
                for (int i = 0; i < nv; i++) a0[i] += tstar * pdir[i];
                act = 0;
                for (int rr = 0; rr < nr; rr++) if (rs[rr] * a0[rdof[rr]] - aref[rr] < 0.0) act |= 1u << rr;
            }
        }
        if (converged) {
            for (int rr = 0; rr < nr; rr++) {
                const double sres = rs[rr] * a0[rdof[rr]] - aref[rr];
                if (sres < 0.0) fc[rdof[rr]] += rs[rr] * (-D[rr] * sres);
            }
            if (!damped) { for (int i = 0; i < nv; i++) b[i] = a0[i]; break; }
            phase = 2;
        }
    }
    for (int i = 0; i < N; i++) {
        if (i >= nv) break;
        v[i] += h * b[i];
        q[i] += h * v[i];
    }
    return nr;
}
template <int N>
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#endif
int limits_solve_integrate(int nv, const double* lk, const int* li, double h, double (*M)[N], double* f, bool damped,
                           double* q, double* v) {
    return limits_solve_integrate_inl<N>(nv, lk, li, h, M, f, damped, q, v);
}

}  // namespace tree
}  // namespace mjb
#include "rollout_tree_planar.cuh"
namespace mjb {
namespace tree {

// One mj_step.  NV > 0: compile-time dof count; NV == 0: run-time nv <= MJB_TREE_MAX_LINKS.  SERIAL: link i hangs
// off link i - 1 (the parent's frame and velocity ride in registers instead of per-link arrays).  UNROLL: per-link
// loops unrolled.  u: controls (nu).  q, v advanced in place.  Returns the number of limit rows.
//
// All spatial vectors are world-oriented and refer to the point O = origin of link 0, so handing a wrench or an
// inertia to the parent is a plain addition and an entry of M is one 6-vector dot product.
template <int NV, bool SERIAL, bool UNROLL>
TR_HD int substep(const double* lk, const int* li, const double* g, int nv_rt, double* q, double* v, const double* u) {
    constexpr int N = NV > 0 ? NV : MJB_TREE_MAX_LINKS;
    constexpr int NS = SERIAL ? 1 : N;                      // slots for what a child needs from its parent
    constexpr int UNR = (NV > 0 && UNROLL) ? NV : 1;
    const int nv = NV > 0 ? NV : nv_rt;
    const double h = g[TG_DT], rho = g[TG_RHO], visc = g[TG_VISC];
    double Rw[NS][9], pw[NS][3], V[NS][6], Ab[NS][6];
    double S[N][6], F[N][6], cm[N], ch[N][3], cI[N][6], M[N][N], f[N];
    double O[3] = {0.0, 0.0, 0.0};

    // ---- pass 1: root -> leaves
#pragma unroll(UNR)
    for (int i = 0; i < N; i++) {
        if (i >= nv) break;
        const double* L = lk + i * LK_STRIDE;
        const int* I = li + i * LI_STRIDE;
        const int p = SERIAL ? i - 1 : I[LI_PARENT];
        const int ps = SERIAL ? 0 : (p < 0 ? 0 : p), is = SERIAL ? 0 : i;
        const bool hinge = I[LI_TYPE] == MJB_TREE_HINGE;
        const double* a = L + LK_AXIS;
        // parent frame, velocity, bias acceleration (world: identity at -O, at rest, accelerating against gravity)
        double Rp[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, pp[3] = {-O[0], -O[1], -O[2]};
        double Vp[6] = {0, 0, 0, 0, 0, 0}, Ap[6] = {0, 0, 0, -g[TG_GRAV], -g[TG_GRAV + 1], -g[TG_GRAV + 2]};
        if (p >= 0) {
#pragma unroll
            for (int x = 0; x < 9; x++) Rp[x] = Rw[ps][x];
#pragma unroll
            for (int x = 0; x < 3; x++) pp[x] = pw[ps][x];
#pragma unroll
            for (int x = 0; x < 6; x++) { Vp[x] = V[ps][x]; Ap[x] = Ab[ps][x]; }
        }
        // link frame: R = Rp Rfix Rjoint, origin = pp + Rp (off [+ Rfix a q])
        double Rl[9], ol[3] = {L[LK_OFF], L[LK_OFF + 1], L[LK_OFF + 2]};
        if (hinge) {
            double sn, cs;
            sincos_lean(q[i], &sn, &cs);
            const double t = 1.0 - cs;
            const double J[9] = {cs + t * a[0] * a[0],        t * a[0] * a[1] - sn * a[2], t * a[0] * a[2] + sn * a[1],
                                 t * a[0] * a[1] + sn * a[2], cs + t * a[1] * a[1],        t * a[1] * a[2] - sn * a[0],
                                 t * a[0] * a[2] - sn * a[1], t * a[1] * a[2] + sn * a[0], cs + t * a[2] * a[2]};
            if (I[LI_BODY] & 2) {                               // Rfix = identity
#pragma unroll
                for (int x = 0; x < 9; x++) Rl[x] = J[x];
            } else {
#pragma unroll
                for (int x = 0; x < 3; x++)
#pragma unroll
                    for (int y = 0; y < 3; y++)
                        Rl[3 * x + y] = L[LK_RFIX + 3 * x] * J[y] + L[LK_RFIX + 3 * x + 1] * J[3 + y] + L[LK_RFIX + 3 * x + 2] * J[6 + y];
            }
        } else {
            double d[3];
#pragma unroll
            for (int x = 0; x < 9; x++) Rl[x] = L[LK_RFIX + x];
            mv(d, L + LK_RFIX, a);
#pragma unroll
            for (int x = 0; x < 3; x++) ol[x] += d[x] * q[i];
        }
        double R[9], pos[3], aw[3];
#pragma unroll
        for (int x = 0; x < 3; x++)
#pragma unroll
            for (int y = 0; y < 3; y++) R[3 * x + y] = Rp[3 * x] * Rl[y] + Rp[3 * x + 1] * Rl[3 + y] + Rp[3 * x + 2] * Rl[6 + y];
        mv(pos, Rp, ol);
#pragma unroll
        for (int x = 0; x < 3; x++) pos[x] += pp[x];
        if (i == 0) {                                           // the reference point: everything below is relative to it
#pragma unroll
            for (int x = 0; x < 3; x++) { O[x] = pos[x]; pos[x] = 0.0; }
        }
        mv(aw, R, a);
        // motion axis about O, velocity, bias acceleration
        double Si[6];
        if (hinge) { Si[0] = aw[0]; Si[1] = aw[1]; Si[2] = aw[2]; cross(Si + 3, pos, aw); }
        else { Si[0] = 0.0; Si[1] = 0.0; Si[2] = 0.0; Si[3] = aw[0]; Si[4] = aw[1]; Si[5] = aw[2]; }
        double Sd[6], t1[3], t2[3];
        cross(Sd, Vp, Si);
        cross(t1, Vp, Si + 3);
        cross(t2, Vp + 3, Si);
#pragma unroll
        for (int x = 0; x < 3; x++) Sd[3 + x] = t1[x] + t2[x];
        double Vi[6], Ai[6];
#pragma unroll
        for (int x = 0; x < 6; x++) { Vi[x] = Vp[x] + Si[x] * v[i]; Ai[x] = Ap[x] + Sd[x] * v[i]; S[i][x] = Si[x]; }
#pragma unroll
        for (int x = 0; x < 9; x++) Rw[is][x] = R[x];
#pragma unroll
        for (int x = 0; x < 3; x++) pw[is][x] = pos[x];
#pragma unroll
        for (int x = 0; x < 6; x++) { V[is][x] = Vi[x]; Ab[is][x] = Ai[x]; }
        // link wrench about O and its inertia as the seed of the composite
        if (I[LI_BODY] & 1) {
            const double m = L[LK_MASS];
            const double* Ic = L + LK_IC;
            double c[3], hh[3], Io[6];
            mv(c, R, L + LK_COM);
#pragma unroll
            for (int x = 0; x < 3; x++) { c[x] += pos[x]; hh[x] = m * c[x]; }
            {   // R Ic R' + m (|c|^2 1 - c c')
                double RS[9];
#pragma unroll
                for (int x = 0; x < 3; x++) {
                    RS[3 * x] = R[3 * x] * Ic[0] + R[3 * x + 1] * Ic[3] + R[3 * x + 2] * Ic[4];
                    RS[3 * x + 1] = R[3 * x] * Ic[3] + R[3 * x + 1] * Ic[1] + R[3 * x + 2] * Ic[5];
                    RS[3 * x + 2] = R[3 * x] * Ic[4] + R[3 * x + 1] * Ic[5] + R[3 * x + 2] * Ic[2];
                }
                const double cc = dot(c, c);
                Io[0] = RS[0] * R[0] + RS[1] * R[1] + RS[2] * R[2] + m * (cc - c[0] * c[0]);
                Io[1] = RS[3] * R[3] + RS[4] * R[4] + RS[5] * R[5] + m * (cc - c[1] * c[1]);
                Io[2] = RS[6] * R[6] + RS[7] * R[7] + RS[8] * R[8] + m * (cc - c[2] * c[2]);
                Io[3] = RS[0] * R[3] + RS[1] * R[4] + RS[2] * R[5] - m * c[0] * c[1];
                Io[4] = RS[0] * R[6] + RS[1] * R[7] + RS[2] * R[8] - m * c[0] * c[2];
                Io[5] = RS[3] * R[6] + RS[4] * R[7] + RS[5] * R[8] - m * c[1] * c[2];
            }
            double mn[3], ml[3], fn[3], fl[3], t3[3];
            symv(mn, Io, Vi);                       // momentum (angular about O, linear)
            cross(t3, hh, Vi + 3);
            cross(ml, Vi, hh);
#pragma unroll
            for (int x = 0; x < 3; x++) { mn[x] += t3[x]; ml[x] += m * Vi[3 + x]; }
            symv(fn, Io, Ai);                       // I A
            cross(t3, hh, Ai + 3);
            cross(fl, Ai, hh);
#pragma unroll
            for (int x = 0; x < 3; x++) { fn[x] += t3[x]; fl[x] += m * Ai[3 + x]; }
            cross(t1, Vi, mn);                      // + V x* (I V)
            cross(t2, Vi + 3, ml);
            cross(t3, Vi, ml);
#pragma unroll
            for (int x = 0; x < 3; x++) { fn[x] += t1[x] + t2[x]; fl[x] += t3[x]; }
            if (rho > 0.0 || visc > 0.0) {
                const double* B = L + LK_BOX;
                double vc[3], bw[3], bv[3], lw[3], lv[3], lT[3] = {0, 0, 0}, lF[3] = {0, 0, 0}, T[3], Fo[3];
                cross(vc, Vi, c);
#pragma unroll
                for (int x = 0; x < 3; x++) vc[x] += Vi[3 + x];
                mtv(bw, R, Vi);
                mtv(bv, R, vc);
                mv(lw, L + LK_RIN, bw);
                mv(lv, L + LK_RIN, bv);
                if (visc > 0.0) {
                    const double PI = 3.14159265358979323846, d = (B[0] + B[1] + B[2]) / 3.0;
#pragma unroll
                    for (int x = 0; x < 3; x++) { lT[x] = -PI * d * d * d * visc * lw[x]; lF[x] = -3.0 * PI * d * visc * lv[x]; }
                }
                if (rho > 0.0) {
                    const double b0 = B[0], b1 = B[1], b2 = B[2];
                    const double q0 = b0 * b0 * b0 * b0, q1 = b1 * b1 * b1 * b1, q2 = b2 * b2 * b2 * b2;
                    lF[0] -= 0.5 * rho * b1 * b2 * fabs(lv[0]) * lv[0];
                    lF[1] -= 0.5 * rho * b0 * b2 * fabs(lv[1]) * lv[1];
                    lF[2] -= 0.5 * rho * b0 * b1 * fabs(lv[2]) * lv[2];
                    lT[0] -= rho * b0 * (q1 + q2) * fabs(lw[0]) * lw[0] / 64.0;
                    lT[1] -= rho * b1 * (q0 + q2) * fabs(lw[1]) * lw[1] / 64.0;
                    lT[2] -= rho * b2 * (q0 + q1) * fabs(lw[2]) * lw[2] / 64.0;
                }
                mtv(bw, L + LK_RIN, lT);
                mtv(bv, L + LK_RIN, lF);
                mv(T, R, bw);
                mv(Fo, R, bv);
                cross(t1, c, Fo);
#pragma unroll
                for (int x = 0; x < 3; x++) { fl[x] -= Fo[x]; fn[x] -= T[x] + t1[x]; }
            }
            cm[i] = m;
#pragma unroll
            for (int x = 0; x < 3; x++) { ch[i][x] = hh[x]; F[i][x] = fn[x]; F[i][3 + x] = fl[x]; }
#pragma unroll
            for (int x = 0; x < 6; x++) cI[i][x] = Io[x];
        } else {
            cm[i] = 0.0;
#pragma unroll
            for (int x = 0; x < 3; x++) ch[i][x] = 0.0;
#pragma unroll
            for (int x = 0; x < 6; x++) { cI[i][x] = 0.0; F[i][x] = 0.0; }
        }
    }

    // ---- pass 2: leaves -> root
    bool damped = false;
#pragma unroll(UNR)
    for (int ii = 0; ii < N; ii++) {
        const int i = (NV > 0 ? NV : nv) - 1 - ii;
        if (i < 0) break;
        const double* L = lk + i * LK_STRIDE;
        const int* I = li + i * LI_STRIDE;
        const int p = SERIAL ? i - 1 : I[LI_PARENT];
        const double* Si = S[i];
        const double tau = Si[0] * F[i][0] + Si[1] * F[i][1] + Si[2] * F[i][2] + Si[3] * F[i][3] + Si[4] * F[i][4] + Si[5] * F[i][5];
        double act = 0.0;
        if (I[LI_ACT] >= 0) act = L[LK_GEAR] * fmin(fmax(u[I[LI_ACT]], L[LK_CLO]), L[LK_CHI]);
        f[i] = act - L[LK_STIFF] * (q[i] - L[LK_SREF]) - L[LK_DAMP] * v[i] - tau;
        damped = damped || L[LK_DAMP] != 0.0;
        // momentum of the composite under unit joint velocity; its projections on the ancestors' axes are column i of M
        double mn[3], ml[3], t3[3];
        symv(mn, cI[i], Si);
        cross(t3, ch[i], Si + 3);
        cross(ml, Si, ch[i]);
#pragma unroll
        for (int x = 0; x < 3; x++) { mn[x] += t3[x]; ml[x] += cm[i] * Si[3 + x]; }
        int j = i;
#pragma unroll(UNR)
        for (int step = 0; step < N; step++) {
            if (j < 0) break;
            const double* Sj = S[j];
            const double e = Sj[0] * mn[0] + Sj[1] * mn[1] + Sj[2] * mn[2] + Sj[3] * ml[0] + Sj[4] * ml[1] + Sj[5] * ml[2];
            M[i][j] = e;
            M[j][i] = e;
            j = SERIAL ? j - 1 : li[j * LI_STRIDE + LI_PARENT];
        }
        M[i][i] += L[LK_ARM];
        if (p >= 0) {
#pragma unroll
            for (int x = 0; x < 6; x++) { F[p][x] += F[i][x]; cI[p][x] += cI[i][x]; }
#pragma unroll
            for (int x = 0; x < 3; x++) ch[p][x] += ch[i][x];
            cm[p] += cm[i];
        }
    }
    if (!SERIAL) {
        // dofs on different branches do not couple: M[i][j] was only written along ancestor chains
#pragma unroll(UNR)
        for (int i = 0; i < N; i++) {
            if (i >= nv) break;
            for (int j = 0; j < i; j++) {
                bool anc = false;
                for (int k = li[i * LI_STRIDE + LI_PARENT]; k >= 0; k = li[k * LI_STRIDE + LI_PARENT]) anc = anc || k == j;
                if (!anc) { M[i][j] = 0.0; M[j][i] = 0.0; }
            }
        }
    }

    return limits_solve_integrate_inl<N>(nv, lk, li, h, M, f, damped, q, v);
}

// PLANAR: the planar instantiation (rollout_tree_planar.cuh; NV > 0), else the general 3-D one; CONTACTS: with the
// contact candidates of the model (planar only).
template <int NV, bool SERIAL, bool PLANAR, bool CONTACTS>
__global__ void __launch_bounds__(64) rollout_tree_kernel(const double* __restrict__ g_lk, const int* __restrict__ g_li,
                                                          const double* __restrict__ g_g, const double* __restrict__ g_pk,
                                                          const int* __restrict__ g_anc, int ncand, const int* __restrict__ g_cti,
                                                          const double* __restrict__ g_ctd, int nv_rt, int nu,
                                                          mjb_tree_rollout_args a) {
    constexpr int N = NV > 0 ? NV : MJB_TREE_MAX_LINKS;
    __shared__ double s_lk[N * LK_STRIDE];
    __shared__ int s_li[N * LI_STRIDE];
    __shared__ double s_g[TG_STRIDE];
    __shared__ double s_pk[PLANAR ? N * PK_STRIDE + 2 : 1];
    __shared__ int s_anc[PLANAR ? N : 1];
    __shared__ int s_cti[CONTACTS ? MJB_TREE_MAX_CAND * CTI_STRIDE : 1];
    __shared__ double s_ctd[CONTACTS ? MJB_TREE_MAX_CAND * CT_STRIDE : 1];
    const int nv = NV > 0 ? NV : nv_rt;
    // blocks are model-uniform: instance = block / blocks-per-model (one model: the plain particle numbering)
    const int bpm = (a.particles_per_model + (int)blockDim.x - 1) / (int)blockDim.x;
    const int inst = blockIdx.x / bpm, local = (blockIdx.x % bpm) * blockDim.x + threadIdx.x;
    g_lk += (long long)inst * nv * LK_STRIDE;
    for (int i = threadIdx.x; i < nv * LK_STRIDE; i += blockDim.x) s_lk[i] = g_lk[i];
    for (int i = threadIdx.x; i < nv * LI_STRIDE; i += blockDim.x) s_li[i] = g_li[i];
    if (threadIdx.x < TG_STRIDE) s_g[threadIdx.x] = g_g[threadIdx.x];
    if (PLANAR) {
        g_pk += (long long)inst * (nv * PK_STRIDE + 2);
        for (int i = threadIdx.x; i < nv * PK_STRIDE + 2; i += blockDim.x) s_pk[i] = g_pk[i];
        for (int i = threadIdx.x; i < nv; i += blockDim.x) s_anc[i] = g_anc[i];
        if (CONTACTS) {
            for (int i = threadIdx.x; i < ncand * CTI_STRIDE; i += blockDim.x) s_cti[i] = g_cti[i];
            for (int i = threadIdx.x; i < ncand * CT_STRIDE; i += blockDim.x) s_ctd[i] = g_ctd[i];
        }
    }
    __syncthreads();
    const long long k = (long long)inst * a.particles_per_model + local;
    if (local >= a.particles_per_model || k >= a.K) return;
    const int ctrl = (int)(k / a.particles_per_ctrl);
    const double* s0 = a.state + (long long)ctrl * 2 * nv;
    const double* mean = a.mean + (long long)ctrl * a.H * nu;
    double q[N], v[N], u[N];
#pragma unroll
    for (int i = 0; i < N; i++) if (i < nv) { q[i] = s0[i]; v[i] = s0[nv + i]; }
    const double inv_dt = 1.0 / (a.frame_skip * s_g[TG_DT]);
    const int d_obs = 2 * nv - a.obs_qpos_start;
    int nefc = 0;
    for (int t = 0; t < a.H; t++) {
        double a2 = 0.0;
#pragma unroll
        for (int j = 0; j < N; j++) {
            if (j >= nu) break;
            double x = mean[t * nu + j];
            if (a.noise) x += a.noise[k * a.noise_sk + t * a.noise_st + j * a.noise_sj];
            u[j] = x;
            a2 += x * x;
            if (a.actions) a.actions[k * a.act_sk + t * a.act_st + j * a.act_sj] = x;     // unclipped (wrapper :150)
        }
        double uf[PLANAR ? N : 1];
        if constexpr (PLANAR) {
            // actuator force per dof (mj_fwdActuation: the control clamped to ctrlrange, times the gear), held over the substeps
#pragma unroll
            for (int i = 0; i < N; i++) {
                const int ai = s_li[i * LI_STRIDE + LI_ACT];
                double c = 0.0;
#pragma unroll
                for (int j = 0; j < N; j++) if (j == ai) c = u[j];
                uf[i] = ai >= 0 ? s_lk[i * LK_STRIDE + LK_GEAR] * fmin(fmax(c, s_lk[i * LK_STRIDE + LK_CLO]), s_lk[i * LK_STRIDE + LK_CHI]) : 0.0;
            }
        }
        double before = 0.0, after = 0.0;           // (selects, not q[fwd_dof]: a run-time index would push q out of registers)
#pragma unroll
        for (int i = 0; i < N; i++) if (i == a.fwd_dof) before = q[i];
        for (int s = 0; s < a.frame_skip; s++) {
            if constexpr (PLANAR) nefc += planar_substep<NV, SERIAL, CONTACTS>(s_lk, s_li, s_pk, s_anc, s_g, s_pk + NV * PK_STRIDE, ncand, s_cti, s_ctd, q, v, uf);
            else nefc += substep<NV, SERIAL, false>(s_lk, s_li, s_g, nv, q, v, u);
        }
#pragma unroll
        for (int i = 0; i < N; i++) if (i == a.fwd_dof) after = q[i];
        const double reward = a.w_fwd * (after - before) * inv_dt - a.w_ctrl * a2;
        a.costs[k * a.costs_sk + t * a.costs_st] = -reward;
        if (a.states_out) {
            double* so = a.states_out + (k * a.H + t) * 2 * nv;
#pragma unroll
            for (int i = 0; i < N; i++) if (i < nv) { so[i] = q[i]; so[nv + i] = v[i]; }
        }
        if (a.next_obs) {
            double* ob = a.next_obs + (k * a.H + t) * d_obs;
#pragma unroll
            for (int i = 0; i < N; i++) if (i < nv) {
                if (i >= a.obs_qpos_start) ob[i - a.obs_qpos_start] = q[i];
                ob[nv - a.obs_qpos_start + i] = v[i];
            }
        }
    }
    if (a.nefc) a.nefc[k] = nefc;
}

}  // namespace tree
}  // namespace mjb

#ifndef MJB_HOST_EMU
namespace {
template <class T> bool upload(T** dst, const T* src, size_t n) {
    return cudaMalloc(dst, sizeof(T) * n) == cudaSuccess &&
           cudaMemcpyAsync(*dst, src, sizeof(T) * n, cudaMemcpyHostToDevice, 0) == cudaSuccess;
}
int g_use_planar = 1;
}  // namespace

extern "C" mjb_tree_model* mjb_tree_model_create(int nv, int nu, const double* link_params, const int* link_ints,
                                                 const double* globals, const double* planar_params,
                                                 const int* planar_anc, const double* planar_gravity, int n_contacts,
                                                 const int* contact_ints, const double* contact_params, int n_instances,
                                                 int device) {
    if (nv < 1 || nv > MJB_TREE_MAX_LINKS || nu < 0 || nu > nv || !link_params || !link_ints || !globals || n_instances < 1) {
        mjb::set_error(MJB_EINVAL, "mjb_tree_model_create: 1 <= nv <= %d, nu <= nv, n_instances >= 1, non-null blocks", MJB_TREE_MAX_LINKS);
        return nullptr;
    }
    if ((planar_params != nullptr) != (planar_anc != nullptr) || (planar_params != nullptr) != (planar_gravity != nullptr)) {
        mjb::set_error(MJB_EINVAL, "mjb_tree_model_create: the three planar blocks come together or not at all");
        return nullptr;
    }
    int serial = 1;
    for (int i = 0; i < nv; i++) {
        const int p = link_ints[i * LI_STRIDE + LI_PARENT];
        if (p >= i || p < -1) { mjb::set_error(MJB_EINVAL, "mjb_tree_model_create: link %d has parent %d (must precede it)", i, p); return nullptr; }
        if (p != i - 1) serial = 0;
        const int act = link_ints[i * LI_STRIDE + LI_ACT];
        if (act >= nu) { mjb::set_error(MJB_EINVAL, "mjb_tree_model_create: link %d names actuator %d of %d", i, act, nu); return nullptr; }
    }
    if (n_contacts < 0 || n_contacts > MJB_TREE_MAX_CAND || (n_contacts > 0 && (!planar_params || !contact_ints || !contact_params)) ||
        (n_contacts > 0 && !((nv == 7 && serial) || nv == 9))) {
        mjb::set_error(MJB_ENOTIMPL, "mjb_tree_model_create: contacts run in the planar instantiation only, for 7 dofs in series "
                       "and for 9-dof trees (the shapes of the reference's models), at most %d candidate pairs", MJB_TREE_MAX_CAND);
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) { mjb::set_error(MJB_ECUDA, "mjb_tree_model_create: cudaSetDevice(%d) failed", device); return nullptr; }
    mjb_tree_model* m = new mjb_tree_model{device, nv, nu, serial, planar_params != nullptr, n_instances,
                                           nullptr, nullptr, nullptr, nullptr, nullptr, n_contacts, nullptr, nullptr};
    bool ok = upload(&m->d_lk, link_params, (size_t)n_instances * nv * LK_STRIDE) && upload(&m->d_li, link_ints, (size_t)nv * LI_STRIDE) &&
              upload(&m->d_g, globals, (size_t)TG_STRIDE);
    if (ok && m->planar) {
        const size_t stride = (size_t)nv * PK_STRIDE + 2;
        double* pk = (double*)malloc(sizeof(double) * stride * n_instances);
        for (int n = 0; n < n_instances && pk; n++) {
            for (int i = 0; i < nv * PK_STRIDE; i++) pk[n * stride + i] = planar_params[(size_t)n * nv * PK_STRIDE + i];
            pk[n * stride + nv * PK_STRIDE] = planar_gravity[0];
            pk[n * stride + nv * PK_STRIDE + 1] = planar_gravity[1];
        }
        ok = pk && upload(&m->d_pk, pk, stride * n_instances) && upload(&m->d_anc, planar_anc, (size_t)nv);
        ok = ok && cudaStreamSynchronize(0) == cudaSuccess;      // pk is freed right here
        free(pk);
        if (ok && n_contacts > 0)
            ok = upload(&m->d_cti, contact_ints, (size_t)n_contacts * CTI_STRIDE) && upload(&m->d_ctd, contact_params, (size_t)n_contacts * CT_STRIDE);
    }
    if (!ok || cudaStreamSynchronize(0) != cudaSuccess) {
        mjb::set_error(MJB_ECUDA, "mjb_tree_model_create: device allocation / upload failed");
        mjb_tree_model_destroy(m);
        return nullptr;
    }
    return m;
}

extern "C" void mjb_tree_model_destroy(mjb_tree_model* m) {
    if (!m) return;
    cudaFree(m->d_lk); cudaFree(m->d_li); cudaFree(m->d_g);
    if (m->d_pk) cudaFree(m->d_pk);
    if (m->d_anc) cudaFree(m->d_anc);
    if (m->d_cti) cudaFree(m->d_cti);
    if (m->d_ctd) cudaFree(m->d_ctd);
    delete m;
}

extern "C" int mjb_tree_use_planar(int on) {
    const int old = g_use_planar;
    if (on >= 0) g_use_planar = on != 0;
    return old;
}

extern "C" void mjb_tree_layout(int* out) {
    const int v[] = {LK_RFIX, LK_OFF, LK_AXIS, LK_MASS, LK_COM, LK_IC, LK_RIN, LK_BOX, LK_ARM, LK_DAMP, LK_STIFF, LK_SREF, LK_LO,
                     LK_HI, LK_INVW, LK_SOLK, LK_SOLB, LK_SOLIMP, LK_GEAR, LK_CLO, LK_CHI, LK_STRIDE, LI_PARENT, LI_TYPE,
                     LI_LIMITED, LI_ACT, LI_BODY, LI_STRIDE, TG_DT, TG_GRAV, TG_RHO, TG_VISC, TG_STRIDE, MJB_TREE_MAX_LINKS,
                     PK_OFF, PK_DIR, PK_MASS, PK_COM, PK_INN, PK_CLIN, PK_KV1, PK_KV2, PK_E, PK_AK, PK_STRIDE,
                     CT_A, CT_HA, CT_RA, CT_B, CT_HB, CT_RB, CT_MU, CT_K, CT_BB, CT_SOLIMP, CT_INVW, CT_BOUND, CT_STRIDE, CTI_STRIDE,
                     MJB_TREE_MAX_CAND};
    for (unsigned i = 0; i < sizeof(v) / sizeof(v[0]); i++) out[i] = v[i];
}

extern "C" int mjb_rollout_tree(const mjb_tree_model* m, const mjb_tree_rollout_args* a, void* stream) {
    MJB_REQUIRE(m && a && a->state && a->mean && a->costs, "mjb_rollout_tree: null pointer");
    MJB_REQUIRE(a->K >= 1 && a->H >= 1 && a->frame_skip >= 1, "mjb_rollout_tree: K, H and frame_skip must be positive");
    MJB_REQUIRE(a->particles_per_ctrl >= 1 && a->K % a->particles_per_ctrl == 0,
                "Number of particles must be divisible by number of controllers");
    MJB_REQUIRE(a->fwd_dof >= 0 && a->fwd_dof < m->nv && a->obs_qpos_start >= 0 && a->obs_qpos_start <= m->nv,
                "mjb_rollout_tree: fwd_dof / obs_qpos_start out of range");
    MJB_REQUIRE(a->particles_per_model >= 1 && a->K % a->particles_per_model == 0 && a->K / a->particles_per_model <= m->n_inst,
                "Number of particles must be divisible by number of cpus");      /* (subproc_vec_env.py:140-141) */
    const int blocks = (a->K / a->particles_per_model) * ((a->particles_per_model + 63) / 64);
    MJB_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    using namespace mjb::tree;
    MJB_REQUIRE(m->ncand == 0 || g_use_planar, "mjb_rollout_tree: a model with contacts runs in the planar instantiation only");
    const bool planar = m->planar && g_use_planar;
#define MJB_TREE_LAUNCH(NV, SERIAL, PLANAR, CONTACTS) \
    rollout_tree_kernel<NV, SERIAL, PLANAR, CONTACTS><<<blocks, 64, 0, st>>>(m->d_lk, m->d_li, m->d_g, m->d_pk, m->d_anc, m->ncand, m->d_cti, \
                                                                         m->d_ctd, m->nv, m->nu, *a)
    if (m->ncand > 0) {
        // contacts: the two shapes the reference's models have (swimmer: 7 dofs in series; half-cheetah: 9, two legs)
        if (m->nv == 7) MJB_TREE_LAUNCH(7, true, true, true);         // (mjb_tree_model_create admits nothing else)
        else MJB_TREE_LAUNCH(9, false, true, true);
    }
    else if (planar && m->nv == 7 && m->serial) MJB_TREE_LAUNCH(7, true, true, false);
    else if (planar && m->nv == 7) MJB_TREE_LAUNCH(7, false, true, false);
    else if (planar && m->nv == 9) MJB_TREE_LAUNCH(9, false, true, false);
    // other planar mechanisms of up to 8 dofs (n-link swimmers, hoppers, walkers without contacts): run-time parents
    else if (planar && m->nv == 3) MJB_TREE_LAUNCH(3, false, true, false);
    else if (planar && m->nv == 4) MJB_TREE_LAUNCH(4, false, true, false);
    else if (planar && m->nv == 5) MJB_TREE_LAUNCH(5, false, true, false);
    else if (planar && m->nv == 6) MJB_TREE_LAUNCH(6, false, true, false);
    else if (planar && m->nv == 8) MJB_TREE_LAUNCH(8, false, true, false);
    else if (m->nv == 7 && m->serial) MJB_TREE_LAUNCH(7, true, false, false);
    else MJB_TREE_LAUNCH(0, false, false, false);
#undef MJB_TREE_LAUNCH
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}
#endif  // MJB_HOST_EMU
