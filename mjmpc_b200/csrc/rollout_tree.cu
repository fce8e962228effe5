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
//   mj_Euler                 implicit in joint damping: (M + h B) qacc = f + J'lambda
// The model block (tree_model.h) is staged in shared memory once per block: link parameters are indexed with
// run-time link numbers and every lane reads the same address (broadcast).
// Instantiations: <7, serial chain> with every loop unrolled and the per-link state in registers (the reference's
// swimmer is a serial chain of 7 dofs); <run-time nv <= 12, any tree> with the per-link state in local memory.
#include "common.h"
#include "tree_model.h"

struct mjb_tree_model {
    int device, nv, nu, serial;
    double* d_lk;   // nv x LK_STRIDE
    int* d_li;      // nv x LI_STRIDE
    double* d_g;    // TG_STRIDE
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

// One mj_step.  NV > 0: compile-time dof count (loops unroll); NV == 0: run-time nv <= MJB_TREE_MAX_LINKS.
// SERIAL: link i hangs off link i - 1.  u: controls (nu).  q, v advanced in place.  Returns the number of limit rows.
template <int NV, bool SERIAL>
TR_HD int substep(const double* lk, const int* li, const double* g, int nv_rt, double* q, double* v, const double* u) {
    constexpr int N = NV > 0 ? NV : MJB_TREE_MAX_LINKS;
    constexpr int UNR = NV > 0 ? NV : 1;      // per-link loops unroll only when the dof count is a compile-time constant
    const int nv = NV > 0 ? NV : nv_rt;
    const double h = g[TG_DT], rho = g[TG_RHO], visc = g[TG_VISC];
    double R[N][9], r[N][3], w[N][3], vl[N][3], aa[N][3], al[N][3], fl[N][3], fn[N][3];
    double cm[N], ch[N][3], cI[N][6], M[N][N], f[N];

    // ---- pass 1: root -> leaves
#pragma unroll(UNR)
    for (int i = 0; i < N; i++) {
        if (i >= nv) break;
        const double* L = lk + i * LK_STRIDE;
        const int* I = li + i * LI_STRIDE;
        const int p = SERIAL ? i - 1 : I[LI_PARENT];
        const bool hinge = I[LI_TYPE] == MJB_TREE_HINGE;
        const double* a = L + LK_AXIS;
        if (hinge) {
            double s, c;
            sincos(q[i], &s, &c);
            const double t = 1.0 - c;
            const double J[9] = {c + t * a[0] * a[0],        t * a[0] * a[1] - s * a[2], t * a[0] * a[2] + s * a[1],
                                 t * a[0] * a[1] + s * a[2], c + t * a[1] * a[1],        t * a[1] * a[2] - s * a[0],
                                 t * a[0] * a[2] - s * a[1], t * a[1] * a[2] + s * a[0], c + t * a[2] * a[2]};
#pragma unroll
            for (int x = 0; x < 3; x++)
#pragma unroll
                for (int y = 0; y < 3; y++)
                    R[i][3 * x + y] = L[LK_RFIX + 3 * x] * J[y] + L[LK_RFIX + 3 * x + 1] * J[3 + y] + L[LK_RFIX + 3 * x + 2] * J[6 + y];
            r[i][0] = L[LK_OFF]; r[i][1] = L[LK_OFF + 1]; r[i][2] = L[LK_OFF + 2];
        } else {
            double d[3];
#pragma unroll
            for (int x = 0; x < 9; x++) R[i][x] = L[LK_RFIX + x];
            mv(d, L + LK_RFIX, a);
#pragma unroll
            for (int x = 0; x < 3; x++) r[i][x] = L[LK_OFF + x] + d[x] * q[i];
        }
        double wp[3] = {0, 0, 0}, vp[3] = {0, 0, 0}, ap[3] = {0, 0, 0}, lp[3] = {-g[TG_GRAV], -g[TG_GRAV + 1], -g[TG_GRAV + 2]};
        if (p >= 0) {
#pragma unroll
            for (int x = 0; x < 3; x++) { wp[x] = w[p][x]; vp[x] = vl[p][x]; ap[x] = aa[p][x]; lp[x] = al[p][x]; }
        }
        double t1[3], t2[3];
        cross(t1, wp, r[i]);
        cross(t2, ap, r[i]);
#pragma unroll
        for (int x = 0; x < 3; x++) { t1[x] += vp[x]; t2[x] += lp[x]; }
        mtv(w[i], R[i], wp);
        mtv(vl[i], R[i], t1);
        mtv(aa[i], R[i], ap);
        mtv(al[i], R[i], t2);
        const double sj[3] = {a[0] * v[i], a[1] * v[i], a[2] * v[i]};
        if (hinge) {
            cross(t1, w[i], sj);
            cross(t2, vl[i], sj);
#pragma unroll
            for (int x = 0; x < 3; x++) { aa[i][x] += t1[x]; al[i][x] += t2[x]; w[i][x] += sj[x]; }
        } else {
            cross(t1, w[i], sj);
#pragma unroll
            for (int x = 0; x < 3; x++) { al[i][x] += t1[x]; vl[i][x] += sj[x]; }
        }
        // link wrench and its own inertia as the seed of the composite
        if (I[LI_BODY]) {
            const double m = L[LK_MASS];
            const double* c = L + LK_COM;
            const double* Ic = L + LK_IC;
            double vc[3], hl[3], hn[3], ac[3], t3[3];
            cross(vc, w[i], c);
            cross(ac, aa[i], c);
#pragma unroll
            for (int x = 0; x < 3; x++) { vc[x] += vl[i][x]; ac[x] += al[i][x]; hl[x] = m * vc[x]; fl[i][x] = m * ac[x]; }
            symv(hn, Ic, w[i]);
            cross(t3, c, hl);
#pragma unroll
            for (int x = 0; x < 3; x++) hn[x] += t3[x];
            symv(fn[i], Ic, aa[i]);
            cross(t3, c, fl[i]);
            cross(t1, w[i], hn);
            cross(t2, vl[i], hl);
#pragma unroll
            for (int x = 0; x < 3; x++) fn[i][x] += t3[x] + t1[x] + t2[x];
            cross(t1, w[i], hl);
#pragma unroll
            for (int x = 0; x < 3; x++) fl[i][x] += t1[x];
            if (rho > 0.0 || visc > 0.0) {
                const double* B = L + LK_BOX;
                double lw[3], lv[3], lT[3] = {0, 0, 0}, lF[3] = {0, 0, 0}, T[3], F[3];
                mv(lw, L + LK_RIN, w[i]);
                mv(lv, L + LK_RIN, vc);
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
                mtv(T, L + LK_RIN, lT);
                mtv(F, L + LK_RIN, lF);
                cross(t1, c, F);
#pragma unroll
                for (int x = 0; x < 3; x++) { fl[i][x] -= F[x]; fn[i][x] -= T[x] + t1[x]; }
            }
            const double cc = dot(c, c);
            cm[i] = m;
#pragma unroll
            for (int x = 0; x < 3; x++) ch[i][x] = m * c[x];
            cI[i][0] = Ic[0] + m * (cc - c[0] * c[0]); cI[i][1] = Ic[1] + m * (cc - c[1] * c[1]); cI[i][2] = Ic[2] + m * (cc - c[2] * c[2]);
            cI[i][3] = Ic[3] - m * c[0] * c[1]; cI[i][4] = Ic[4] - m * c[0] * c[2]; cI[i][5] = Ic[5] - m * c[1] * c[2];
        } else {
            cm[i] = 0.0;
#pragma unroll
            for (int x = 0; x < 3; x++) { fl[i][x] = 0.0; fn[i][x] = 0.0; ch[i][x] = 0.0; }
#pragma unroll
            for (int x = 0; x < 6; x++) cI[i][x] = 0.0;
        }
    }

    // ---- pass 2: leaves -> root
#pragma unroll(UNR)
    for (int ii = 0; ii < N; ii++) {
        const int i = (NV > 0 ? NV : nv) - 1 - ii;
        if (i < 0) break;
        const double* L = lk + i * LK_STRIDE;
        const int* I = li + i * LI_STRIDE;
        const int p = SERIAL ? i - 1 : I[LI_PARENT];
        const bool hinge = I[LI_TYPE] == MJB_TREE_HINGE;
        const double* a = L + LK_AXIS;
        // generalised force of the link wrench (bias + gravity - fluid), passive joint forces, actuation
        double tau = hinge ? dot(a, fn[i]) : dot(a, fl[i]);
        double act = 0.0;
        if (I[LI_ACT] >= 0) act = L[LK_GEAR] * fmin(fmax(u[I[LI_ACT]], L[LK_CLO]), L[LK_CHI]);
        f[i] = act - L[LK_STIFF] * (q[i] - L[LK_SREF]) - L[LK_DAMP] * v[i] - tau;
        // column i of M: momentum of the composite under unit joint velocity, walked down to the root
        double Fl[3], Fn[3];
        if (hinge) { cross(Fl, a, ch[i]); symv(Fn, cI[i], a); }
        else { Fl[0] = cm[i] * a[0]; Fl[1] = cm[i] * a[1]; Fl[2] = cm[i] * a[2]; cross(Fn, ch[i], a); }
        M[i][i] = (hinge ? dot(a, Fn) : dot(a, Fl)) + L[LK_ARM];
#pragma unroll(UNR)
        for (int jj = 0; jj < N; jj++) if (jj < i) { M[i][jj] = 0.0; }
        int j = i;
#pragma unroll(UNR)
        for (int step = 0; step < N; step++) {
            const int pj = SERIAL ? j - 1 : li[j * LI_STRIDE + LI_PARENT];
            if (pj < 0) break;
            double t1[3], t2[3];
            mv(t1, R[j], Fl);
            mv(t2, R[j], Fn);
            cross(Fn, r[j], t1);
#pragma unroll
            for (int x = 0; x < 3; x++) { Fl[x] = t1[x]; Fn[x] += t2[x]; }
            j = pj;
            const double* aj = lk + j * LK_STRIDE + LK_AXIS;
            M[i][j] = li[j * LI_STRIDE + LI_TYPE] == MJB_TREE_HINGE ? dot(aj, Fn) : dot(aj, Fl);
        }
        if (p >= 0) {
            // hand wrench and composite inertia to the parent
            double F[3], T[3], t1[3], hp[3];
            mv(F, R[i], fl[i]);
            mv(T, R[i], fn[i]);
            cross(t1, r[i], F);
#pragma unroll
            for (int x = 0; x < 3; x++) { fl[p][x] += F[x]; fn[p][x] += T[x] + t1[x]; }
            mv(hp, R[i], ch[i]);
            // rotate the inertia: R S R'
            const double* S = cI[i];
            const double* Q = R[i];
            double RS[9];
#pragma unroll
            for (int x = 0; x < 3; x++) {
                RS[3 * x] = Q[3 * x] * S[0] + Q[3 * x + 1] * S[3] + Q[3 * x + 2] * S[4];
                RS[3 * x + 1] = Q[3 * x] * S[3] + Q[3 * x + 1] * S[1] + Q[3 * x + 2] * S[5];
                RS[3 * x + 2] = Q[3 * x] * S[4] + Q[3 * x + 1] * S[5] + Q[3 * x + 2] * S[2];
            }
            const double m = cm[i];
            const double* o = r[i];
            const double oo = dot(o, o), oh = dot(o, hp);
            const double Ixx = RS[0] * Q[0] + RS[1] * Q[1] + RS[2] * Q[2], Iyy = RS[3] * Q[3] + RS[4] * Q[4] + RS[5] * Q[5];
            const double Izz = RS[6] * Q[6] + RS[7] * Q[7] + RS[8] * Q[8], Ixy = RS[0] * Q[3] + RS[1] * Q[4] + RS[2] * Q[5];
            const double Ixz = RS[0] * Q[6] + RS[1] * Q[7] + RS[2] * Q[8], Iyz = RS[3] * Q[6] + RS[4] * Q[7] + RS[5] * Q[8];
            cI[p][0] += Ixx + m * (oo - o[0] * o[0]) + 2.0 * (oh - o[0] * hp[0]);
            cI[p][1] += Iyy + m * (oo - o[1] * o[1]) + 2.0 * (oh - o[1] * hp[1]);
            cI[p][2] += Izz + m * (oo - o[2] * o[2]) + 2.0 * (oh - o[2] * hp[2]);
            cI[p][3] += Ixy - m * o[0] * o[1] - o[0] * hp[1] - hp[0] * o[1];
            cI[p][4] += Ixz - m * o[0] * o[2] - o[0] * hp[2] - hp[0] * o[2];
            cI[p][5] += Iyz - m * o[1] * o[2] - o[1] * hp[2] - hp[1] * o[2];
            cm[p] += m;
#pragma unroll
            for (int x = 0; x < 3; x++) ch[p][x] += hp[x] + m * o[x];
        }
    }
    // symmetric fill (M[i][j] set for j < i)
#pragma unroll(UNR)
    for (int i = 0; i < N; i++)
#pragma unroll(UNR)
        for (int j = 0; j < N; j++) if (j > i && j < nv) M[i][j] = M[j][i];

    // ---- joint-limit rows
    int nr = 0, rdof[N];
    double rs[N], aref[N], D[N];
    bool damped = false;
#pragma unroll(UNR)
    for (int i = 0; i < N; i++) {
        if (i >= nv) break;
        const double* L = lk + i * LK_STRIDE;
        damped = damped || L[LK_DAMP] != 0.0;
        if (!li[i * LI_STRIDE + LI_LIMITED]) continue;
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
    double A[N][N], qa[N];
    if (nr > 0) {
        double a0[N];
        for (int i = 0; i < nv; i++) { a0[i] = f[i]; for (int j = 0; j <= i; j++) A[i][j] = M[i][j]; }
        chol_solve<N>(nv, A, a0);                       // unconstrained acceleration
        double fc[N];
        unsigned act = 0;
        for (int rr = 0; rr < nr; rr++) if (rs[rr] * a0[rdof[rr]] - aref[rr] < 0.0) act |= 1u << rr;
        for (int iter = 0; iter < 40 && (act || iter); iter++) {
            // minimiser of the quadratic that holds on the current active set
            double a1[N];
            for (int i = 0; i < nv; i++) { a1[i] = f[i]; for (int j = 0; j <= i; j++) A[i][j] = M[i][j]; }
            for (int rr = 0; rr < nr; rr++)
                if (act >> rr & 1) { A[rdof[rr]][rdof[rr]] += D[rr]; a1[rdof[rr]] += rs[rr] * D[rr] * aref[rr]; }
            chol_solve<N>(nv, A, a1);
            unsigned act1 = 0;
            for (int rr = 0; rr < nr; rr++) if (rs[rr] * a1[rdof[rr]] - aref[rr] < 0.0) act1 |= 1u << rr;
            if (act1 == act) { for (int i = 0; i < nv; i++) a0[i] = a1[i]; break; }
            // the set changes along the step: exact minimiser of the piecewise quadratic on the ray a0 + t (a1 - a0)
            double p[N], g0 = 0.0, h0 = 0.0;
            for (int i = 0; i < nv; i++) p[i] = a1[i] - a0[i];
            for (int i = 0; i < nv; i++) {
                double Mp = 0.0, Ma = -f[i];
                for (int k = 0; k < nv; k++) { Mp += M[i][k] * p[k]; Ma += M[i][k] * a0[k]; }
                g0 += p[i] * Ma; h0 += p[i] * Mp;
            }
            double res[N], Jp[N], tcur = 0.0;
            for (int rr = 0; rr < nr; rr++) { res[rr] = rs[rr] * a0[rdof[rr]] - aref[rr]; Jp[rr] = rs[rr] * p[rdof[rr]]; }
            double tstar = 1.0;
            for (int seg = 0; seg <= nr; seg++) {
                // next breakpoint after tcur
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
            for (int i = 0; i < nv; i++) a0[i] += tstar * p[i];
            act = 0;
            for (int rr = 0; rr < nr; rr++) if (rs[rr] * a0[rdof[rr]] - aref[rr] < 0.0) act |= 1u << rr;
        }
        for (int i = 0; i < nv; i++) fc[i] = 0.0;
        for (int rr = 0; rr < nr; rr++) {
            const double s = rs[rr] * a0[rdof[rr]] - aref[rr];
            if (s < 0.0) fc[rdof[rr]] += rs[rr] * (-D[rr] * s);
        }
        if (!damped) {
            for (int i = 0; i < nv; i++) qa[i] = a0[i];
        } else {
            for (int i = 0; i < nv; i++) f[i] += fc[i];
        }
    }
    if (nr == 0 || damped) {
        // ---- mj_Euler, implicit in joint damping
#pragma unroll
        for (int i = 0; i < N; i++) {
            if (i >= nv) break;
            qa[i] = f[i];
#pragma unroll(UNR)
            for (int j = 0; j < N; j++) if (j <= i) A[i][j] = M[i][j];
            A[i][i] += h * lk[i * LK_STRIDE + LK_DAMP];
        }
        chol_solve<N>(nv, A, qa);
    }
#pragma unroll(UNR)
    for (int i = 0; i < N; i++) {
        if (i >= nv) break;
        v[i] += h * qa[i];
        q[i] += h * v[i];
    }
    return nr;
}

template <int NV, bool SERIAL>
__global__ void __launch_bounds__(64) rollout_tree_kernel(const double* __restrict__ g_lk, const int* __restrict__ g_li,
                                                          const double* __restrict__ g_g, int nv_rt, int nu,
                                                          mjb_tree_rollout_args a) {
    constexpr int N = NV > 0 ? NV : MJB_TREE_MAX_LINKS;
    __shared__ double s_lk[N * LK_STRIDE];
    __shared__ int s_li[N * LI_STRIDE];
    __shared__ double s_g[TG_STRIDE];
    const int nv = NV > 0 ? NV : nv_rt;
    for (int i = threadIdx.x; i < nv * LK_STRIDE; i += blockDim.x) s_lk[i] = g_lk[i];
    for (int i = threadIdx.x; i < nv * LI_STRIDE; i += blockDim.x) s_li[i] = g_li[i];
    if (threadIdx.x < TG_STRIDE) s_g[threadIdx.x] = g_g[threadIdx.x];
    __syncthreads();
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.K) return;
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
        const double before = q[a.fwd_dof];
        for (int s = 0; s < a.frame_skip; s++) nefc += substep<NV, SERIAL>(s_lk, s_li, s_g, nv, q, v, u);
        const double reward = a.w_fwd * (q[a.fwd_dof] - before) * inv_dt - a.w_ctrl * a2;
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
extern "C" mjb_tree_model* mjb_tree_model_create(int nv, int nu, const double* link_params, const int* link_ints,
                                                 const double* globals, int device) {
    if (nv < 1 || nv > MJB_TREE_MAX_LINKS || nu < 0 || nu > nv || !link_params || !link_ints || !globals) {
        mjb::set_error(MJB_EINVAL, "mjb_tree_model_create: 1 <= nv <= %d, nu <= nv, non-null blocks", MJB_TREE_MAX_LINKS);
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
    if (cudaSetDevice(device) != cudaSuccess) { mjb::set_error(MJB_ECUDA, "mjb_tree_model_create: cudaSetDevice(%d) failed", device); return nullptr; }
    mjb_tree_model* m = new mjb_tree_model{device, nv, nu, serial, nullptr, nullptr, nullptr};
    if (cudaMalloc(&m->d_lk, sizeof(double) * nv * LK_STRIDE) != cudaSuccess || cudaMalloc(&m->d_li, sizeof(int) * nv * LI_STRIDE) != cudaSuccess ||
        cudaMalloc(&m->d_g, sizeof(double) * TG_STRIDE) != cudaSuccess ||
        cudaMemcpyAsync(m->d_lk, link_params, sizeof(double) * nv * LK_STRIDE, cudaMemcpyHostToDevice, 0) != cudaSuccess ||
        cudaMemcpyAsync(m->d_li, link_ints, sizeof(int) * nv * LI_STRIDE, cudaMemcpyHostToDevice, 0) != cudaSuccess ||
        cudaMemcpyAsync(m->d_g, globals, sizeof(double) * TG_STRIDE, cudaMemcpyHostToDevice, 0) != cudaSuccess ||
        cudaStreamSynchronize(0) != cudaSuccess) {
        mjb::set_error(MJB_ECUDA, "mjb_tree_model_create: device allocation / upload failed");
        cudaFree(m->d_lk); cudaFree(m->d_li); cudaFree(m->d_g);
        delete m;
        return nullptr;
    }
    return m;
}

extern "C" void mjb_tree_model_destroy(mjb_tree_model* m) {
    if (!m) return;
    cudaFree(m->d_lk); cudaFree(m->d_li); cudaFree(m->d_g);
    delete m;
}

extern "C" void mjb_tree_layout(int* out) {
    const int v[] = {LK_RFIX, LK_OFF, LK_AXIS, LK_MASS, LK_COM, LK_IC, LK_RIN, LK_BOX, LK_ARM, LK_DAMP, LK_STIFF, LK_SREF, LK_LO,
                     LK_HI, LK_INVW, LK_SOLK, LK_SOLB, LK_SOLIMP, LK_GEAR, LK_CLO, LK_CHI, LK_STRIDE, LI_PARENT, LI_TYPE,
                     LI_LIMITED, LI_ACT, LI_BODY, LI_STRIDE, TG_DT, TG_GRAV, TG_RHO, TG_VISC, TG_STRIDE, MJB_TREE_MAX_LINKS};
    for (unsigned i = 0; i < sizeof(v) / sizeof(v[0]); i++) out[i] = v[i];
}

extern "C" int mjb_rollout_tree(const mjb_tree_model* m, const mjb_tree_rollout_args* a, void* stream) {
    MJB_REQUIRE(m && a && a->state && a->mean && a->costs, "mjb_rollout_tree: null pointer");
    MJB_REQUIRE(a->K >= 1 && a->H >= 1 && a->frame_skip >= 1, "mjb_rollout_tree: K, H and frame_skip must be positive");
    MJB_REQUIRE(a->particles_per_ctrl >= 1 && a->K % a->particles_per_ctrl == 0,
                "Number of particles must be divisible by number of controllers");
    MJB_REQUIRE(a->fwd_dof >= 0 && a->fwd_dof < m->nv && a->obs_qpos_start >= 0 && a->obs_qpos_start <= m->nv,
                "mjb_rollout_tree: fwd_dof / obs_qpos_start out of range");
    const int blocks = (a->K + 63) / 64;
    cudaStream_t st = (cudaStream_t)stream;
    if (m->nv == 7 && m->serial)
        mjb::tree::rollout_tree_kernel<7, true><<<blocks, 64, 0, st>>>(m->d_lk, m->d_li, m->d_g, m->nv, m->nu, *a);
    else
        mjb::tree::rollout_tree_kernel<0, false><<<blocks, 64, 0, st>>>(m->d_lk, m->d_li, m->d_g, m->nv, m->nu, *a);
    MJB_CUDA(cudaGetLastError());
    return MJB_OK;
}
#endif  // MJB_HOST_EMU
