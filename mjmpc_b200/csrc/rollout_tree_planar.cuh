// Planar instantiation of the tree rollout kernel (included by rollout_tree.cu).
//
// A tree whose hinge axes are all parallel (to n) and whose slides are all perpendicular to n moves in a plane -- the
// reference's swimmer.xml (n = z) and half_cheetah.xml (n = y) both do.  The 6-vectors of the general kernel collapse
// to (omega; vx, vy), rotations to (cos, sin) of the accumulated hinge angles, spatial inertias to (m, m c, I_nn): ~4 x
// fewer FP64 operations, ~10 doubles of state per link -- so every per-link loop unrolls with the state in registers AND
// the code stays inside the instruction caches (the supply of instructions, not the FP64 pipe, is what the
// general instantiation waits for).  The reduction is exact: only n'I n enters the in-plane equations, and the
// inertia-box fluid model of mj_passive needs the projections of the body's inertial axes on the plane (they turn with
// the body) and their constant components along n (mjmpc_b200/envs/mjcf_tree.py: pack_planar).
// Link frames are world-aligned at q = 0.  Parents are run-time data: a child picks its parent's state with predicated
// selects over the links before it (static register indexing), or i - 1 when SERIAL.
// The constrained solve and mj_Euler share one unrolled factorisation in registers, iterated over the active set of
// limit rows; an active set that does not settle within 10 iterations goes to limits_solve_integrate (Newton with an
// exact line search, out of line).
#pragma once

enum { PK_OFF = 0, PK_DIR = 2, PK_MASS = 4, PK_COM = 5, PK_INN = 7, PK_CLIN = 8, PK_KV1 = 9, PK_KV2 = 10, PK_E = 11, PK_AK = 17,
       PK_STRIDE = 20 };

namespace mjb {
namespace tree {

TR_HD double inv_sqrt(double x) {
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}

template <int NV, bool SERIAL>
TR_HD int planar_substep(const double* lk, const int* li, const double* pk, const int* anc, const double* g, const double* gp,
                         double* q, double* v, const double* u /* per dof: gear * clamp(ctrl) */) {
    const double h = g[TG_DT];
    const bool fluid = g[TG_RHO] > 0.0 || g[TG_VISC] > 0.0;
    double cs[NV], sn[NV], px[NV], py[NV], Vw[NV], Vx[NV], Vy[NV], Ax[NV], Ay[NV];      // bias acceleration has no angular part
    double Sw[NV], Sx[NV], Sy[NV], Fw[NV], Fx[NV], Fy[NV], cm[NV], chx[NV], chy[NV], cI[NV];
    double M[NV][NV], f[NV];
    double rD[NV], rS[NV], rA[NV];              // limit row of dof i (J = rS e_i): 1 / R, side, reference acceleration
    double Ox = 0.0, Oy = 0.0;
    unsigned rows = 0;
    bool damped = false;

#pragma unroll
    for (int i = 0; i < NV; i++) {
        const double* L = lk + i * LK_STRIDE;
        const double* P = pk + i * PK_STRIDE;
        const int* I = li + i * LI_STRIDE;
        const int p = SERIAL ? i - 1 : I[LI_PARENT];
        const bool hinge = I[LI_TYPE] == MJB_TREE_HINGE;
        double cp = 1.0, sp = 0.0, ppx = -Ox, ppy = -Oy, wp = 0.0, vxp = 0.0, vyp = 0.0, axp = -gp[0], ayp = -gp[1];
        if (SERIAL) {
            if (i > 0) { cp = cs[i - 1]; sp = sn[i - 1]; ppx = px[i - 1]; ppy = py[i - 1]; wp = Vw[i - 1]; vxp = Vx[i - 1]; vyp = Vy[i - 1]; axp = Ax[i - 1]; ayp = Ay[i - 1]; }
        } else {
#pragma unroll
            for (int j = 0; j < i; j++)
                if (j == p) { cp = cs[j]; sp = sn[j]; ppx = px[j]; ppy = py[j]; wp = Vw[j]; vxp = Vx[j]; vyp = Vy[j]; axp = Ax[j]; ayp = Ay[j]; }
        }
        // link origin and orientation
        double ox = P[PK_OFF], oy = P[PK_OFF + 1], c = cp, s = sp, dx = 0.0, dy = 0.0;
        if (hinge) {
            double sj, cj;
            sincos_lean(P[PK_DIR] * q[i], &sj, &cj);
            c = cp * cj - sp * sj;
            s = sp * cj + cp * sj;
        } else {
            dx = cp * P[PK_DIR] - sp * P[PK_DIR + 1];       // slide direction, turned with the parent
            dy = sp * P[PK_DIR] + cp * P[PK_DIR + 1];
        }
        double x = ppx + cp * ox - sp * oy, y = ppy + sp * ox + cp * oy;
        if (!hinge) { x += dx * q[i]; y += dy * q[i]; }
        if (i == 0) { Ox = x; Oy = y; x = 0.0; y = 0.0; }
        // motion axis about O: hinge sigma (1; y, -x), slide (0; d)
        const double sw = hinge ? P[PK_DIR] : 0.0;
        const double sx = hinge ? sw * y : dx, sy = hinge ? -sw * x : dy;
        // bias acceleration: A += (V_parent x S) qdot = ( wp (-sy, sx) + sw (vyp, -vxp) ) qdot
        const double qd = v[i];
        const double ax = axp + (wp * -sy + sw * vyp) * qd, ay = ayp + (wp * sx - sw * vxp) * qd;
        const double w = wp + sw * qd, vx = vxp + sx * qd, vy = vyp + sy * qd;
        cs[i] = c; sn[i] = s; px[i] = x; py[i] = y; Vw[i] = w; Vx[i] = vx; Vy[i] = vy; Ax[i] = ax; Ay[i] = ay;
        Sw[i] = sw; Sx[i] = sx; Sy[i] = sy;
        if (I[LI_BODY] & 1) {
            const double m = P[PK_MASS];
            const double cx = x + c * P[PK_COM] - s * P[PK_COM + 1], cy = y + s * P[PK_COM] + c * P[PK_COM + 1];
            const double hx = m * cx, hy = m * cy, Io = P[PK_INN] + m * (cx * cx + cy * cy);
            // momentum (about O) and wrench: f = I A + V x* (I V); the bias acceleration is purely linear
            const double mn = Io * w + hx * vy - hy * vx, mlx = m * vx - w * hy, mly = m * vy + w * hx;
            double fw = hx * ay - hy * ax + vx * mly - vy * mlx;
            double fx = m * ax - w * mly, fy = m * ay + w * mlx;
            (void)mn;
            if (fluid) {
                const double vcx = vx - w * cy, vcy = vy + w * cx;      // velocity of the centre of mass
                double Fx_ = 0.0, Fy_ = 0.0;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const double e0x = P[PK_E + 2 * k], e0y = P[PK_E + 2 * k + 1];
                    const double ex = c * e0x - s * e0y, ey = s * e0x + c * e0y;
                    const double lv = ex * vcx + ey * vcy;
                    const double lf = -P[PK_CLIN] * lv - P[PK_AK + k] * fabs(lv) * lv;
                    Fx_ += lf * ex; Fy_ += lf * ey;
                }
                const double T = -P[PK_KV1] * w - P[PK_KV2] * fabs(w) * w;
                fx -= Fx_; fy -= Fy_;
                fw -= T + cx * Fy_ - cy * Fx_;
            }
            Fw[i] = fw; Fx[i] = fx; Fy[i] = fy; cm[i] = m; chx[i] = hx; chy[i] = hy; cI[i] = Io;
        } else {
            Fw[i] = 0.0; Fx[i] = 0.0; Fy[i] = 0.0; cm[i] = 0.0; chx[i] = 0.0; chy[i] = 0.0; cI[i] = 0.0;
        }
        rD[i] = 0.0; rS[i] = 0.0; rA[i] = 0.0;
        if (I[LI_LIMITED]) {
            const double dlo = q[i] - L[LK_LO], dhi = L[LK_HI] - q[i];
            if (dlo < 0.0 || dhi < 0.0) {
                const double side = dlo < 0.0 ? 1.0 : -1.0, dist = dlo < 0.0 ? dlo : dhi;
                const double imp = impedance_call(L + LK_SOLIMP, dist);
                rD[i] = fmin(1e15, imp / ((1.0 - imp) * L[LK_INVW]));       // 1 / max(1e-15, (1 - imp) invweight / imp)
                rS[i] = side;
                rA[i] = -L[LK_SOLB] * (side * v[i]) - L[LK_SOLK] * imp * dist;
                rows |= 1u << i;
            }
        }
        damped = damped || L[LK_DAMP] != 0.0;
    }

#pragma unroll
    for (int ii = 0; ii < NV; ii++) {
        const int i = NV - 1 - ii;
        const double* L = lk + i * LK_STRIDE;
        const int* I = li + i * LI_STRIDE;
        const int p = SERIAL ? i - 1 : I[LI_PARENT];
        const double tau = Sw[i] * Fw[i] + Sx[i] * Fx[i] + Sy[i] * Fy[i];
        const double act = u[i];                    // per-dof actuator force, clamped and geared once per env step
        f[i] = act - L[LK_STIFF] * (q[i] - L[LK_SREF]) - L[LK_DAMP] * v[i] - tau;
        // momentum of the composite under unit joint velocity
        const double mn = cI[i] * Sw[i] + chx[i] * Sy[i] - chy[i] * Sx[i];
        const double mlx = cm[i] * Sx[i] - Sw[i] * chy[i], mly = cm[i] * Sy[i] + Sw[i] * chx[i];
        M[i][i] = Sw[i] * mn + Sx[i] * mlx + Sy[i] * mly + L[LK_ARM];
        const int am = SERIAL ? ((1 << i) - 1) : anc[i];
#pragma unroll
        for (int j = 0; j < i; j++) {
            const double e = (am >> j & 1) ? Sw[j] * mn + Sx[j] * mlx + Sy[j] * mly : 0.0;
            M[i][j] = e;
            M[j][i] = e;
        }
        if (SERIAL) {
            if (i > 0) { Fw[i - 1] += Fw[i]; Fx[i - 1] += Fx[i]; Fy[i - 1] += Fy[i]; cm[i - 1] += cm[i]; chx[i - 1] += chx[i]; chy[i - 1] += chy[i]; cI[i - 1] += cI[i]; }
        } else {
#pragma unroll
            for (int j = 0; j < i; j++)
                if (j == p) { Fw[j] += Fw[i]; Fx[j] += Fx[i]; Fy[j] += Fy[i]; cm[j] += cm[i]; chx[j] += chx[i]; chy[j] += chy[i]; cI[j] += cI[i]; }
        }
    }

    // Constrained solve + mj_Euler around ONE factorisation in registers.  phase 0: active-set iterations on
    // M + sum_active D e e' (a set that reproduces itself is the exact minimiser of the convex piecewise-quadratic
    // problem, whatever set the iteration started from: the first guess comes from the decoupled accelerations
    // f_i / M_ii, which is right for most rows and saves the unconstrained solve); phase 2: mj_Euler's solve, implicit in joint
    // damping, with the constraint force on the right-hand side.  No row: phase 2 at once.  A set that keeps changing
    // (it can cycle without a line search) goes to the out-of-line Newton solver with its exact line search.
    double A[NV][NV], b[NV], fc[NV];
    unsigned act = 0;
    int phase = rows ? 0 : 2;
    bool done = false;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        fc[i] = 0.0;
        if ((rows >> i & 1) && rS[i] * f[i] < rA[i] * M[i][i]) act |= 1u << i;      // M_ii > 0
    }
#pragma unroll 1
    for (int it = 0; it < 10 && !done; it++) {
#pragma unroll
        for (int i = 0; i < NV; i++) {
            const bool on = phase == 0 && (act >> i & 1);
            b[i] = f[i] + (phase == 2 ? fc[i] : 0.0) + (on ? rS[i] * rD[i] * rA[i] : 0.0);
#pragma unroll
            for (int j = 0; j < i; j++) A[i][j] = M[i][j];
            A[i][i] = M[i][i] + (phase == 2 ? h * lk[i * LK_STRIDE + LK_DAMP] : 0.0) + (on ? rD[i] : 0.0);
        }
#pragma unroll
        for (int j = 0; j < NV; j++) {
            double d = A[j][j];
#pragma unroll
            for (int k = 0; k < j; k++) d -= A[j][k] * A[j][k];
            const double inv = inv_sqrt(d);
            A[j][j] = inv;
#pragma unroll
            for (int i = j + 1; i < NV; i++) {
                double t = A[i][j];
#pragma unroll
                for (int k = 0; k < j; k++) t -= A[i][k] * A[j][k];
                A[i][j] = t * inv;
            }
        }
#pragma unroll
        for (int i = 0; i < NV; i++) {
            double t = b[i];
#pragma unroll
            for (int k = 0; k < i; k++) t -= A[i][k] * b[k];
            b[i] = t * A[i][i];
        }
#pragma unroll
        for (int ii = 0; ii < NV; ii++) {
            const int i = NV - 1 - ii;
            double t = b[i];
#pragma unroll
            for (int k = i + 1; k < NV; k++) t -= A[k][i] * b[k];
            b[i] = t * A[i][i];
        }
        if (phase == 2) { done = true; break; }
        unsigned na = 0;
#pragma unroll
        for (int i = 0; i < NV; i++) if ((rows >> i & 1) && rS[i] * b[i] - rA[i] < 0.0) na |= 1u << i;
        if (na == act) {
            if (!damped) { done = true; break; }
#pragma unroll
            for (int i = 0; i < NV; i++) fc[i] = (act >> i & 1) ? rS[i] * (-rD[i] * (rS[i] * b[i] - rA[i])) : 0.0;
            phase = 2;
        } else {
            act = na;
        }
    }
    if (!done) {
        // (copies, so that M itself never has its address taken and stays in registers)
        double Ml[NV][NV], fl[NV];
#pragma unroll
        for (int i = 0; i < NV; i++) {
            fl[i] = f[i];
#pragma unroll
            for (int j = 0; j < NV; j++) Ml[i][j] = M[i][j];
        }
        return limits_solve_integrate<NV>(NV, lk, li, h, Ml, fl, damped, q, v);
    }
    int nr = 0;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        v[i] += h * b[i];
        q[i] += h * v[i];
        nr += rows >> i & 1;
    }
    return nr;
}

}  // namespace tree
}  // namespace mjb
