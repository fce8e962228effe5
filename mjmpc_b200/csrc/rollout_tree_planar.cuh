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

// Contact candidates of a planar mechanism (mjcf_tree.pack_planar_contacts).  Per candidate, ints: kind (0 segment /
// segment, 1 world plane / the two end spheres of a segment), link 1 (-1 = world), link 2; doubles (link frames
// world-aligned at q = 0, relative to the link's anchor): segment A centre (2), half axis (2), radius -- for a plane:
// a point (2, absolute) and the normal (2) -- segment B centre, half axis, radius, then mu, K, B, solimp (5),
// body_invweight0 sum.
enum { CT_A = 0, CT_HA = 2, CT_RA = 4, CT_B = 5, CT_HB = 7, CT_RB = 9, CT_MU = 10, CT_K = 11, CT_BB = 12, CT_SOLIMP = 13, CT_INVW = 18,
       CT_BOUND = 19,       // |half axis A| + |half axis B| + radius A + radius B: centres farther apart than this cannot touch
       CT_STRIDE = 20, CTI_STRIDE = 3, MJB_TREE_MAX_CAND = 16, MJB_TREE_MAX_DENSE = 3 * 2 * MJB_TREE_MAX_CAND };

namespace mjb {
namespace tree {

// Constrained solve + mj_Euler with DENSE rows (contacts) next to the joint-limit rows: Newton on the convex
// piecewise-quadratic problem with an exact line search along every step (the oracle's algorithm), out of line.  Limit
// rows come per dof (rD = 0: none): J = rS e_i; dense rows as (Jd, Dd, Ad)[nd].  M, f are clobbered.
template <int NV>
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#endif
void dense_solve_integrate(const double* lk, double h, double (*M)[NV], double* f, bool damped, const double* rD, const double* rS,
                           const double* rA, int nd, const double (*Jd)[NV], const double* Dd, const double* Ad, double* q, double* v) {
    double A[NV][NV], a[NV], g[NV], p[NV], fc[NV];
    double res[NV + MJB_TREE_MAX_DENSE], Jp[NV + MJB_TREE_MAX_DENSE];
    for (int i = 0; i < NV; i++) { a[i] = f[i]; for (int j = 0; j <= i; j++) A[i][j] = M[i][j]; }
    chol_solve<NV>(NV, A, a);
    double fn = 0.0;
    for (int i = 0; i < NV; i++) fn += f[i] * f[i];
    for (int iter = 0; iter < 100; iter++) {
        // residuals: rows 0..NV-1 are the limit rows (absent: rD = 0), NV.. the dense ones
        for (int i = 0; i < NV; i++) res[i] = rD[i] > 0.0 ? rS[i] * a[i] - rA[i] : 1.0;
        for (int r = 0; r < nd; r++) {
            double sres = -Ad[r];
            for (int k = 0; k < NV; k++) sres += Jd[r][k] * a[k];
            res[NV + r] = sres;
        }
        double gn = 0.0;
        for (int i = 0; i < NV; i++) {
            double t = -f[i];
            for (int k = 0; k < NV; k++) t += M[i][k] * a[k];
            if (res[i] < 0.0) t += rD[i] * res[i] * rS[i];
            for (int r = 0; r < nd; r++) if (res[NV + r] < 0.0) t += Dd[r] * res[NV + r] * Jd[r][i];
            g[i] = t; gn += t * t;
        }
        if (sqrt(gn) <= 1e-15 * (1.0 + sqrt(fn))) break;
        for (int i = 0; i < NV; i++) {
            p[i] = -g[i];
            for (int j = 0; j <= i; j++) {
                double t = M[i][j];
                for (int r = 0; r < nd; r++) if (res[NV + r] < 0.0) t += Dd[r] * Jd[r][i] * Jd[r][j];
                A[i][j] = t;
            }
            if (res[i] < 0.0) A[i][i] += rD[i];
        }
        chol_solve<NV>(NV, A, p);
        double g0 = 0.0, h0 = 0.0;
        for (int i = 0; i < NV; i++) {
            double Mp = 0.0, Ma = -f[i];
            for (int k = 0; k < NV; k++) { Mp += M[i][k] * p[k]; Ma += M[i][k] * a[k]; }
            g0 += p[i] * Ma; h0 += p[i] * Mp;
        }
        for (int i = 0; i < NV; i++) Jp[i] = rD[i] > 0.0 ? rS[i] * p[i] : 0.0;
        for (int r = 0; r < nd; r++) { double t = 0.0; for (int k = 0; k < NV; k++) t += Jd[r][k] * p[k]; Jp[NV + r] = t; }
        const int nrow = NV + nd;
        double tcur = 0.0, tstar = 1.0;
        for (int seg = 0; seg <= nrow; seg++) {
            double tnext = TR_BIG;
            for (int r = 0; r < nrow; r++)
                if (Jp[r] != 0.0) { const double tb = -res[r] / Jp[r]; if (tb > tcur && tb < tnext) tnext = tb; }
            const double tmid = tnext >= TR_BIG ? tcur + 1.0 : 0.5 * (tcur + tnext);
            double c0 = g0, c1 = h0;
            for (int r = 0; r < nrow; r++) {
                const double Dr = r < NV ? rD[r] : Dd[r - NV];
                if (Dr > 0.0 && res[r] + tmid * Jp[r] < 0.0) { c0 += Dr * res[r] * Jp[r]; c1 += Dr * Jp[r] * Jp[r]; }
            }
            const double t = -c0 / c1;
            if (t <= tnext || tnext >= TR_BIG) { tstar = t < tcur ? tcur : t; break; }
            tcur = tnext;
        }
        for (int i = 0; i < NV; i++) a[i] += tstar * p[i];
        // a full Newton step that crossed no breakpoint landed on the minimiser of the quadratic it was built from, and
        // that quadratic still holds there: optimal (what remains of the gradient is rounding)
        if (fabs(tstar - 1.0) < 1e-9 && tcur == 0.0) {
            bool same = true;
            for (int r = 0; r < nrow && same; r++) {
                const double Dr = r < NV ? rD[r] : Dd[r - NV];
                if (Dr > 0.0 && ((res[r] < 0.0) != (res[r] + tstar * Jp[r] < 0.0))) same = false;
            }
            if (same) break;
        }
    }
    if (damped) {
        for (int i = 0; i < NV; i++) fc[i] = rD[i] > 0.0 && rS[i] * a[i] - rA[i] < 0.0 ? rS[i] * (-rD[i] * (rS[i] * a[i] - rA[i])) : 0.0;
        for (int r = 0; r < nd; r++) {
            double sres = -Ad[r];
            for (int k = 0; k < NV; k++) sres += Jd[r][k] * a[k];
            if (sres < 0.0) for (int k = 0; k < NV; k++) fc[k] += Jd[r][k] * (-Dd[r] * sres);
        }
        for (int i = 0; i < NV; i++) {
            a[i] = f[i] + fc[i];
            for (int j = 0; j <= i; j++) A[i][j] = M[i][j];
            A[i][i] += h * lk[i * LK_STRIDE + LK_DAMP];
        }
        chol_solve<NV>(NV, A, a);
    }
    for (int i = 0; i < NV; i++) { v[i] += h * a[i]; q[i] += h * v[i]; }
}

TR_HD double inv_sqrt(double x) {
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}

template <int NV, bool SERIAL, bool CONTACTS>
TR_HD int planar_substep(const double* lk, const int* li, const double* pk, const int* anc, const double* g, const double* gp,
                         int ncand, const int* cti, const double* ctd,
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

    // ---- contacts (models compiled with allow_contacts="model"): detection on the link frames, three rows per
    // contact -- n + mu t, n - mu t and the pair n +- mu t_out merged into one row of twice the weight (t_out is normal
    // to the plane of motion: both rows have the Jacobian of n) -- and the dense Newton, out of line
    double Jd[CONTACTS ? MJB_TREE_MAX_DENSE : 1][NV], Dd[CONTACTS ? MJB_TREE_MAX_DENSE : 1], Ad[CONTACTS ? MJB_TREE_MAX_DENSE : 1];
    int nd = 0, ncontact = 0;
    if (CONTACTS) {
        double lf[NV][4];
#pragma unroll
        for (int j = 0; j < NV; j++) { lf[j][0] = cs[j]; lf[j][1] = sn[j]; lf[j][2] = px[j]; lf[j][3] = py[j]; }
#pragma unroll 1
        for (int cnd = 0; cnd < ncand; cnd++) {
            const int* CI = cti + cnd * CTI_STRIDE;
            const double* C = ctd + cnd * CT_STRIDE;
            const int l1 = CI[1], l2 = CI[2];
            // link frames by run-time link number: from the local-memory copy (a select chain over the register arrays
            // cost 56 predicated moves per candidate)
            double c1 = 1.0, s1 = 0.0, x1 = -Ox, y1 = -Oy;
            if (l1 >= 0) { c1 = lf[l1][0]; s1 = lf[l1][1]; x1 = lf[l1][2]; y1 = lf[l1][3]; }
            const double c2 = lf[l2][0], s2 = lf[l2][1], x2 = lf[l2][2], y2 = lf[l2][3];
            const unsigned m1 = l1 < 0 ? 0u : ((SERIAL ? ((1u << l1) - 1u) : (unsigned)anc[l1]) | (1u << l1));
            const unsigned m2 = (SERIAL ? ((1u << l2) - 1u) : (unsigned)anc[l2]) | (1u << l2);
            // segment B in the world (relative to O)
            const double bx = x2 + c2 * C[CT_B] - s2 * C[CT_B + 1], by = y2 + s2 * C[CT_B] + c2 * C[CT_B + 1];
            const double hbx = c2 * C[CT_HB] - s2 * C[CT_HB + 1], hby = s2 * C[CT_HB] + c2 * C[CT_HB + 1];
            double cpx[2], cpy[2], cnx[2], cny[2], cdist[2];
            int nc = 0;
            if (CI[0] == 1) {
                const double nx = C[CT_HA], ny = C[CT_HA + 1];
                const bool sphere = C[CT_HB] == 0.0 && C[CT_HB + 1] == 0.0;
                // the whole capsule above the plane: neither end sphere can touch
                if ((bx + Ox - C[CT_A]) * nx + (by + Oy - C[CT_A + 1]) * ny - fabs(hbx * nx + hby * ny) - C[CT_RB] >= 0.0) continue;
                for (int e = 0; e < (sphere ? 1 : 2); e++) {
                    const double sg = sphere ? 0.0 : (e ? -1.0 : 1.0);
                    const double ex = bx + sg * hbx, ey = by + sg * hby;
                    const double dist = (ex + Ox - C[CT_A]) * nx + (ey + Oy - C[CT_A + 1]) * ny - C[CT_RB];
                    if (dist < 0.0) {
                        cnx[nc] = nx; cny[nc] = ny; cdist[nc] = dist;
                        cpx[nc] = ex - nx * (C[CT_RB] + 0.5 * dist); cpy[nc] = ey - ny * (C[CT_RB] + 0.5 * dist);
                        nc++;
                    }
                }
            } else {
                const double ax = x1 + c1 * C[CT_A] - s1 * C[CT_A + 1], ay = y1 + s1 * C[CT_A] + c1 * C[CT_A + 1];
                const double hax = c1 * C[CT_HA] - s1 * C[CT_HA + 1], hay = s1 * C[CT_HA] + c1 * C[CT_HA + 1];
                const double dfx = ax - bx, dfy = ay - by;
                if (dfx * dfx + dfy * dfy >= C[CT_BOUND] * C[CT_BOUND]) continue;      // bounding circles apart: no contact
                const double ma = hax * hax + hay * hay, mb = -(hax * hbx + hay * hby), mc = hbx * hbx + hby * hby;
                const double uu = -(hax * dfx + hay * dfy), ww = hbx * dfx + hby * dfy;
                const double det = ma * mc - mb * mb;
                double t1 = 0.0, t2 = 0.0;
                if (fabs(det) >= 1e-15) {                 // mjc_CapsuleCapsule, general position: sequential clamping
                    t1 = (mc * uu - mb * ww) / det; t2 = (ma * ww - mb * uu) / det;
                    if (t1 > 1.0) { t1 = 1.0; t2 = (ww - mb) / mc; } else if (t1 < -1.0) { t1 = -1.0; t2 = (ww + mb) / mc; }
                    if (t2 > 1.0) { t2 = 1.0; t1 = fmin(1.0, fmax(-1.0, (uu - mb) / ma)); }
                    else if (t2 < -1.0) { t2 = -1.0; t1 = fmin(1.0, fmax(-1.0, (uu + mb) / ma)); }
                } else {                                  // a sphere against a segment, or parallel segments
                    if (ma > 1e-15) t1 = fmin(1.0, fmax(-1.0, uu / ma));
                    if (mc > 1e-15) t2 = fmin(1.0, fmax(-1.0, (ww - mb * t1) / mc));
                }
                const double v1x = ax + hax * t1, v1y = ay + hay * t1, ddx = bx + hbx * t2 - v1x, ddy = by + hby * t2 - v1y;
                const double cd = sqrt(ddx * ddx + ddy * ddy), dist = cd - C[CT_RA] - C[CT_RB];
                if (dist < 0.0 && cd > 1e-15) {
                    cnx[0] = ddx / cd; cny[0] = ddy / cd; cdist[0] = dist;
                    cpx[0] = v1x + cnx[0] * (C[CT_RA] + 0.5 * dist); cpy[0] = v1y + cny[0] * (C[CT_RA] + 0.5 * dist);
                    nc = 1;
                }
            }
            for (int e = 0; e < nc && nd + 3 <= MJB_TREE_MAX_DENSE; e++) {
                const double nx = cnx[e], ny = cny[e], tx = -ny, ty = nx, mu = C[CT_MU];
                const double imp = impedance_call(C + CT_SOLIMP, cdist[e]);
                const double R0 = fmax(1e-15, (1.0 - imp) * ((1.0 + mu * mu) * C[CT_INVW]) / imp);
                const double Dr = 1.0 / (2.0 * mu * mu * R0);
                double vn = 0.0, vt = 0.0;
#pragma unroll
                for (int j = 0; j < NV; j++) {
                    // velocity of the contact point per unit qdot_j: body 2 minus body 1
                    const double sgn = ((m2 >> j & 1u) ? 1.0 : 0.0) - ((m1 >> j & 1u) ? 1.0 : 0.0);
                    const double jx = sgn * (Sx[j] - Sw[j] * cpy[e]), jy = sgn * (Sy[j] + Sw[j] * cpx[e]);
                    const double jn = nx * jx + ny * jy, jt = tx * jx + ty * jy;
                    Jd[nd][j] = jn + mu * jt; Jd[nd + 1][j] = jn - mu * jt; Jd[nd + 2][j] = jn;
                    vn += jn * v[j]; vt += jt * v[j];
                }
                const double pen = C[CT_K] * imp * cdist[e];
                Dd[nd] = Dr; Dd[nd + 1] = Dr; Dd[nd + 2] = 2.0 * Dr;
                Ad[nd] = -C[CT_BB] * (vn + mu * vt) - pen; Ad[nd + 1] = -C[CT_BB] * (vn - mu * vt) - pen; Ad[nd + 2] = -C[CT_BB] * vn - pen;
                nd += 3;
                ncontact++;
            }
        }
    }

    // Constrained solve + mj_Euler around ONE factorisation in registers.  phase 0: active-set iterations on
    // M + sum_active D e e' (a set that reproduces itself is the exact minimiser of the convex piecewise-quadratic
    // problem, whatever set the iteration started from: the first guess comes from the decoupled accelerations
    // f_i / M_ii, which is right for most rows and saves the unconstrained solve); phase 2: mj_Euler's solve, implicit in joint
    // damping, with the constraint force on the right-hand side.  No row: phase 2 at once.  A set that keeps changing
    // (it can cycle without a line search) goes to the out-of-line Newton solver with its exact line search.
    // Dense rows (contacts) join the same iteration: active ones add D J J' to the matrix before it is factored -- the
    // row comes from local memory, the matrix stays in registers -- first guess: the reference acceleration pushes apart.
    double A[NV][NV], b[NV], fc[NV];
    unsigned act = 0;
    unsigned long long dact = 0;
    int phase = (rows || nd) ? 0 : 2;
    bool done = false;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        fc[i] = 0.0;
        if ((rows >> i & 1) && rS[i] * f[i] < rA[i] * M[i][i]) act |= 1u << i;      // M_ii > 0
    }
    if (CONTACTS)
        for (int r = 0; r < nd; r++) if (Ad[r] > 0.0) dact |= 1ull << r;
#pragma unroll 1
    for (int it = 0; it < (CONTACTS ? 16 : 10) && !done; it++) {
#pragma unroll
        for (int i = 0; i < NV; i++) {
            const bool on = phase == 0 && (act >> i & 1);
            b[i] = f[i] + (phase == 2 ? fc[i] : 0.0) + (on ? rS[i] * rD[i] * rA[i] : 0.0);
#pragma unroll
            for (int j = 0; j < i; j++) A[i][j] = M[i][j];
            A[i][i] = M[i][i] + (phase == 2 ? h * lk[i * LK_STRIDE + LK_DAMP] : 0.0) + (on ? rD[i] : 0.0);
        }
        if (CONTACTS && phase == 0) {
#pragma unroll 1
            for (int r = 0; r < nd; r++) {
                if (!(dact >> r & 1)) continue;
                double jr[NV];
#pragma unroll
                for (int i = 0; i < NV; i++) jr[i] = Jd[r][i];
                const double dr = Dd[r], da = dr * Ad[r];
#pragma unroll
                for (int i = 0; i < NV; i++) {
                    const double di = dr * jr[i];
                    b[i] += da * jr[i];
#pragma unroll
                    for (int j = 0; j <= i; j++) A[i][j] += di * jr[j];
                }
            }
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
        unsigned long long nda = 0;
        if (CONTACTS) {
#pragma unroll 1
            for (int r = 0; r < nd; r++) {
                double sres = -Ad[r];
#pragma unroll
                for (int i = 0; i < NV; i++) sres += Jd[r][i] * b[i];
                if (sres < 0.0) nda |= 1ull << r;
            }
        }
        if (na == act && nda == dact) {
            if (!damped) { done = true; break; }
#pragma unroll
            for (int i = 0; i < NV; i++) fc[i] = (act >> i & 1) ? rS[i] * (-rD[i] * (rS[i] * b[i] - rA[i])) : 0.0;
            if (CONTACTS) {
#pragma unroll 1
                for (int r = 0; r < nd; r++) {
                    if (!(dact >> r & 1)) continue;
                    double sres = -Ad[r];
#pragma unroll
                    for (int i = 0; i < NV; i++) sres += Jd[r][i] * b[i];
#pragma unroll
                    for (int i = 0; i < NV; i++) fc[i] += Jd[r][i] * (-Dd[r] * sres);
                }
            }
            phase = 2;
        } else {
            act = na;
            dact = nda;
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
        if (CONTACTS && nd > 0) {
            dense_solve_integrate<NV>(lk, h, Ml, fl, damped, rD, rS, rA, nd, Jd, Dd, Ad, q, v);
            int nrc = 4 * ncontact;
#pragma unroll
            for (int i = 0; i < NV; i++) nrc += rows >> i & 1;
            return nrc;
        }
        return limits_solve_integrate<NV>(NV, lk, li, h, Ml, fl, damped, q, v);
    }
    int nr = 4 * ncontact;
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
