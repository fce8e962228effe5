// Articulated-body step for a serial chain of 7 axis-aligned hinge joints, one particle per
// thread, every loop unrolled at compile time so q, v, M and the recursion state live in
// registers.  Replaces, for the reacher_7dof model, what the reference reaches through
//   GymEnvWrapper.rollout          mjmpc/envs/gym_env_wrapper.py:125-153
//   Reacher7DOFEnv.step            mjmpc/envs/basic/reacher_env.py:29-39
//   -> mjrl do_simulation -> mujoco_py MjSim.step -> MuJoCo mj_step   (not in the reference tree)
//
// Formulation (differs on purpose from MuJoCo's / the oracle's world-frame one):
//   * link-frame recursive Newton-Euler for the Coriolis/centrifugal bias,
//   * link-frame composite-rigid-body recursion for M(q) (+ armature),
//   * soft joint limits / sphere-plane contact: exact minimiser of MuJoCo's convex
//     constraint objective by Newton steps on the active set (M + diag(D) + Dc Jc Jc'),
//   * semi-implicit Euler with implicit joint damping: (M + h B) a = f + f_constraint.
// Joint axes are template constants, rotations are 2x2 Givens updates, and structurally
// zero offsets / centre-of-mass components are skipped through the Traits masks.
#pragma once
#include <math.h>
#include <utility>
#include "chain_model.h"

#if defined(__CUDACC__)
#define MJB_HD __host__ __device__ __forceinline__
#else
#define MJB_HD inline
#endif

namespace mjb {

struct V3 { double x, y, z; };
struct S3 { double xx, yy, zz, xy, xz, yz; };

MJB_HD V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
MJB_HD V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
MJB_HD V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
MJB_HD V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
MJB_HD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
MJB_HD V3 mul(const S3& I, V3 w) {
    return {I.xx * w.x + I.xy * w.y + I.xz * w.z, I.xy * w.x + I.yy * w.y + I.yz * w.z,
            I.xz * w.x + I.yz * w.y + I.zz * w.z};
}
template <int AX> MJB_HD double comp(V3 v) { return AX == 0 ? v.x : (AX == 1 ? v.y : v.z); }

// v <- R v, R = rotation by angle (s = sin, c = cos) about coordinate axis AX (child -> parent frame)
template <int AX> MJB_HD V3 rot(V3 v, double s, double c) {
    if constexpr (AX == 0) return {v.x, c * v.y - s * v.z, s * v.y + c * v.z};
    else if constexpr (AX == 1) return {c * v.x + s * v.z, v.y, c * v.z - s * v.x};
    else return {c * v.x - s * v.y, s * v.x + c * v.y, v.z};
}
// v <- R' v (parent -> child frame)
template <int AX> MJB_HD V3 rotT(V3 v, double s, double c) { return rot<AX>(v, -s, c); }

// I <- R I R'
template <int AX> MJB_HD S3 rotS(const S3& I, double s, double c) {
    // 2x2 block [a b; b d] of the two rotated axes, and the coupling (p, r) with the fixed axis
    double a, b, d, p, r;
    if constexpr (AX == 0) { a = I.yy; b = I.yz; d = I.zz; p = I.xy; r = I.xz; }
    else if constexpr (AX == 1) { a = I.zz; b = I.xz; d = I.xx; p = I.yz; r = I.xy; }
    else { a = I.xx; b = I.xy; d = I.yy; p = I.xz; r = I.yz; }
    // rotation in the (u,w) plane: u' = c u - s w, w' = s u + c w
    const double t1 = c * a - s * b, t2 = c * b - s * d;   // row u of G*B
    const double t3 = s * a + c * b, t4 = s * b + c * d;   // row w of G*B
    const double na = c * t1 - s * t2, nb = s * t1 + c * t2, nd = s * t3 + c * t4;
    const double np = c * p - s * r, nr = s * p + c * r;
    S3 o;
    if constexpr (AX == 0) { o.xx = I.xx; o.yy = na; o.yz = nb; o.zz = nd; o.xy = np; o.xz = nr; }
    else if constexpr (AX == 1) { o.yy = I.yy; o.zz = na; o.xz = nb; o.xx = nd; o.yz = np; o.xy = nr; }
    else { o.zz = I.zz; o.xx = na; o.xy = nb; o.yy = nd; o.xz = np; o.yz = nr; }
    return o;
}

// compile-time loop
template <class F, int... Is> MJB_HD void static_for_impl(F&& f, std::integer_sequence<int, Is...>) {
    (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, class F> MJB_HD void static_for(F&& f) { static_for_impl(f, std::make_integer_sequence<int, N>{}); }

// Structure of the reference's arm (sawyer.xml:15-59): axes z,y,x,y,x,y,x; after merging the two
// welded bodies, link offsets are (x,0,0) or 0 and COMs are on the link x axis (or at the origin),
// so inertias about the link origin stay diagonal for links 1..6.  Bit k of a mask = component k.
struct SawyerTraits {
    static constexpr int axis(int i) { constexpr int a[7] = {2, 1, 0, 1, 0, 1, 0}; return a[i]; }
    static constexpr int off_mask(int i) { constexpr int m[7] = {7, 1, 0, 1, 0, 1, 0}; return m[i]; }
    static constexpr int com_mask(int i) { constexpr int m[7] = {7, 0, 1, 0, 1, 0, 1}; return m[i]; }
    static constexpr bool full_inertia(int i) { return i == 0; }
};
// Same axis pattern, no structural zeros assumed (any offsets / COMs / inertias).
struct DenseTraits {
    static constexpr int axis(int i) { constexpr int a[7] = {2, 1, 0, 1, 0, 1, 0}; return a[i]; }
    static constexpr int off_mask(int) { return 7; }
    static constexpr int com_mask(int) { return 7; }
    static constexpr bool full_inertia(int) { return true; }
};

template <class T, int I, class P> MJB_HD V3 link_offset(const P& prm) {
    constexpr int m = T::off_mask(I);
    return {(m & 1) ? prm[CH_OFF + 3 * I] : 0.0, (m & 2) ? prm[CH_OFF + 3 * I + 1] : 0.0,
            (m & 4) ? prm[CH_OFF + 3 * I + 2] : 0.0};
}
// o x f with masked offset components
template <class T, int I, class P> MJB_HD V3 off_cross(const P& prm, V3 f) {
    constexpr int m = T::off_mask(I);
    V3 r = {0.0, 0.0, 0.0};
    if constexpr (m & 1) { const double ox = prm[CH_OFF + 3 * I]; r.y -= ox * f.z; r.z += ox * f.y; }
    if constexpr (m & 2) { const double oy = prm[CH_OFF + 3 * I + 1]; r.x += oy * f.z; r.z -= oy * f.x; }
    if constexpr (m & 4) { const double oz = prm[CH_OFF + 3 * I + 2]; r.x -= oz * f.y; r.y += oz * f.x; }
    return r;
}
template <class T, int I, class P> MJB_HD V3 link_h(const P& prm) {
    constexpr int m = T::com_mask(I);
    return {(m & 1) ? prm[CH_H + 3 * I] : 0.0, (m & 2) ? prm[CH_H + 3 * I + 1] : 0.0,
            (m & 4) ? prm[CH_H + 3 * I + 2] : 0.0};
}
template <class T, int I, class P> MJB_HD S3 link_inertia_o(const P& prm) {
    S3 s;
    s.xx = prm[CH_IO + 6 * I]; s.yy = prm[CH_IO + 6 * I + 1]; s.zz = prm[CH_IO + 6 * I + 2];
    if constexpr (T::full_inertia(I)) { s.xy = prm[CH_IO + 6 * I + 3]; s.xz = prm[CH_IO + 6 * I + 4]; s.yz = prm[CH_IO + 6 * I + 5]; }
    else { s.xy = 0.0; s.xz = 0.0; s.yz = 0.0; }
    return s;
}
// a x h and w x (w x h) with masked h
template <class T, int I> MJB_HD V3 cross_h(V3 a, V3 h) {
    constexpr int m = T::com_mask(I);
    V3 r = {0.0, 0.0, 0.0};
    if constexpr (m & 1) { r.y += a.z * h.x; r.z -= a.y * h.x; }
    if constexpr (m & 2) { r.x -= a.z * h.y; r.z += a.x * h.y; }
    if constexpr (m & 4) { r.x += a.y * h.z; r.y -= a.x * h.z; }
    return r;
}

// ---------------------------------------------------------------------------------------------
// M(q) lower triangle (with armature) and bias(q,v).  sn/cs = sin/cos of the joint angles.
// ---------------------------------------------------------------------------------------------
template <class T, class P>
MJB_HD void chain_mass_bias(const P& prm, const double (&sn)[7], const double (&cs)[7], const double (&qd)[7],
                            double (&M)[7][7], double (&bias)[7]) {
    // ---- recursive Newton-Euler, outward pass: link velocities / bias accelerations -> link wrenches
    V3 lf[7], ln[7];
    {
        V3 w = {0, 0, 0}, al = {0, 0, 0}, ac = {0, 0, 0};
        static_for<7>([&](auto Ic) {
            constexpr int i = decltype(Ic)::value;
            constexpr int AX = T::axis(i);
            if constexpr (i == 0) {
                // fixed base: only the joint's own rate
                w = {AX == 0 ? qd[0] : 0.0, AX == 1 ? qd[0] : 0.0, AX == 2 ? qd[0] : 0.0};
            } else {
                const V3 o = link_offset<T, i>(prm);
                // acceleration of this joint's origin, parent frame: a + al x o + w x (w x o)
                V3 ao = ac;
                if constexpr (T::off_mask(i) != 0) {
                    const V3 wxo = cross(w, o);
                    ao = ao + cross(al, o) + cross(w, wxo);
                }
                const V3 wp = rotT<AX>(w, sn[i], cs[i]);
                const V3 alp = rotT<AX>(al, sn[i], cs[i]);
                ac = rotT<AX>(ao, sn[i], cs[i]);
                const double r = qd[i];
                // al = alp + (wp x e) r ; w = wp + e r
                if constexpr (AX == 0) { al = {alp.x, alp.y + wp.z * r, alp.z - wp.y * r}; w = {wp.x + r, wp.y, wp.z}; }
                else if constexpr (AX == 1) { al = {alp.x - wp.z * r, alp.y, alp.z + wp.x * r}; w = {wp.x, wp.y + r, wp.z}; }
                else { al = {alp.x + wp.y * r, alp.y - wp.x * r, alp.z}; w = {wp.x, wp.y, wp.z + r}; }
            }
            // wrench about the link origin: f = m a + al x h + w x (w x h); n = I al + w x (I w) + h x a
            const V3 h = link_h<T, i>(prm);
            const S3 Io = link_inertia_o<T, i>(prm);
            const double m = prm[CH_MASS + i];
            V3 f = m * ac;
            V3 n = mul(Io, al) + cross(w, mul(Io, w));
            if constexpr (T::com_mask(i) != 0) {
                const V3 wxh = cross_h<T, i>(w, h);
                f = f + cross_h<T, i>(al, h) + cross(w, wxh);
                n = n - cross_h<T, i>(ac, h);   // h x a = -(a x h)
            }
            lf[i] = f; ln[i] = n;
        });
    }
    // ---- inward pass: accumulate child wrenches, project on the joint axis
    {
        V3 fa = {0, 0, 0}, na = {0, 0, 0};
        static_for<7>([&](auto Ic) {
            constexpr int i = 6 - decltype(Ic)::value;
            constexpr int AX = T::axis(i);
            fa = fa + lf[i];
            na = na + ln[i];
            bias[i] = comp<AX>(na);
            if constexpr (i > 0) {
                fa = rot<AX>(fa, sn[i], cs[i]);
                na = rot<AX>(na, sn[i], cs[i]) + off_cross<T, i>(prm, fa);
            }
        });
    }
    // ---- composite rigid body recursion (inward): composite (mass, first moment, inertia about origin)
    {
        double cm = 0.0;
        V3 ch = {0, 0, 0};
        S3 cI = {0, 0, 0, 0, 0, 0};
        static_for<7>([&](auto Ic) {
            constexpr int i = 6 - decltype(Ic)::value;
            constexpr int AX = T::axis(i);
            // add link i's own inertia (composite is expressed in frame i here)
            {
                const V3 h = link_h<T, i>(prm);
                const S3 Io = link_inertia_o<T, i>(prm);
                cm += prm[CH_MASS + i];
                ch = ch + h;
                cI.xx += Io.xx; cI.yy += Io.yy; cI.zz += Io.zz; cI.xy += Io.xy; cI.xz += Io.xz; cI.yz += Io.yz;
            }
            // unit acceleration about the joint axis: F = e x ch, N = cI e
            V3 F, N;
            if constexpr (AX == 0) { F = {0.0, -ch.z, ch.y}; N = {cI.xx, cI.xy, cI.xz}; }
            else if constexpr (AX == 1) { F = {ch.z, 0.0, -ch.x}; N = {cI.xy, cI.yy, cI.yz}; }
            else { F = {-ch.y, ch.x, 0.0}; N = {cI.xz, cI.yz, cI.zz}; }
            M[i][i] = comp<AX>(N) + prm[CH_ARMATURE + i];
            // carry the wrench down to every ancestor joint
            static_for<i>([&](auto Jc) {
                constexpr int j = i - decltype(Jc)::value;      // hop through joint j: frame j -> j-1
                constexpr int AJ = T::axis(j);
                F = rot<AJ>(F, sn[j], cs[j]);
                N = rot<AJ>(N, sn[j], cs[j]) + off_cross<T, j>(prm, F);
                M[i][j - 1] = comp<T::axis(j - 1)>(N);
            });
            // move the composite to the parent frame
            if constexpr (i > 0) {
                const V3 hr = rot<AX>(ch, sn[i], cs[i]);
                cI = rotS<AX>(cI, sn[i], cs[i]);
                if constexpr (T::off_mask(i) != 0) {
                    const V3 o = link_offset<T, i>(prm);
                    // reference point moves by -o: I += (2 h.o + m o.o) 1 - (h o' + o h') - m o o'
                    const V3 mo = cm * o;
                    const V3 g = hr + hr + mo;                  // 2h + m o
                    const double tr = dot(g, o);
                    cI.xx += tr - g.x * o.x; cI.yy += tr - g.y * o.y; cI.zz += tr - g.z * o.z;
                    // off-diagonals: -(h_a o_b + o_a h_b) - m o_a o_b
                    cI.xy -= hr.x * o.y + o.x * hr.y + mo.x * o.y;
                    cI.xz -= hr.x * o.z + o.x * hr.z + mo.x * o.z;
                    cI.yz -= hr.y * o.z + o.y * hr.z + mo.y * o.z;
                    ch = hr + mo;
                } else {
                    ch = hr;
                }
            }
        });
    }
}

// In-place LDL' of the lower triangle of a 7x7 SPD matrix: A[i][j] (i>j) <- L[i][j], dinv[j] = 1/D[j].
MJB_HD void ldl7(double (&A)[7][7], double (&dinv)[7]) {
#pragma unroll
    for (int j = 0; j < 7; j++) {
        double w[7];
        double dj = A[j][j];
#pragma unroll
        for (int k = 0; k < j; k++) { w[k] = A[j][k] * A[k][k]; dj -= A[j][k] * w[k]; }
        A[j][j] = dj;
        const double inv = 1.0 / dj;
        dinv[j] = inv;
#pragma unroll
        for (int i = j + 1; i < 7; i++) {
            double t = A[i][j];
#pragma unroll
            for (int k = 0; k < j; k++) t -= A[i][k] * w[k];
            A[i][j] = t * inv;
        }
    }
}
MJB_HD void ldl7_solve(const double (&L)[7][7], const double (&dinv)[7], double (&b)[7]) {
#pragma unroll
    for (int i = 1; i < 7; i++) {
#pragma unroll
        for (int k = 0; k < i; k++) b[i] -= L[i][k] * b[k];
    }
#pragma unroll
    for (int i = 0; i < 7; i++) b[i] *= dinv[i];
#pragma unroll
    for (int i = 5; i >= 0; i--) {
#pragma unroll
        for (int k = i + 1; k < 7; k++) b[i] -= L[k][i] * b[k];
    }
}

// MuJoCo's constraint impedance as a function of penetration (solimp d0,dwidth,width,midpoint,power)
template <class P> MJB_HD double impedance(const P& prm, double dist_minus_margin) {
    const double d0 = prm[CS_IMP_D0], dw = prm[CS_IMP_DW], width = prm[CS_IMP_WIDTH];
    if (d0 == dw || width <= 1e-15) return 0.5 * (d0 + dw);
    const double x = fabs(dist_minus_margin / width);
    if (x >= 1.0) return dw;
    if (x <= 0.0) return d0;
    const double mid = prm[CS_IMP_MID], pw = prm[CS_IMP_POWER];
    double y;
    if (pw == 1.0) y = x;
    else if (x <= mid) y = pow(x, pw) / pow(mid, pw - 1.0);
    else y = 1.0 - pow(1.0 - x, pw) / pow(1.0 - mid, pw - 1.0);
    return d0 + y * (dw - d0);
}

// World position of a point given in the last link's frame (nested evaluation, no matrices).
template <class T, class P>
MJB_HD V3 chain_point_world(const P& prm, const double (&sn)[7], const double (&cs)[7], V3 p) {
    static_for<7>([&](auto Ic) {
        constexpr int i = 6 - decltype(Ic)::value;
        p = rot<T::axis(i)>(p, sn[i], cs[i]);
        if constexpr (T::off_mask(i) != 0) p = p + link_offset<T, i>(prm);
    });
    return p;
}

// Constraint rows of one particle: 7 joint-limit rows (at most one side of a joint can be violated;
// row j acts on dof j with Jacobian entry sg[j] = -side) and one frictionless contact row.
struct Rows {
    double D[7], aref[7], sg[7];   // D == 0: row absent
    double Dc, arefc, Jc[7];       // Dc == 0: no contact
    bool any;
};

template <class T, class P>
MJB_HD void make_rows(const P& prm, const double (&q)[7], const double (&qd)[7], const double (&sn)[7],
                      const double (&cs)[7], Rows& R) {
    R.any = false;
    const double K = prm[CS_SOLK], B = prm[CS_SOLB];
    const int limited = (int)prm[CS_LIMITED_MASK];
#pragma unroll
    for (int j = 0; j < 7; j++) {
        const double lo = prm[CH_RANGE_LO + j], hi = prm[CH_RANGE_HI + j];
        double dist = 0.0, side = 0.0;
        if (q[j] < lo) { dist = q[j] - lo; side = -1.0; }          // side * (range - q), side = -1
        else if (q[j] > hi) { dist = hi - q[j]; side = 1.0; }
        R.D[j] = 0.0; R.aref[j] = 0.0; R.sg[j] = -side;
        if (dist < 0.0 && ((limited >> j) & 1)) {
            const double imp = impedance(prm, dist);
            double r = (1.0 - imp) * prm[CH_INVW0 + j] / imp;
            r = r < 1e-15 ? 1e-15 : r;
            R.D[j] = 1.0 / r;
            R.aref[j] = -B * (-side * qd[j]) - K * imp * dist;
            R.any = true;
        }
    }
    R.Dc = 0.0; R.arefc = 0.0;
#pragma unroll
    for (int j = 0; j < 7; j++) R.Jc[j] = 0.0;
    const double radius = prm[CS_CON_RADIUS];
    if (radius > 0.0) {
        const V3 c = chain_point_world<T>(prm, sn, cs, V3{prm[CS_CON_POS], prm[CS_CON_POS + 1], prm[CS_CON_POS + 2]});
        const double margin = prm[CS_CON_MARGIN];
        const double dist = c.z - prm[CS_CON_PLANE_Z] - radius;
        if (dist < margin) {
            // contact point (mid-surface) and plane normal carried into each link frame:
            // Jc[j] = n_j . (e_j x r_j), r_j = contact point relative to joint j, link-j coordinates
            V3 n = {0.0, 0.0, 1.0};
            V3 r = {c.x, c.y, c.z - (radius + 0.5 * dist)};
            static_for<7>([&](auto Ic) {
                constexpr int i = decltype(Ic)::value;
                constexpr int AX = T::axis(i);
                r = r - link_offset<T, i>(prm);
                r = rotT<AX>(r, sn[i], cs[i]);
                n = rotT<AX>(n, sn[i], cs[i]);
                // n . (e x r)
                if constexpr (AX == 0) R.Jc[i] = n.z * r.y - n.y * r.z;
                else if constexpr (AX == 1) R.Jc[i] = n.x * r.z - n.z * r.x;
                else R.Jc[i] = n.y * r.x - n.x * r.y;
            });
            const double imp = impedance(prm, dist - margin);
            double rr = (1.0 - imp) * prm[CS_CON_INVW] / imp;
            rr = rr < 1e-15 ? 1e-15 : rr;
            R.Dc = 1.0 / rr;
            double vel = 0.0;
#pragma unroll
            for (int j = 0; j < 7; j++) vel += R.Jc[j] * qd[j];
            R.arefc = -B * vel - K * imp * (dist - margin);
            R.any = true;
        }
    }
}

// Exact minimiser over a of  1/2 a'Ma - f'a + sum_r 1/2 D_r min(0, J_r a - aref_r)^2 ; returns the
// constraint force J' lambda in fc.  Newton on the active set with an exact line search (each
// piece of the objective is quadratic, so a step that keeps its active set lands on the optimum).
MJB_HD void solve_constraints(const double (&M)[7][7], const double (&f)[7], const Rows& R, double (&fc)[7]) {
    double a[7];
#pragma unroll
    for (int j = 0; j < 7; j++) a[j] = 0.0;
    unsigned act = 0;            // bit j: limit row j active, bit 7: contact
    for (int it = 0; it < 64; it++) {
        // active set at the current point
        double jc = -R.arefc;
#pragma unroll
        for (int j = 0; j < 7; j++) jc += R.Jc[j] * a[j];
        act = 0;
#pragma unroll
        for (int j = 0; j < 7; j++) if (R.D[j] > 0.0 && R.sg[j] * a[j] - R.aref[j] < 0.0) act |= 1u << j;
        if (R.Dc > 0.0 && jc < 0.0) act |= 128u;
        // minimiser of the current quadratic piece
        double H[7][7], dinv[7], x[7];
        const double dc = (act & 128u) ? R.Dc : 0.0;
#pragma unroll
        for (int i = 0; i < 7; i++) {
            const double di = ((act >> i) & 1u) ? R.D[i] : 0.0;
#pragma unroll
            for (int j = 0; j <= i; j++) H[i][j] = M[i][j] + dc * R.Jc[i] * R.Jc[j];
            H[i][i] += di;
            x[i] = f[i] + di * R.aref[i] * R.sg[i] + dc * R.arefc * R.Jc[i];
        }
        ldl7(H, dinv);
        ldl7_solve(H, dinv, x);
        // active set at the candidate
        double jcx = -R.arefc;
#pragma unroll
        for (int j = 0; j < 7; j++) jcx += R.Jc[j] * x[j];
        unsigned actx = 0;
#pragma unroll
        for (int j = 0; j < 7; j++) if (R.D[j] > 0.0 && R.sg[j] * x[j] - R.aref[j] < 0.0) actx |= 1u << j;
        if (R.Dc > 0.0 && jcx < 0.0) actx |= 128u;
        if (actx == act) {
#pragma unroll
            for (int j = 0; j < 7; j++) a[j] = x[j];
            break;
        }
        // exact line search on phi(t) = obj(a + t p), p = x - a:  phi'(t) piecewise linear, increasing
        double p[7], g0 = 0.0, h0 = 0.0;
#pragma unroll
        for (int j = 0; j < 7; j++) p[j] = x[j] - a[j];
#pragma unroll
        for (int i = 0; i < 7; i++) {
            double s = 0.0, ga = -f[i];
#pragma unroll
            for (int k = 0; k < 7; k++) {
                const double mik = k <= i ? M[i][k] : M[k][i];
                s += mik * p[k]; ga += mik * a[k];
            }
            g0 += p[i] * ga; h0 += p[i] * s;
        }
        double jar[8], jp[8], Dr[8];
#pragma unroll
        for (int j = 0; j < 7; j++) { jar[j] = R.sg[j] * a[j] - R.aref[j]; jp[j] = R.sg[j] * p[j]; Dr[j] = R.D[j]; }
        jar[7] = jc; jp[7] = jcx - jc; Dr[7] = R.Dc;
        // bracket the root between consecutive breakpoints: lo = largest breakpoint with phi' <= 0
        double lo = 0.0, hi = INFINITY;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            if (Dr[r] > 0.0 && jp[r] != 0.0) {
                const double t = -jar[r] / jp[r];
                if (t > 0.0) {
                    double d = g0 + t * h0;
#pragma unroll
                    for (int s = 0; s < 8; s++) {
                        const double js = jar[s] + t * jp[s];
                        if (Dr[s] > 0.0 && js < 0.0) d += Dr[s] * js * jp[s];
                    }
                    if (d <= 0.0) { if (t > lo) lo = t; }
                    else if (t < hi) hi = t;
                }
            }
        }
        const double mid = (hi == INFINITY) ? lo + 1.0 : 0.5 * (lo + hi);
        double c0 = g0, c1 = h0;
#pragma unroll
        for (int s = 0; s < 8; s++) {
            if (Dr[s] > 0.0 && jar[s] + mid * jp[s] < 0.0) { c0 += Dr[s] * jar[s] * jp[s]; c1 += Dr[s] * jp[s] * jp[s]; }
        }
        double t = -c0 / c1;
        t = t < lo ? lo : (t > hi ? hi : t);
#pragma unroll
        for (int j = 0; j < 7; j++) a[j] += t * p[j];
    }
    // constraint force at the optimum: lambda_r = -D_r (J_r a - aref_r) on active rows
    double jc = -R.arefc;
#pragma unroll
    for (int j = 0; j < 7; j++) jc += R.Jc[j] * a[j];
    const double lc = (R.Dc > 0.0 && jc < 0.0) ? -R.Dc * jc : 0.0;
#pragma unroll
    for (int j = 0; j < 7; j++) {
        const double jr = R.sg[j] * a[j] - R.aref[j];
        const double lam = (R.D[j] > 0.0 && jr < 0.0) ? -R.D[j] * jr : 0.0;
        fc[j] = R.sg[j] * lam + lc * R.Jc[j];
    }
}

// One mj_step of the chain: forward dynamics at (q, qd) under control u, then Euler advance.
// sn/cs must hold sin/cos of q on entry.  Returns true when a constraint row was present.
template <class T, class P>
MJB_HD bool chain_substep(const P& prm, double (&q)[7], double (&qd)[7], const double (&sn)[7],
                          const double (&cs)[7], const double (&u)[7]) {
    double M[7][7], f[7];
    chain_mass_bias<T>(prm, sn, cs, qd, M, f);
#pragma unroll
    for (int j = 0; j < 7; j++) {
        double c = u[j];
        c = c < prm[CH_CTRL_LO + j] ? prm[CH_CTRL_LO + j] : c;
        c = c > prm[CH_CTRL_HI + j] ? prm[CH_CTRL_HI + j] : c;
        f[j] = prm[CH_GEAR + j] * c - prm[CH_DAMPING + j] * qd[j] - f[j];
    }
    Rows R;
    make_rows<T>(prm, q, qd, sn, cs, R);
    if (R.any) {
        double fc[7];
        solve_constraints(M, f, R, fc);
#pragma unroll
        for (int j = 0; j < 7; j++) f[j] += fc[j];
    }
    double dinv[7];
#pragma unroll
    for (int j = 0; j < 7; j++) M[j][j] += prm[CH_HDAMP + j];
    ldl7(M, dinv);
    ldl7_solve(M, dinv, f);
    const double h = prm[CS_TIMESTEP];
#pragma unroll
    for (int j = 0; j < 7; j++) { qd[j] += h * f[j]; q[j] += h * qd[j]; }
    return R.any;
}

// The reference's step cost (reacher_env.py:31-35): -reward = |h-g|_1 + 5 |h-g|_2
MJB_HD double reach_cost(V3 hand, V3 target) {
    const double dx = hand.x - target.x, dy = hand.y - target.y, dz = hand.z - target.z;
    return (fabs(dx) + fabs(dy) + fabs(dz)) + 5.0 * sqrt(dx * dx + dy * dy + dz * dz);
}

}  // namespace mjb
