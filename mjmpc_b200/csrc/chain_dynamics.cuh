// Articulated-body step for a serial chain of 7 axis-aligned hinge joints, one particle per
// thread, the kinematic recursions unrolled at compile time so q, v, M and the recursion state
// live in registers.  Replaces, for the reacher_7dof model, what the reference reaches through
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
// Joint axes are template constants, rotations are 2x2 Givens updates, and structurally zero
// offsets / centre-of-mass components / inertia products are skipped through the Traits masks.
//
// Code size matters as much as FLOPs here: with 255 registers per thread only 8 warps share an
// SM, so the loop body has to stay inside the instruction cache (the first version of this file
// inlined 140 KB of SASS and spent 7 of every 8 issue slots waiting for instructions).  Hence one
// shared factor/solve site for the Newton and Euler solves, a rolled line search, and a compact
// sincos for the bounded joint angles.
#pragma once
#include <math.h>
#include <utility>
#include "chain_model.h"

// hide a value from constant propagation (keeps the compiler from cloning the factor/solve loop body
// per phase: one copy of that code is the point)
#if defined(__CUDA_ARCH__)
#define MJB_OPAQUE(x) asm volatile("" : "+r"(x))
#else
#define MJB_OPAQUE(x) ((void)0)
#endif

#if defined(__CUDACC__)
#define MJB_HD __host__ __device__ __forceinline__
#define MJB_NOINLINE static __host__ __device__ __noinline__
#else
#define MJB_HD inline
#define MJB_NOINLINE __attribute__((noinline)) inline
#endif

namespace mjb {

#if defined(MJB_HOST_STATS)
// host test harness only: solver statistics [substeps, substeps with rows, factor/solve passes, line searches,
// rank-one repairs tried, repairs confirmed]
static long long g_stats[6] = {0, 0, 0, 0, 0, 0};
static int* g_trips = nullptr;      // optional: factor/solve passes of every substep, in call order
static long long g_ntrips = 0;
#define MJB_STAT(i) (g_stats[i]++)
#define MJB_STAT_TRIPS(n) do { if (g_trips) g_trips[g_ntrips++] = (n); } while (0)
#else
#define MJB_STAT(i) ((void)0)
#define MJB_STAT_TRIPS(n) ((void)0)
#endif

struct V3 { double x, y, z; };
struct S3 { double xx, yy, zz, xy, xz, yz; };

MJB_HD V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
MJB_HD V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
MJB_HD V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
MJB_HD V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
// min / max without the NaN-propagation fix-ups of fmin/fmax (9 instructions each in FP64 on sm_100a)
MJB_HD double dmin(double a, double b) { return a < b ? a : b; }
MJB_HD double dmax(double a, double b) { return a > b ? a : b; }
// v with its sign bit XOR-ed with `mask` (0 or 0x80000000): multiplication by +-1 as one integer
// instruction on the high word
MJB_HD double sflip(double v, unsigned mask) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(v) ^ (int)mask, __double2loint(v));
#else
    return mask ? -v : v;
#endif
}
template <int AX> MJB_HD double comp(V3 v) { return AX == 0 ? v.x : (AX == 1 ? v.y : v.z); }

// a x b where b's components outside MASK are structurally zero (bit k = component k)
template <int MASK> MJB_HD V3 cross_m(V3 a, V3 b) {
    V3 r = {0.0, 0.0, 0.0};
    if constexpr (MASK & 1) { r.y += a.z * b.x; r.z -= a.y * b.x; }
    if constexpr (MASK & 2) { r.x -= a.z * b.y; r.z += a.x * b.y; }
    if constexpr (MASK & 4) { r.x += a.y * b.z; r.y -= a.x * b.z; }
    return r;
}
// acc + a x b  and  acc - a x b  as two FMAs per component (a bare cross product followed by a vector
// add costs three FP64 operations per component)
MJB_HD V3 cross_acc(V3 acc, V3 a, V3 b) {
    return {fma(a.y, b.z, fma(-a.z, b.y, acc.x)), fma(a.z, b.x, fma(-a.x, b.z, acc.y)), fma(a.x, b.y, fma(-a.y, b.x, acc.z))};
}
// acc + SIGN * (a x b), b's components outside MASK structurally zero
template <int MASK, int SIGN> MJB_HD V3 cross_m_acc(V3 acc, V3 a, V3 b) {
    const double sx = SIGN > 0 ? b.x : -b.x, sy = SIGN > 0 ? b.y : -b.y, sz = SIGN > 0 ? b.z : -b.z;
    if constexpr (MASK & 1) { acc.y = fma(a.z, sx, acc.y); acc.z = fma(-a.y, sx, acc.z); }
    if constexpr (MASK & 2) { acc.x = fma(-a.z, sy, acc.x); acc.z = fma(a.x, sy, acc.z); }
    if constexpr (MASK & 4) { acc.x = fma(a.y, sz, acc.x); acc.y = fma(-a.x, sz, acc.y); }
    return acc;
}
// I w for a symmetric 3x3; FULL = false: diagonal only
template <bool FULL> MJB_HD V3 mul_s(const S3& I, V3 w) {
    if constexpr (FULL)
        return {I.xx * w.x + I.xy * w.y + I.xz * w.z, I.xy * w.x + I.yy * w.y + I.yz * w.z,
                I.xz * w.x + I.yz * w.y + I.zz * w.z};
    else
        return {I.xx * w.x, I.yy * w.y, I.zz * w.z};
}

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

// sin and cos of a joint angle.  Joint angles are bounded by the (soft) joint limits, so a two-term
// Cody-Waite reduction by pi/2 and the fdlibm kernel polynomials on [-pi/4, pi/4] are accurate to
// ~1 ulp for |x| < 1e5 rad; there is no huge-argument path.
MJB_HD void sincos_joint(double x, double& s, double& c) {
    const double kd = rint(x * 6.36619772367581382433e-01);          // x * 2/pi
    double r = fma(kd, -1.57079632673412561417e+00, x);              // pi/2, first 33 bits
    r = fma(kd, -6.07710050650619224932e-11, r);                     // pi/2 tail
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
    s = (k & 2) ? -a : a;
    c = ((k + 1) & 2) ? -b : b;
}

// sin / cos of q + dq from sin / cos of q by the angle-addition formulas, sin(dq) and cos(dq) - 1 from their Taylor
// series (to dq^13 and dq^14: truncation below 2e-17 for |dq| <= 0.25 rad, one integration step of a joint slower
// than 25 rad/s).  21 FP64 instructions instead of the 41 of sincos_joint and no range reduction; each update
// rounds s and c once more, so the pair drifts from the exact values by O(updates x 1e-16) -- the role-split
// rollout uses it (64 updates per horizon: < 1e-14), callers fall back to sincos_joint beyond 0.25 rad.
MJB_HD void sincos_advance(double dq, double& s, double& c) {
    const double z = dq * dq;
    const double S = z * fma(z, fma(z, fma(z, fma(z, fma(z, 1.60590438368216146e-10, -2.50521083854417188e-08),
                                                   2.75573192239858907e-06), -1.98412698412698413e-04),
                                    8.33333333333333333e-03), -1.66666666666666667e-01);          // sin(dq)/dq - 1
    const double C = z * fma(z, fma(z, fma(z, fma(z, fma(z, fma(z, -1.14707455977297247e-11, 2.08767569878680990e-09),
                                                          -2.75573192239858907e-07), 2.48015873015873016e-05),
                                           -1.38888888888888889e-03), 4.16666666666666667e-02), -0.5);   // cos(dq) - 1
    const double sd = fma(dq, S, dq);
    const double s0 = s;
    s = s0 + fma(c, sd, s0 * C);
    c = c + fma(-s0, sd, c * C);
}
// exact re-evaluation of all seven pairs, out of line (rare path of the callers of sincos_advance):
// io[0..7) angles in, io[7..14) sin, io[14..21) cos out
MJB_NOINLINE void sincos_all(double* io) {
#pragma unroll 1
    for (int j = 0; j < 7; j++) {
        double s, c;
        sincos_joint(io[j], s, c);
        io[7 + j] = s; io[14 + j] = c;
    }
}

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
// p + o with masked o
template <int MASK> MJB_HD V3 add_m(V3 p, V3 o) {
    if constexpr (MASK & 1) p.x += o.x;
    if constexpr (MASK & 2) p.y += o.y;
    if constexpr (MASK & 4) p.z += o.z;
    return p;
}

// Per-particle scratch outside the register file.  Values with a long life and a single use (the
// link wrenches between the two Newton-Euler passes, M between its construction and the solves, the
// contact Jacobian, the control held over both substeps) are parked here explicitly instead of
// being spilled by the compiler: on the device a slot is one column of shared memory
// (slot * BLOCK + thread: conflict-free), on the host a plain array.
// SC_M2 (M + Dc Jc Jc', only while a contact row exists) reuses the link-wrench slots, which are dead once
// the bias forces are known.
// SC_JC (contact Jacobian, only while a contact row exists) sits in the tail of the link-wrench region
// that M2 leaves free.  SC_NZ: landing zone of the asynchronous copy of the NEXT env step's noise row
// (kernel wrapper only; compiled out with MJB_NO_PREFETCH to fit one more block per SM).
#ifdef MJB_NO_PREFETCH
enum { SC_LF = 0, SC_LN = 18, SC_M = 36, SC_JC = 28, SC_U = 64, SC_NZ = 71, SC_NSLOT = 71, SC_M2 = 0 };
#else
enum { SC_LF = 0, SC_LN = 18, SC_M = 36, SC_JC = 28, SC_U = 64, SC_NZ = 71, SC_NSLOT = 78, SC_M2 = 0 };
#endif
// fused in-kernel noise: filtered history eps[t-1] lives in the SC_NZ slots, eps[t-2] in SC_E2
enum { SC_E2 = SC_NSLOT, SC_NSLOT_FUSED = SC_NSLOT + 7 };
MJB_HD constexpr int sc_m(int i, int j) { return SC_M + i * (i + 1) / 2 + j; }   // lower triangle, i >= j

struct HostScratch {
    double buf[SC_NSLOT];
    double ld(int slot) const { return buf[slot]; }
    void st(int slot, double v) { buf[slot] = v; }
};

// ---------------------------------------------------------------------------------------------
// M(q) lower triangle (with armature) -> scratch, and bias(q,v).  sn/cs = sin/cos of the joint angles.
// PARTS: bit 0 = the bias forces (the two Newton-Euler passes), bit 1 = the mass matrix (composite-rigid-body
// recursion); the two halves are independent of each other (role-split experiments compile them apart).
// ---------------------------------------------------------------------------------------------
template <class T, int PARTS = 3, class P, class S>
MJB_HD void chain_mass_bias(const P& prm, S& sc, const double (&sn)[7], const double (&cs)[7], const double (&qd)[7],
                            double (&bias)[7]) {
    // ---- recursive Newton-Euler, outward pass: link velocities / bias accelerations -> link wrenches.
    // Link 0 hangs off the fixed base: its own wrench never reaches a joint axis (its z torque is
    // zero and it has no parent), so only its angular velocity is carried on.
    if constexpr (PARTS & 1) {
        V3 w = {0, 0, 0}, al = {0, 0, 0}, ac = {0, 0, 0};
        static_for<7>([&](auto Ic) {
            constexpr int i = decltype(Ic)::value;
            constexpr int AX = T::axis(i);
            constexpr int OM = T::off_mask(i), CM = T::com_mask(i);
            if constexpr (i == 0) {
                w = {AX == 0 ? qd[0] : 0.0, AX == 1 ? qd[0] : 0.0, AX == 2 ? qd[0] : 0.0};
            } else {
                // acceleration of this joint's origin, parent frame: a + al x o + w x (w x o)
                V3 ao = ac;
                if constexpr (OM != 0) {
                    const V3 o = link_offset<T, i>(prm);
                    const V3 wxo = cross_m<OM>(w, o);
                    ao = cross_acc(cross_m_acc<OM, 1>(ac, al, o), w, wxo);
                }
                const V3 wp = rotT<AX>(w, sn[i], cs[i]);
                const V3 alp = rotT<AX>(al, sn[i], cs[i]);
                ac = rotT<AX>(ao, sn[i], cs[i]);
                const double r = qd[i];
                // al = alp + (wp x e) r ; w = wp + e r
                if constexpr (AX == 0) { al = {alp.x, alp.y + wp.z * r, alp.z - wp.y * r}; w = {wp.x + r, wp.y, wp.z}; }
                else if constexpr (AX == 1) { al = {alp.x - wp.z * r, alp.y, alp.z + wp.x * r}; w = {wp.x, wp.y + r, wp.z}; }
                else { al = {alp.x + wp.y * r, alp.y - wp.x * r, alp.z}; w = {wp.x, wp.y, wp.z + r}; }
                // wrench about the link origin: f = m a + al x h + w x (w x h); n = I al + w x (I w) + h x a
                const S3 Io = link_inertia_o<T, i>(prm);
                const double m = prm[CH_MASS + i];
                V3 f = m * ac;
                V3 n = cross_acc(mul_s<T::full_inertia(i)>(Io, al), w, mul_s<T::full_inertia(i)>(Io, w));
                if constexpr (CM != 0) {
                    const V3 h = link_h<T, i>(prm);
                    const V3 wxh = cross_m<CM>(w, h);
                    f = cross_acc(cross_m_acc<CM, 1>(f, al, h), w, wxh);
                    n = cross_m_acc<CM, -1>(n, ac, h);   // h x a = -(a x h)
                }
                sc.st(SC_LF + 3 * (i - 1), f.x); sc.st(SC_LF + 3 * (i - 1) + 1, f.y); sc.st(SC_LF + 3 * (i - 1) + 2, f.z);
                sc.st(SC_LN + 3 * (i - 1), n.x); sc.st(SC_LN + 3 * (i - 1) + 1, n.y); sc.st(SC_LN + 3 * (i - 1) + 2, n.z);
            }
        });
    }
    // ---- inward pass: accumulate child wrenches, project on the joint axis
    if constexpr (PARTS & 1) {
        V3 fa = {0, 0, 0}, na = {0, 0, 0};
        static_for<7>([&](auto Ic) {
            constexpr int i = 6 - decltype(Ic)::value;
            constexpr int AX = T::axis(i);
            if constexpr (i > 0) {
                fa = fa + V3{sc.ld(SC_LF + 3 * (i - 1)), sc.ld(SC_LF + 3 * (i - 1) + 1), sc.ld(SC_LF + 3 * (i - 1) + 2)};
                na = na + V3{sc.ld(SC_LN + 3 * (i - 1)), sc.ld(SC_LN + 3 * (i - 1) + 1), sc.ld(SC_LN + 3 * (i - 1) + 2)};
            }
            bias[i] = comp<AX>(na);
            if constexpr (i > 0) {
                fa = rot<AX>(fa, sn[i], cs[i]);
                na = rot<AX>(na, sn[i], cs[i]);
                if constexpr (T::off_mask(i) != 0) na = cross_m_acc<T::off_mask(i), -1>(na, fa, link_offset<T, i>(prm));   // + o x f
            }
        });
    }
    // ---- composite rigid body recursion (inward): composite (mass, first moment, inertia about origin)
    if constexpr (PARTS & 2) {
        double cm = 0.0;
        V3 ch = {0, 0, 0};
        S3 cI = {0, 0, 0, 0, 0, 0};
        static_for<7>([&](auto Ic) {
            constexpr int i = 6 - decltype(Ic)::value;
            constexpr int AX = T::axis(i);
            // add link i's own inertia (composite is expressed in frame i here)
            {
                const S3 Io = link_inertia_o<T, i>(prm);
                cm += prm[CH_MASS + i];
                if constexpr (T::com_mask(i) != 0) ch = add_m<T::com_mask(i)>(ch, link_h<T, i>(prm));
                cI.xx += Io.xx; cI.yy += Io.yy; cI.zz += Io.zz;
                if constexpr (T::full_inertia(i)) { cI.xy += Io.xy; cI.xz += Io.xz; cI.yz += Io.yz; }
            }
            // unit acceleration about the joint axis: F = e x ch, N = cI e
            V3 F, N;
            if constexpr (AX == 0) { F = {0.0, -ch.z, ch.y}; N = {cI.xx, cI.xy, cI.xz}; }
            else if constexpr (AX == 1) { F = {ch.z, 0.0, -ch.x}; N = {cI.xy, cI.yy, cI.yz}; }
            else { F = {-ch.y, ch.x, 0.0}; N = {cI.xz, cI.yz, cI.zz}; }
            sc.st(sc_m(i, i), comp<AX>(N) + prm[CH_ARMATURE + i]);
            // carry the wrench down to every ancestor joint
            static_for<i>([&](auto Jc) {
                constexpr int j = i - decltype(Jc)::value;      // hop through joint j: frame j -> j-1
                constexpr int AJ = T::axis(j);
                F = rot<AJ>(F, sn[j], cs[j]);
                N = rot<AJ>(N, sn[j], cs[j]);
                if constexpr (T::off_mask(j) != 0) N = cross_m_acc<T::off_mask(j), -1>(N, F, link_offset<T, j>(prm));
                sc.st(sc_m(i, j - 1), comp<T::axis(j - 1)>(N));
            });
            // move the composite to the parent frame
            if constexpr (i > 0) {
                constexpr int OM = T::off_mask(i);
                const V3 hr = rot<AX>(ch, sn[i], cs[i]);
                cI = rotS<AX>(cI, sn[i], cs[i]);
                if constexpr (OM != 0) {
                    const V3 o = link_offset<T, i>(prm);
                    // reference point moves by -o: I += (2 h.o + m o.o) 1 - (h o' + o h') - m o o'
                    const V3 mo = cm * o;
                    const V3 g = hr + hr + mo;                  // 2h + m o
                    double tr = 0.0;
                    if constexpr (OM & 1) tr += g.x * o.x;
                    if constexpr (OM & 2) tr += g.y * o.y;
                    if constexpr (OM & 4) tr += g.z * o.z;
                    if constexpr (OM & 1) cI.xx -= g.x * o.x;
                    if constexpr (OM & 2) cI.yy -= g.y * o.y;
                    if constexpr (OM & 4) cI.zz -= g.z * o.z;
                    cI.xx += tr; cI.yy += tr; cI.zz += tr;
                    // off-diagonals: -(h_a o_b + o_a h_b) - m o_a o_b
                    if constexpr (OM & 2) cI.xy -= hr.x * o.y;
                    if constexpr (OM & 1) cI.xy -= o.x * hr.y;
                    if constexpr ((OM & 3) == 3) cI.xy -= mo.x * o.y;
                    if constexpr (OM & 4) cI.xz -= hr.x * o.z;
                    if constexpr (OM & 1) cI.xz -= o.x * hr.z;
                    if constexpr ((OM & 5) == 5) cI.xz -= mo.x * o.z;
                    if constexpr (OM & 4) cI.yz -= hr.y * o.z;
                    if constexpr (OM & 2) cI.yz -= o.y * hr.z;
                    if constexpr ((OM & 6) == 6) cI.yz -= mo.y * o.z;
                    ch = add_m<OM>(hr, mo);
                } else {
                    ch = hr;
                }
            }
        });
    }
}

// Reciprocal of a well-scaled positive number (mass-matrix pivots, regularisers): hardware seed
// (rcp.approx.ftz.f64 = MUFU.RCP64H, ~20 bits) refined by one third-order step
// r = r0 (1 + e + e^2), e = 1 - x r0  (error e^3, below 1 ulp) -- three dependent FP64 operations on the
// critical path of every pivot instead of the ~30 instructions (with a slow-path call) of an IEEE
// division.  No denormal / special-case handling: callers pass well-scaled positive values only.
MJB_HD double rcp_pos(double x) {
#if defined(__CUDA_ARCH__)
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
    const double e = fma(-x, r0, 1.0);
    const double t = fma(e, e, e);
    return fma(r0, t, r0);
#else
    return 1.0 / x;
#endif
}

// In-place LDL' of the lower triangle of a 7x7 SPD matrix: A[i][j] (i>j) <- L[i][j], dinv[j] = 1/D[j].
// Right-looking (outer-product) order: as soon as a pivot's reciprocal is known its column is scaled
// and the trailing block updated with independent FMAs, so the dependency chain from one pivot to the
// next is reciprocal -> scale -> one FMA (the left-looking form chains up to six FMAs in front of
// every reciprocal).  With two warps per scheduler this chain length is what the solve costs.
MJB_HD void ldl7(double (&A)[7][7], double (&dinv)[7]) {
#pragma unroll
    for (int j = 0; j < 7; j++) {
        const double inv = rcp_pos(A[j][j]);
        dinv[j] = inv;
        double c[7];
#pragma unroll
        for (int i = j + 1; i < 7; i++) { c[i] = A[i][j]; A[i][j] = c[i] * inv; }
#pragma unroll
        for (int i = j + 1; i < 7; i++) {
#pragma unroll
            for (int k = j + 1; k <= i; k++) A[i][k] = fma(-A[i][j], c[k], A[i][k]);
        }
    }
}
// L y = b, z = D^-1 y, L' x = z.  Both substitutions apply the unknown found LAST through the final FMA of a
// row (forward: k ascending; backward: k descending), so consecutive rows are one FMA apart on the
// dependency chain; with k ascending the backward pass chained 7 - i operations per row (27 instead of 7
// dependent FP64 operations at 8.2 cycles each: measured 3.5 % of the whole rollout kernel).
MJB_HD void ldl7_solve(const double (&L)[7][7], const double (&dinv)[7], double (&b)[7]) {
#pragma unroll
    for (int i = 1; i < 7; i++) {
#pragma unroll
        for (int k = 0; k < i; k++) b[i] = fma(-L[i][k], b[k], b[i]);
    }
#pragma unroll
    for (int i = 0; i < 7; i++) b[i] *= dinv[i];
#pragma unroll
    for (int i = 5; i >= 0; i--) {
#pragma unroll
        for (int k = 6; k > i; k--) b[i] = fma(-L[k][i], b[k], b[i]);
    }
}

// MuJoCo's constraint impedance / regulariser / reference acceleration of one soft row
// (mj_makeImpedance + mj_referenceConstraint): generic form, out of line.  The hot path below uses a
// branch-free specialisation for the default power-2 impedance and falls back to this one otherwise.
struct RowParams { double D, aref; };
template <class P> MJB_NOINLINE RowParams soft_row(P prm, double pos_minus_margin, double vel, double invweight) {
    const double d0 = prm[CS_IMP_D0], dw = prm[CS_IMP_DW], width = prm[CS_IMP_WIDTH];
    double imp;
    if (d0 == dw || width <= 1e-15) imp = 0.5 * (d0 + dw);
    else {
        const double x = fabs(pos_minus_margin / width);
        if (x >= 1.0) imp = dw;
        else if (x <= 0.0) imp = d0;
        else {
            const double mid = prm[CS_IMP_MID], pw = prm[CS_IMP_POWER];
            double y;
            if (pw == 1.0) y = x;
            else if (pw == 2.0) y = x <= mid ? (x * x) / mid : 1.0 - ((1.0 - x) * (1.0 - x)) / (1.0 - mid);
            else if (x <= mid) y = pow(x, pw) / pow(mid, pw - 1.0);
            else y = 1.0 - pow(1.0 - x, pw) / pow(1.0 - mid, pw - 1.0);
            imp = d0 + y * (dw - d0);
        }
    }
    double r = (1.0 - imp) * invweight / imp;      // regulariser R; D = 1/R
    r = r < 1e-15 ? 1e-15 : r;
    RowParams o;
    o.D = 1.0 / r;
    o.aref = -prm[CS_SOLB] * vel - prm[CS_SOLK] * imp * pos_minus_margin;
    return o;
}

// World position of a point given in the last link's frame (nested evaluation, no matrices).
template <class T, class P>
MJB_HD V3 chain_point_world(const P& prm, const double (&sn)[7], const double (&cs)[7], V3 p) {
    static_for<7>([&](auto Ic) {
        constexpr int i = 6 - decltype(Ic)::value;
        p = rot<T::axis(i)>(p, sn[i], cs[i]);
        if constexpr (T::off_mask(i) != 0) p = add_m<T::off_mask(i)>(p, link_offset<T, i>(prm));
    });
    return p;
}

// Jacobian row of the sphere-plane contact: the contact point and the plane normal are carried into
// every link frame, Jc[j] = n_j . (e_j x r_j).  Rare path, kept out of line; it works on a packed copy
// (io[0..7) sin, [7..14) cos, [14..21) qvel in; [0..7) Jc, [7] D, [8] aref out) so that the caller's
// register-resident arrays never have their address taken.
template <class T, class P>
MJB_NOINLINE void contact_row(P prm, double* io, double cx, double cy, double cz, double dist) {
    const double radius = prm[CS_CON_RADIUS], margin = prm[CS_CON_MARGIN];
    V3 n = {0.0, 0.0, 1.0};
    V3 r = {cx, cy, cz - (radius + 0.5 * dist)};
    double vel = 0.0, J[7];
    static_for<7>([&](auto Ic) {
        constexpr int i = decltype(Ic)::value;
        constexpr int AX = T::axis(i);
        if constexpr (T::off_mask(i) != 0) r = r - link_offset<T, i>(prm);
        r = rotT<AX>(r, io[i], io[7 + i]);
        n = rotT<AX>(n, io[i], io[7 + i]);
        double j;
        if constexpr (AX == 0) j = n.z * r.y - n.y * r.z;
        else if constexpr (AX == 1) j = n.x * r.z - n.z * r.x;
        else j = n.y * r.x - n.x * r.y;
        J[i] = j;
        vel += j * io[14 + i];
    });
    const RowParams rp = soft_row(prm, dist - margin, vel, prm[CS_CON_INVW]);
#pragma unroll
    for (int i = 0; i < 7; i++) io[i] = J[i];
    io[7] = rp.D;
    io[8] = rp.aref;
}

// Constraint rows of one particle.  Limit row j acts on dof j alone with Jacobian entry sg = +1 below
// the range, -1 above it (at most one side can be violated).  Stored in the row's own sign convention
// so that one comparison decides activity:
//   D[j]   = 1/R_j, 0 when the row is absent
//   bs[j]  = aref_j (= sg * b_j with b_j the joint-space reference acceleration), -DBL_MAX when the row
//            is absent: the row is active at acceleration a iff sg a_j < bs_j; it then adds D_j to H_jj,
//            sg D_j bs_j to the right-hand side, and exerts the joint force sg D_j (bs_j - sg a_j).
//   below  = bit j: joint j is below its range (sg = +1); sgn(j) is the sign mask of sg
// The contact row's Jacobian lives in the scratch (SC_JC) and is only touched when Dc != 0.
struct Rows {
    double D[7], bs[7];
    unsigned below;
    double Dc, arefc;
    MJB_HD unsigned sgn(int j) const { return (~below << (31 - j)) & 0x80000000u; }
};
#define MJB_ROW_ABSENT (-1.7976931348623157e308)

// generic-impedance fallback of make_rows (non-default solimp shapes): out of line
template <class P>
MJB_NOINLINE void limit_rows_generic(P prm, const double* qv, double* Db, unsigned* below_any) {
    const int limited = (int)prm[CS_LIMITED_MASK];
    unsigned below = 0, any = 0;
    for (int j = 0; j < 7; j++) {
        const double lo = prm[CH_RANGE_LO + j], hi = prm[CH_RANGE_HI + j];
        const bool bl = qv[j] < lo;
        const double dist = bl ? qv[j] - lo : hi - qv[j];
        Db[j] = 0.0; Db[7 + j] = 0.0;
        below |= bl ? (1u << j) : 0u;
        if (dist < 0.0 && ((limited >> j) & 1)) {
            const double sg = bl ? 1.0 : -1.0;
            const RowParams rp = soft_row(prm, dist, sg * qv[7 + j], prm[CH_INVW0 + j]);
            Db[j] = rp.D; Db[7 + j] = sg * rp.aref;
            any = 1;
        }
    }
    below_any[0] = below; below_any[1] = any;
}

template <class T, class P, class S>
MJB_HD bool make_rows(const P& prm, S& sc, const double (&q)[7], const double (&qd)[7], const double (&sn)[7],
                      const double (&cs)[7], Rows& R) {
    bool any = false;
    const double d0 = prm[CS_IMP_D0], dw = prm[CS_IMP_DW], width = prm[CS_IMP_WIDTH];
    R.below = 0;
    // Centre of the contact sphere: a chain of 7 dependent rotations that nothing below depends on until the
    // distance test.  It is evaluated inside the branch that computes the limit rows so that ptxas interleaves
    // it with their independent per-joint work (as a basic block of its own it cost 128 cycles of exposed
    // latency per substep for 42 instructions).
    V3 c;
    const V3 cpos = {prm[CS_CON_POS], prm[CS_CON_POS + 1], prm[CS_CON_POS + 2]};
    if (prm[CS_IMP_POWER] == 2.0 && d0 != dw && width > 1e-15) {
        c = chain_point_world<T>(prm, sn, cs, cpos);
        // default impedance shape (power 2)
        const int limited = (int)prm[CS_LIMITED_MASK];
        const double mid = prm[CS_IMP_MID];
        const double K = prm[CS_SOLK], B = prm[CS_SOLB];
        const double niw = -1.0 / width, imid = 1.0 / mid, i1mid = 1.0 / (1.0 - mid), dd = dw - d0;
#pragma unroll
        for (int j = 0; j < 7; j++) {
            const double dlo = q[j] - prm[CH_RANGE_LO + j], dhi = prm[CH_RANGE_HI + j] - q[j];
            const bool below = dlo < 0.0;
            const double dist = dmin(dlo, dhi);                    // side * (range - q): < 0 when violated
            const bool viol = dist < 0.0 && ((limited >> j) & 1);
            R.below |= below ? (1u << j) : 0u;
            any |= viol;
#ifdef MJB_ROWS_BRANCH
            // variant: the row's impedance / regulariser / reference acceleration only where a limit is
            // violated (a warp skips the block for joints none of its particles violates)
            R.D[j] = 0.0; R.bs[j] = MJB_ROW_ABSENT;
            if (viol)
#endif
            {
                const double x = dmin(dist * niw, 1.0);            // penetration / width, saturating (imp = dw at 1)
                const double xm = 1.0 - x;
                const double y = x <= mid ? (x * x) * imid : fma(-(xm * xm), i1mid, 1.0);
                const double imp = fma(y, dd, d0);
                // D = 1 / max(1e-15, (1-imp)*invweight/imp)
                const double den = dmax((1.0 - imp) * prm[CH_INVW0 + j], 1e-15 * imp);
                const double kd = K * imp * dist;
                const double dj = imp * rcp_pos(den);
                const double bsj = -B * sflip(qd[j], below ? 0u : 0x80000000u) - kd;     // aref = -B sg qd - K imp dist
#ifdef MJB_ROWS_BRANCH
                R.D[j] = dj; R.bs[j] = bsj;
#else
                R.D[j] = viol ? dj : 0.0;                          // branch-free: every lane does the same work
                R.bs[j] = viol ? bsj : MJB_ROW_ABSENT;
#endif
            }
        }
    } else {
        c = chain_point_world<T>(prm, sn, cs, cpos);
        double qv[14], Db[14];
        unsigned ba[2];
#pragma unroll
        for (int j = 0; j < 7; j++) { qv[j] = q[j]; qv[7 + j] = qd[j]; }
        limit_rows_generic(prm, qv, Db, ba);
        R.below = ba[0]; any = ba[1] != 0;
#pragma unroll
        for (int j = 0; j < 7; j++) { R.D[j] = Db[j]; R.bs[j] = Db[j] > 0.0 ? sflip(Db[7 + j], R.sgn(j)) : MJB_ROW_ABSENT; }
    }
    R.Dc = 0.0; R.arefc = 0.0;
    const double radius = prm[CS_CON_RADIUS];
    const double dist = c.z - prm[CS_CON_PLANE_Z] - radius;
    if (radius > 0.0 && dist < prm[CS_CON_MARGIN]) {
        double io[21];
#pragma unroll
        for (int j = 0; j < 7; j++) { io[j] = sn[j]; io[7 + j] = cs[j]; io[14 + j] = qd[j]; }
        contact_row<T>(prm, io, c.x, c.y, c.z, dist);
#pragma unroll
        for (int j = 0; j < 7; j++) sc.st(SC_JC + j, io[j]);
        R.Dc = io[7]; R.arefc = io[8];
        any = true;
    }
    return any;
}

// --------------------------------------------------------------------------------------------------
// Slow path of the constraint solve, out of line and rolled: used when the sphere-plane contact row
// exists or when the unit active-set steps of the fast path did not settle.  Same problem, solved
// robustly: Newton on the active set from a = 0 with an exact line search (phi'(t) is piecewise
// linear and increasing: bracket the root between breakpoints, solve the linear piece), then the
// implicit-damping Euler solve.  Packed buffer (doubles):
//   w[0..49)   M (lower triangle valid)      w[49..56)  f            w[56..63) h*damping
//   w[63..70)  D (limit rows)                w[70..77)  b            w[77..84) sg (+1 below / -1 above)
//   w[84..91)  Jc                            w[91] Dc   w[92] arefc
//   out: w[49..56) = qacc of the Euler solve
// --------------------------------------------------------------------------------------------------
MJB_NOINLINE void chol_solve7(double (*A)[7], double* x) {      // SPD solve, lower triangle of A, in place
    #pragma unroll 1
    for (int j = 0; j < 7; j++) {
        double d = A[j][j];
        #pragma unroll 1
        for (int k = 0; k < j; k++) d -= A[j][k] * A[j][k];
        d = sqrt(d);
        A[j][j] = d;
        #pragma unroll 1
        for (int i = j + 1; i < 7; i++) {
            double t = A[i][j];
            #pragma unroll 1
            for (int k = 0; k < j; k++) t -= A[i][k] * A[j][k];
            A[i][j] = t / d;
        }
    }
#pragma unroll 1
    for (int i = 0; i < 7; i++) { double t = x[i];
#pragma unroll 1
        for (int k = 0; k < i; k++) t -= A[i][k] * x[k]; x[i] = t / A[i][i]; }
#pragma unroll 1
    for (int i = 6; i >= 0; i--) { double t = x[i];
#pragma unroll 1
        for (int k = i + 1; k < 7; k++) t -= A[k][i] * x[k]; x[i] = t / A[i][i]; }
}

// force_only (role-split rollout): w[49..56) <- the constraint force J'lambda, no Euler solve.
MJB_NOINLINE void constrained_solve_slow(double* w, int force_only = 0) {
    const double *M49 = w, *f = w + 49, *hd = w + 56, *D = w + 63, *b = w + 70, *sg = w + 77, *Jc = w + 84;
    const double Dc = w[91], arefc = w[92];
    double a[7], A[7][7], x[7], jar[8], jp[8], Dr[8];
    #pragma unroll 1
    for (int j = 0; j < 7; j++) { a[j] = 0.0; Dr[j] = D[j]; }
    Dr[7] = Dc;
    #pragma unroll 1
    for (int it = 0; it < 60; it++) {
        // residuals J_r a - aref_r and active set at a
        #pragma unroll 1
        for (int j = 0; j < 7; j++) jar[j] = sg[j] * (a[j] - b[j]);
        jar[7] = -arefc;
        #pragma unroll 1
        for (int j = 0; j < 7; j++) jar[7] += Jc[j] * a[j];
        #pragma unroll 1
        for (int i = 0; i < 7; i++) {
            #pragma unroll 1
            for (int j = 0; j <= i; j++) A[i][j] = M49[i * 7 + j] + ((Dc > 0.0 && jar[7] < 0.0) ? Dc * Jc[i] * Jc[j] : 0.0);
            const bool on = D[i] > 0.0 && jar[i] < 0.0;
            if (on) A[i][i] += D[i];
            x[i] = f[i] + (on ? D[i] * b[i] : 0.0) + ((Dc > 0.0 && jar[7] < 0.0) ? Dc * arefc * Jc[i] : 0.0);
        }
        chol_solve7(A, x);
        // same active set at the candidate -> optimum
        bool same = true;
        double jcx = -arefc;
        #pragma unroll 1
        for (int j = 0; j < 7; j++) jcx += Jc[j] * x[j];
        #pragma unroll 1
        for (int j = 0; j < 7; j++) same = same && ((D[j] > 0.0 && sg[j] * (x[j] - b[j]) < 0.0) == (D[j] > 0.0 && jar[j] < 0.0));
        same = same && ((Dc > 0.0 && jcx < 0.0) == (Dc > 0.0 && jar[7] < 0.0));
        if (same) { for (int j = 0; j < 7; j++) a[j] = x[j]; break; }
        // exact line search along p = x - a
        double g0 = 0.0, h0 = 0.0;
        #pragma unroll 1
        for (int i = 0; i < 7; i++) {
            double s = 0.0, ga = -f[i];
            #pragma unroll 1
            for (int k = 0; k < 7; k++) {
                const double mik = k <= i ? M49[i * 7 + k] : M49[k * 7 + i];
                s += mik * (x[k] - a[k]); ga += mik * a[k];
            }
            g0 += (x[i] - a[i]) * ga; h0 += (x[i] - a[i]) * s;
        }
        #pragma unroll 1
        for (int j = 0; j < 7; j++) jp[j] = sg[j] * (x[j] - a[j]);
        jp[7] = jcx - jar[7];
        double lo = 0.0, hi = INFINITY;
        #pragma unroll 1
        for (int r = 0; r < 8; r++) {
            if (Dr[r] > 0.0 && jp[r] != 0.0) {
                const double t = -jar[r] / jp[r];
                if (t > 0.0) {
                    double d = g0 + t * h0;
                    #pragma unroll 1
                    for (int q = 0; q < 8; q++) {
                        const double js = jar[q] + t * jp[q];
                        if (Dr[q] > 0.0 && js < 0.0) d += Dr[q] * js * jp[q];
                    }
                    if (d <= 0.0) { if (t > lo) lo = t; }
                    else if (t < hi) hi = t;
                }
            }
        }
        const double mid = (hi == INFINITY) ? lo + 1.0 : 0.5 * (lo + hi);
        double c0 = g0, c1 = h0;
        #pragma unroll 1
        for (int q = 0; q < 8; q++)
            if (Dr[q] > 0.0 && jar[q] + mid * jp[q] < 0.0) { c0 += Dr[q] * jar[q] * jp[q]; c1 += Dr[q] * jp[q] * jp[q]; }
        double t = -c0 / c1;
        t = t < lo ? lo : (t > hi ? hi : t);
        #pragma unroll 1
        for (int j = 0; j < 7; j++) a[j] += t * (x[j] - a[j]);
    }
    // constraint force at the optimum, then (M + hB) qacc = f + J'lambda
    double jc = -arefc;
    #pragma unroll 1
    for (int j = 0; j < 7; j++) jc += Jc[j] * a[j];
    const double lc = (Dc > 0.0 && jc < 0.0) ? -Dc * jc : 0.0;
    if (force_only) {
        #pragma unroll 1
        for (int i = 0; i < 7; i++) {
            const double r = sg[i] * (a[i] - b[i]);
            x[i] = ((D[i] > 0.0 && r < 0.0) ? D[i] * (b[i] - a[i]) : 0.0) + lc * Jc[i];
        }
        #pragma unroll 1
        for (int j = 0; j < 7; j++) w[49 + j] = x[j];
        return;
    }
    #pragma unroll 1
    for (int i = 0; i < 7; i++) {
        const double r = sg[i] * (a[i] - b[i]);
        x[i] = f[i] + ((D[i] > 0.0 && r < 0.0) ? D[i] * (b[i] - a[i]) : 0.0) + lc * Jc[i];
        #pragma unroll 1
        for (int j = 0; j <= i; j++) A[i][j] = M49[i * 7 + j];
        A[i][i] += hd[i];
    }
    chol_solve7(A, x);
    #pragma unroll 1
    for (int j = 0; j < 7; j++) w[49 + j] = x[j];
}

// One mj_step of the chain: forward dynamics at (q, qd) under the actuator torques held in the scratch
// (SC_U = gear * clip(ctrl), computed once per env step), then Euler advance.  sn/cs must hold sin/cos
// of q on entry.  Returns true when a constraint row was present.
//
// Constraint forces: exact minimiser over a of 1/2 a'Ma - f'a + sum_r 1/2 D_r min(0, J_r a - aref_r)^2
// by Newton on the active set (each piece of the objective is quadratic, so a solve whose own active
// set equals the set it was built from satisfies the optimality conditions).  Fast path (limit rows
// only -- they act on single dofs, so H = M + diagonal): the Newton solves and the final
// implicit-damping Euler solve (M + hB) a = f + J'lambda share ONE factor/solve site fed by a diagonal
// increment and a right-hand-side increment:
//     Newton:  H = M + diag(dadd),  rhs = f + radd,   dadd = D on the active rows, radd = D b
//     Euler:   H = M + diag(h B),   rhs = f + fc,     fc = D (b - a) on the active rows
// REPAIR = false: without the rank-one repair (a wrong first guess costs another factor/solve trip instead; measured
// round 2: within 2 % either way at every launch size, kept as a compile-time switch for experiments).
template <class T, bool REPAIR = true, class P, class S>
MJB_HD bool chain_substep(const P& prm, S& sc, double (&q)[7], double (&qd)[7], const double (&sn)[7],
                          const double (&cs)[7]) {
    double f[7];
    chain_mass_bias<T>(prm, sc, sn, cs, qd, f);
#pragma unroll
    for (int j = 0; j < 7; j++) f[j] = sc.ld(SC_U + j) - prm[CH_DAMPING + j] * qd[j] - f[j];
    Rows R;
    const bool any = make_rows<T>(prm, sc, q, qd, sn, cs, R);
    double dadd[7], radd[7], x[7];
    unsigned act = 0;
    bool slow = false;
    int phase = 1;
    if (!any) {
#pragma unroll
        for (int j = 0; j < 7; j++) { dadd[j] = prm[CH_HDAMP + j]; radd[j] = 0.0; }
    } else {
        // first active-set guess from the decoupled accelerations f_j / M_jj; any guess is admissible,
        // it only has to be confirmed by its own solve
        phase = 0;
#pragma unroll
        for (int j = 0; j < 7; j++) {
            const double bm = R.bs[j] * sc.ld(sc_m(j, j));
            const bool on = sflip(f[j], R.sgn(j)) < bm;            // absent rows (bs = -DBL_MAX) are never on
            act |= on ? (1u << j) : 0u;
            dadd[j] = on ? R.D[j] : 0.0;
            radd[j] = sflip(dadd[j] * R.bs[j], R.sgn(j));
        }
    }
    // Contact row (rare): its rank-one term is not diagonal, so the matrix with the term added is kept
    // as a second copy in the scratch (M2 = M + Dc Jc Jc') and the Newton passes read M or M2 through
    // a slot offset; all contact bookkeeping sits in cold branches around the shared factor/solve.
    int moff = 0;
    bool con_on = false;
    if (R.Dc > 0.0) {
#pragma unroll 1
        for (int i = 0; i < 7; i++) {
            const double ji = R.Dc * sc.ld(SC_JC + i);
#pragma unroll 1
            for (int j = 0; j <= i; j++)
                sc.st(SC_M2 + i * (i + 1) / 2 + j, sc.ld(SC_M + i * (i + 1) / 2 + j) + ji * sc.ld(SC_JC + j));
        }
        con_on = R.arefc > 0.0;          // guess: the reference acceleration alone pushes out of the plane
        if (con_on) {
            moff = SC_M2 - SC_M;
#pragma unroll
            for (int j = 0; j < 7; j++) radd[j] += R.Dc * R.arefc * sc.ld(SC_JC + j);
        }
    }
    MJB_STAT(0); if (any) MJB_STAT(1);
#if defined(MJB_HOST_STATS)
    int ntrip = 0;
#endif
    {
        // Every trip is one factor/solve; a Newton trip then checks its active set and prepares the next
        // trip's increments.  The loop condition is laundered (MJB_OPAQUE) so the compiler cannot thread the
        // "skip the check on the Euler trip" branch into a mid-body exit and rotate the loop into
        // "solve; while (..) { check; solve }" -- that doubled the instruction-cache footprint of the solve.
        int iters = 0;
        int again;
#pragma unroll 1
        do {
            double H[7][7], dinv[7];
            MJB_OPAQUE(phase);
            MJB_STAT(2);
#if defined(MJB_HOST_STATS)
            ntrip++;
#endif
#pragma unroll
            for (int i = 0; i < 7; i++) {
#pragma unroll
                for (int j = 0; j < i; j++) H[i][j] = sc.ld(moff + sc_m(i, j));
                H[i][i] = sc.ld(moff + sc_m(i, i)) + dadd[i];
                x[i] = f[i] + radd[i];
            }
            ldl7(H, dinv);
            ldl7_solve(H, dinv, x);
            again = 0;
            if (phase == 0) {
                unsigned actx = 0;
#pragma unroll
                for (int j = 0; j < 7; j++) actx |= (sflip(x[j], R.sgn(j)) < R.bs[j]) ? (1u << j) : 0u;
                bool con_x = false;
                double jcx = 0.0;
                if (R.Dc > 0.0) {
                    jcx = -R.arefc;
#pragma unroll
                    for (int j = 0; j < 7; j++) jcx += sc.ld(SC_JC + j) * x[j];
                    con_x = jcx < 0.0;
                }
                // same set as the one the solve was built from: optimum, next trip is the Euler solve with
                // the joint-space constraint force D_j (b_j - a_j); otherwise a plain active-set (unit Newton) step
                bool ok = (actx == act) & (con_x == con_on);
#ifndef MJB_NO_REPAIR
                if constexpr (REPAIR) {
                // One limit row misjudged (98 % of the wrong guesses; measured on the host): repair the solution
                // by a rank-one update on the factor at hand instead of a new factor/solve trip.  With
                // H z = e_j:  (H + s D_j e_j e_j') x' = rhs + s rho_j e_j  has  x' = x + alpha z,
                // alpha = s (rho_j - D_j x_j) / (1 + s D_j z_j),  s = +1 row j turns on, -1 off.  MuJoCo's
                // regulariser makes D_j (H^-1)_jj ~ imp/(1-imp) = 9..19, so the denominator of the "off" case
                // (1 - D_j z_j = 1/(1 + D_j h)) loses one digit, not more.  The repaired solution is then checked
                // like any other; a warp whose lanes all settle here skips the third trip the wrong guess cost
                // (64 % of warp-substeps on the bench workload).
                const unsigned fm = actx ^ act;
                if (REPAIR && !ok && R.Dc == 0.0 && (fm & (fm - 1u)) == 0u) {
                    MJB_STAT(4);
                    double z[7];
#pragma unroll
                    for (int j = 0; j < 7; j++) z[j] = ((fm >> j) & 1u) ? 1.0 : 0.0;
                    ldl7_solve(H, dinv, z);
                    double Dj = 0.0, wj = 0.0, zj = 0.0;
#pragma unroll
                    for (int j = 0; j < 7; j++) {
                        const bool b = (fm >> j) & 1u;
                        // rho_j - D_j x_j = D_j sg_j (bs_j - sg_j x_j)
                        Dj = b ? R.D[j] : Dj;
                        wj = b ? sflip(R.bs[j] - sflip(x[j], R.sgn(j)), R.sgn(j)) : wj;
                        zj = b ? z[j] : zj;
                    }
                    const bool on = (actx & fm) != 0u;
                    const double sD = on ? Dj : -Dj;
                    const double alpha = sD * wj * rcp_pos(fma(sD, zj, 1.0));
#pragma unroll
                    for (int j = 0; j < 7; j++) x[j] = fma(alpha, z[j], x[j]);
                    act ^= fm;
                    actx = 0;
#pragma unroll
                    for (int j = 0; j < 7; j++) actx |= (sflip(x[j], R.sgn(j)) < R.bs[j]) ? (1u << j) : 0u;
                    ok = actx == act;
                    if (ok) MJB_STAT(5);
                }
                }
#endif
#pragma unroll
                for (int j = 0; j < 7; j++) {
                    const double De = ((actx >> j) & 1u) ? R.D[j] : 0.0;
                    dadd[j] = ok ? prm[CH_HDAMP + j] : De;
                    radd[j] = sflip(De * (R.bs[j] - (ok ? sflip(x[j], R.sgn(j)) : 0.0)), R.sgn(j));
                }
                if (con_x) {
                    // contact force -Dc (Jc.a - arefc) Jc at the optimum, or the row's right-hand side Dc arefc Jc
                    const double lc = ok ? -R.Dc * jcx : R.Dc * R.arefc;
#pragma unroll
                    for (int j = 0; j < 7; j++) radd[j] += lc * sc.ld(SC_JC + j);
                }
                act = actx;
                con_on = con_x;
                moff = (!ok && con_on) ? SC_M2 - SC_M : 0;
                phase = ok ? 1 : 0;
                again = 1;
                if (!ok && ++iters > 6) { slow = true; again = 0; }    // unit steps did not settle: robust path
            }
            MJB_OPAQUE(again);
        } while (again);
    }
    if (slow) {
        double w[93];
        MJB_STAT(3);
#pragma unroll
        for (int j = 0; j < 7; j++) {
#pragma unroll
            for (int k = 0; k <= j; k++) w[j * 7 + k] = sc.ld(sc_m(j, k));
            w[49 + j] = f[j]; w[56 + j] = prm[CH_HDAMP + j];
            w[63 + j] = R.D[j]; w[70 + j] = R.D[j] > 0.0 ? sflip(R.bs[j], R.sgn(j)) : 0.0;
            w[77 + j] = ((R.below >> j) & 1u) ? 1.0 : -1.0;
            w[84 + j] = R.Dc > 0.0 ? sc.ld(SC_JC + j) : 0.0;
        }
        w[91] = R.Dc; w[92] = R.arefc;
        constrained_solve_slow(w);
#pragma unroll
        for (int j = 0; j < 7; j++) x[j] = w[49 + j];
    }
    MJB_STAT_TRIPS(ntrip);
    const double h = prm[CS_TIMESTEP];
#pragma unroll
    for (int j = 0; j < 7; j++) { qd[j] += h * x[j]; q[j] += h * qd[j]; }
    return any;
}

// The Newton half of chain_substep on its own, for the role-split rollout (rollout_reacher_split.cuh): the
// constraint force J'lambda at the minimiser -- what the loop above hands to its Euler trip as `radd` -- from the
// smooth force f, the rows R and M in the scratch (SC_M; SC_JC / SC_M2 when the contact row exists).  Same
// operations in the same order as chain_substep: first guess f_j / M_jj, factor/solve trips on M + diag(D_active),
// rank-one repair of a single misjudged row, slow path after 6 unsettled trips.
// factor_only (warp-uniform): the caller is the warp that needs the factor of M + hB for the Euler solve; it runs
// through the SAME load / factor instructions as the Newton warp next to it (one copy of that code to fetch for
// both) and leaves with (H, dinv) = LDL'(M + hB).  H / dinv are the caller's arrays for that reason.
template <class T, class P, class S>
MJB_HD void constraint_force(const P& prm, S& sc, const double (&f)[7], const Rows& R, double (&fc)[7], double (&H)[7][7],
                             double (&dinv)[7], bool factor_only = false) {
    double dadd[7], radd[7], x[7];
    unsigned act = 0;
    int moff = 0;
    bool con_on = false;
    if (factor_only) {
#pragma unroll
        for (int j = 0; j < 7; j++) { dadd[j] = prm[CH_HDAMP + j]; radd[j] = 0.0; }
    } else {
#pragma unroll
    for (int j = 0; j < 7; j++) {
        const double bm = R.bs[j] * sc.ld(sc_m(j, j));
        const bool on = sflip(f[j], R.sgn(j)) < bm;
        act |= on ? (1u << j) : 0u;
        dadd[j] = on ? R.D[j] : 0.0;
        radd[j] = sflip(dadd[j] * R.bs[j], R.sgn(j));
    }
    }
    if (!factor_only && R.Dc > 0.0) {
#pragma unroll 1
        for (int i = 0; i < 7; i++) {
            const double ji = R.Dc * sc.ld(SC_JC + i);
#pragma unroll 1
            for (int j = 0; j <= i; j++)
                sc.st(SC_M2 + i * (i + 1) / 2 + j, sc.ld(SC_M + i * (i + 1) / 2 + j) + ji * sc.ld(SC_JC + j));
        }
        con_on = R.arefc > 0.0;
        if (con_on) {
            moff = SC_M2 - SC_M;
#pragma unroll
            for (int j = 0; j < 7; j++) radd[j] += R.Dc * R.arefc * sc.ld(SC_JC + j);
        }
    }
    bool slow = false;
    int iters = 0, again;
#pragma unroll 1
    do {
#pragma unroll
        for (int i = 0; i < 7; i++) {
#pragma unroll
            for (int j = 0; j < i; j++) H[i][j] = sc.ld(moff + sc_m(i, j));
            H[i][i] = sc.ld(moff + sc_m(i, i)) + dadd[i];
        }
        ldl7(H, dinv);
        if (factor_only) return;
#pragma unroll
        for (int i = 0; i < 7; i++) x[i] = f[i] + radd[i];
        ldl7_solve(H, dinv, x);
        unsigned actx = 0;
#pragma unroll
        for (int j = 0; j < 7; j++) actx |= (sflip(x[j], R.sgn(j)) < R.bs[j]) ? (1u << j) : 0u;
        bool con_x = false;
        double jcx = 0.0;
        if (R.Dc > 0.0) {
            jcx = -R.arefc;
#pragma unroll
            for (int j = 0; j < 7; j++) jcx += sc.ld(SC_JC + j) * x[j];
            con_x = jcx < 0.0;
        }
        bool ok = (actx == act) & (con_x == con_on);
#ifndef MJB_NO_REPAIR
        const unsigned fm = actx ^ act;
        if (!ok && R.Dc == 0.0 && (fm & (fm - 1u)) == 0u) {
            double z[7];
#pragma unroll
            for (int j = 0; j < 7; j++) z[j] = ((fm >> j) & 1u) ? 1.0 : 0.0;
            ldl7_solve(H, dinv, z);
            double Dj = 0.0, wj = 0.0, zj = 0.0;
#pragma unroll
            for (int j = 0; j < 7; j++) {
                const bool b = (fm >> j) & 1u;
                Dj = b ? R.D[j] : Dj;
                wj = b ? sflip(R.bs[j] - sflip(x[j], R.sgn(j)), R.sgn(j)) : wj;
                zj = b ? z[j] : zj;
            }
            const bool on = (actx & fm) != 0u;
            const double sD = on ? Dj : -Dj;
            const double alpha = sD * wj * rcp_pos(fma(sD, zj, 1.0));
#pragma unroll
            for (int j = 0; j < 7; j++) x[j] = fma(alpha, z[j], x[j]);
            act ^= fm;
            actx = 0;
#pragma unroll
            for (int j = 0; j < 7; j++) actx |= (sflip(x[j], R.sgn(j)) < R.bs[j]) ? (1u << j) : 0u;
            ok = actx == act;
        }
#endif
#pragma unroll
        for (int j = 0; j < 7; j++) {
            const double De = ((actx >> j) & 1u) ? R.D[j] : 0.0;
            dadd[j] = De;
            radd[j] = sflip(De * (R.bs[j] - (ok ? sflip(x[j], R.sgn(j)) : 0.0)), R.sgn(j));
        }
        if (con_x) {
            const double lc = ok ? -R.Dc * jcx : R.Dc * R.arefc;
#pragma unroll
            for (int j = 0; j < 7; j++) radd[j] += lc * sc.ld(SC_JC + j);
        }
        act = actx;
        con_on = con_x;
        moff = con_on ? SC_M2 - SC_M : 0;
        again = ok ? 0 : 1;
        if (!ok && ++iters > 6) { slow = true; again = 0; }
    } while (again);
    if (slow) {
        double w[93];
#pragma unroll
        for (int j = 0; j < 7; j++) {
#pragma unroll
            for (int k = 0; k <= j; k++) w[j * 7 + k] = sc.ld(sc_m(j, k));
            w[49 + j] = f[j]; w[56 + j] = prm[CH_HDAMP + j];
            w[63 + j] = R.D[j]; w[70 + j] = R.D[j] > 0.0 ? sflip(R.bs[j], R.sgn(j)) : 0.0;
            w[77 + j] = ((R.below >> j) & 1u) ? 1.0 : -1.0;
            w[84 + j] = R.Dc > 0.0 ? sc.ld(SC_JC + j) : 0.0;
        }
        w[91] = R.Dc; w[92] = R.arefc;
        constrained_solve_slow(w, 1);
#pragma unroll
        for (int j = 0; j < 7; j++) radd[j] = w[49 + j];
    }
#pragma unroll
    for (int j = 0; j < 7; j++) fc[j] = radd[j];
}

// Motor torque of joint j under control u: gear * clip(u, ctrlrange) (MuJoCo mj_fwdActuation with
// ctrllimited motors, sawyer.xml:101-109).  The control is held over the frame_skip substeps.
template <class P> MJB_HD double actuator_torque(const P& prm, int j, double u) {
    const double lo = prm[CH_CTRL_LO + j], hi = prm[CH_CTRL_HI + j];
    u = u < lo ? lo : u;
    u = u > hi ? hi : u;
    return prm[CH_GEAR + j] * u;
}

// The reference's step cost (reacher_env.py:31-35): -reward = |h-g|_1 + 5 |h-g|_2
MJB_HD double reach_cost(V3 hand, V3 target) {
    const double dx = hand.x - target.x, dy = hand.y - target.y, dz = hand.z - target.z;
    return (fabs(dx) + fabs(dy) + fabs(dz)) + 5.0 * sqrt(dx * dx + dy * dy + dz * dz);
}

}  // namespace mjb
