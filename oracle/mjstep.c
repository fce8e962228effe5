/*
 * oracle/mjstep.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU (FP64, plain C) restatement of the reference's rollout for the reacher_7dof
 * environment.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker / the timed
 * CPU baseline.  The product path (mjmpc_b200/csrc) never links or calls it.
 *
 * PARITY UNPINNED AGAINST MuJoCo ITSELF for the dynamics (pinned instead, rows / forces / steps / rollouts, against
 * the independent second restatement oracle/efc_ref.py: tests/test_efc_pin.py, tests/golden/efc_pin.npz).
 * The arithmetic of this path lives in MuJoCo 2.0
 * (closed binary, reached through mujoco-py>=2.0,<2.1 and mjrl's MujocoEnv), none of
 * which is under /root/reference or installed here.  The reference ships no golden
 * vectors for it.  This file restates MuJoCo's published algorithm (mj_step for a
 * tree of hinge joints: mj_kinematics, mj_comPos, mj_crb, mj_comVel, mj_rne,
 * mj_passive, mj_fwdActuation, soft joint-limit / frictionless-contact constraints,
 * mj_Euler with implicit joint damping) and is anchored on the reference's own call
 * sites:
 *   mjmpc/envs/gym_env_wrapper.py:123-153   per-particle reset, u = mean[t]+noise[b,t],
 *                                           unclipped action recorded
 *   mjmpc/envs/basic/reacher_env.py:29-39   do_simulation(a, frame_skip=2); reward =
 *                                           -(L1 + 5 L2) between finger and target site
 *   mjmpc/envs/basic/reacher_env.py:41-47   observation layout
 *   mjmpc/envs/basic/reacher_env.py:87-99   set_env_state (sim.reset, forward)
 *   mjmpc/envs/assets/xml/sawyer.xml        the model (compiled by mjmpc_b200/envs/model.py)
 * It deliberately uses MuJoCo's own formulation (world-orientation spatial vectors
 * about the tree's centre of mass, composite-rigid-body M, recursive Newton-Euler bias)
 * on the UN-merged 9-body tree, whereas the CUDA kernel uses a link-frame formulation
 * on the merged 7-link chain -- so the two are independent derivations.
 *
 * Soft constraints: MuJoCo minimises, over qacc,
 *     1/2 (a-a0)' M (a-a0) + sum_i 1/2 D_i min(0, J_i a - aref_i)^2
 * with its Newton solver to tolerance 1e-8.  Here the same strictly convex
 * piecewise-quadratic problem is solved to machine precision (Newton steps with an exact
 * piecewise-linear line search), i.e. the limit MuJoCo's iteration converges to.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define MAXB 16
#define MAXV 8
#define MAXROW (2 * MAXV + 1)
#define MJMINVAL 1e-15

typedef struct {
    int nb, nv;
    int parent[MAXB];
    double pos[MAXB][3], mass[MAXB], ipos[MAXB][3], inertia[MAXB][9];
    int jnt_body[MAXV], jnt_limited[MAXV];
    int body_dof[MAXB]; /* dof owned by the body or -1 */
    double jnt_axis[MAXV][3], jnt_range[MAXV][2];
    double armature[MAXV], damping[MAXV], gear[MAXV], ctrlrange[MAXV][2], invweight0[MAXV];
    double timestep;
    int frame_skip;
    double solK, solB, solimp[5];
    int hand_body;
    double hand_pos[3];
    int con_body;
    double con_pos[3], con_radius, con_plane_z, con_margin, con_invweight;
} ora_model;

/* ---- small vector helpers ---- */
static void cross3(double* r, const double* a, const double* b) {
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    r[0] = x; r[1] = y; r[2] = z;
}
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void matvec3(double* r, const double* M, const double* v) {
    double x = M[0] * v[0] + M[1] * v[1] + M[2] * v[2];
    double y = M[3] * v[0] + M[4] * v[1] + M[5] * v[2];
    double z = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
    r[0] = x; r[1] = y; r[2] = z;
}
static void matmul3(double* R, const double* A, const double* B) {
    double T[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) T[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    memcpy(R, T, sizeof(T));
}
static void rot_axis(double* R, const double* a, double q) {
    double s = sin(q), c = cos(q), t = 1.0 - c;
    R[0] = c + t * a[0] * a[0];        R[1] = t * a[0] * a[1] - s * a[2]; R[2] = t * a[0] * a[2] + s * a[1];
    R[3] = t * a[0] * a[1] + s * a[2]; R[4] = c + t * a[1] * a[1];        R[5] = t * a[1] * a[2] - s * a[0];
    R[6] = t * a[0] * a[2] - s * a[1]; R[7] = t * a[1] * a[2] + s * a[0]; R[8] = c + t * a[2] * a[2];
}

/* spatial inertia about a reference point: mass, first moment h = m*(c - P), inertia about P */
typedef struct { double m, h[3], I[9]; } sinert;
/* momentum = I * motion;  motion = (w, v), force = (torque about P, force) */
static void sinert_mul(double* f, const sinert* S, const double* mv) {
    double t[3], u[3];
    matvec3(t, S->I, mv);          /* I w */
    cross3(u, S->h, mv + 3);       /* h x v */
    f[0] = t[0] + u[0]; f[1] = t[1] + u[1]; f[2] = t[2] + u[2];
    cross3(u, mv, S->h);           /* w x h */
    f[3] = S->m * mv[3] + u[0]; f[4] = S->m * mv[4] + u[1]; f[5] = S->m * mv[5] + u[2];
}
static void cross_motion(double* r, const double* v, const double* m) {
    double a[3], b[3], c[3];
    cross3(a, v, m); cross3(b, v, m + 3); cross3(c, v + 3, m);
    r[0] = a[0]; r[1] = a[1]; r[2] = a[2];
    r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
static void cross_force(double* r, const double* v, const double* f) {
    double a[3], b[3], c[3];
    cross3(a, v, f); cross3(b, v + 3, f + 3); cross3(c, v, f + 3);
    r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2];
    r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}
static double dot6(const double* a, const double* b) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}

/* dense SPD solve (Cholesky), n <= MAXV; A is overwritten */
static void chol_solve(int n, double A[MAXV][MAXV], double* b) {
    for (int j = 0; j < n; j++) {
        double s = A[j][j];
        for (int k = 0; k < j; k++) s -= A[j][k] * A[j][k];
        A[j][j] = sqrt(s);
        for (int i = j + 1; i < n; i++) {
            double t = A[i][j];
            for (int k = 0; k < j; k++) t -= A[i][k] * A[j][k];
            A[i][j] = t / A[j][j];
        }
    }
    for (int i = 0; i < n; i++) {
        double t = b[i];
        for (int k = 0; k < i; k++) t -= A[i][k] * b[k];
        b[i] = t / A[i][i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double t = b[i];
        for (int k = i + 1; k < n; k++) t -= A[k][i] * b[k];
        b[i] = t / A[i][i];
    }
}

typedef struct {
    double xpos[MAXB][3], xmat[MAXB][9];
    double anchor[MAXV][3], axis[MAXV][3];
    double M[MAXV][MAXV], bias[MAXV];
    double hand[3];
    int nefc;
    double qacc[MAXV];
    /* constraint rows of the last step, for row-by-row checks against oracle/efc_ref.py */
    double efc_J[MAXROW][MAXV], efc_aref[MAXROW], efc_D[MAXROW], efc_force[MAXROW], qfrc_constraint[MAXV];
} ora_data;

static double impedance(const double* solimp, double pos, double margin) {
    if (solimp[0] == solimp[1] || solimp[2] <= MJMINVAL) return 0.5 * (solimp[0] + solimp[1]);
    double x = fabs((pos - margin) / solimp[2]);
    if (x >= 1.0) return solimp[1];
    if (x <= 0.0) return solimp[0];
    double y;
    if (solimp[4] == 1.0) y = x;
    else if (x <= solimp[3]) y = pow(x, solimp[4]) / pow(solimp[3], solimp[4] - 1.0);
    else y = 1.0 - pow(1.0 - x, solimp[4]) / pow(1.0 - solimp[3], solimp[4] - 1.0);
    return solimp[0] + y * (solimp[1] - solimp[0]);
}

/* One mj_step: forward dynamics at (q,v) with control u, then Euler advance (in place). */
static void ora_step(const ora_model* m, ora_data* d, double* q, double* v, const double* u) {
    const int nb = m->nb, nv = m->nv;
    const double h = m->timestep;
    /* --- mj_kinematics --- */
    for (int b = 0; b < nb; b++) {
        const int pa = m->parent[b];
        double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        const double* Rp = pa < 0 ? I3 : d->xmat[pa];
        double off[3];
        matvec3(off, Rp, m->pos[b]);
        for (int k = 0; k < 3; k++) d->xpos[b][k] = (pa < 0 ? 0.0 : d->xpos[pa][k]) + off[k];
        const int j = m->body_dof[b];
        if (j >= 0) {
            double Rj[9];
            rot_axis(Rj, m->jnt_axis[j], q[j]);
            matmul3(d->xmat[b], Rp, Rj);
            matvec3(d->axis[j], Rp, m->jnt_axis[j]);
            for (int k = 0; k < 3; k++) d->anchor[j][k] = d->xpos[b][k];
        } else {
            memcpy(d->xmat[b], Rp, 9 * sizeof(double));
        }
    }
    {
        double hp[3];
        matvec3(hp, d->xmat[m->hand_body], m->hand_pos);
        for (int k = 0; k < 3; k++) d->hand[k] = d->xpos[m->hand_body][k] + hp[k];
    }
    /* --- mj_comPos: tree COM, inertias about it, dof axes about it --- */
    double com[3] = {0, 0, 0}, mtot = 0;
    double xipos[MAXB][3];
    for (int b = 0; b < nb; b++) {
        double t[3];
        matvec3(t, d->xmat[b], m->ipos[b]);
        for (int k = 0; k < 3; k++) { xipos[b][k] = d->xpos[b][k] + t[k]; com[k] += m->mass[b] * xipos[b][k]; }
        mtot += m->mass[b];
    }
    for (int k = 0; k < 3; k++) com[k] /= mtot;
    sinert cin[MAXB], crb[MAXB];
    for (int b = 0; b < nb; b++) {
        double RI[9], Rt[9], Iw[9], r[3];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Rt[3 * i + j] = d->xmat[b][3 * j + i];
        matmul3(RI, d->xmat[b], m->inertia[b]);
        matmul3(Iw, RI, Rt);
        for (int k = 0; k < 3; k++) r[k] = xipos[b][k] - com[k];
        const double mm = m->mass[b], rr = dot3(r, r);
        cin[b].m = mm;
        for (int k = 0; k < 3; k++) cin[b].h[k] = mm * r[k];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) cin[b].I[3 * i + j] = Iw[3 * i + j] + mm * ((i == j ? rr : 0.0) - r[i] * r[j]);
        crb[b] = cin[b];
    }
    double cdof[MAXV][6];
    for (int j = 0; j < nv; j++) {
        double off[3];
        for (int k = 0; k < 3; k++) { cdof[j][k] = d->axis[j][k]; off[k] = com[k] - d->anchor[j][k]; }
        cross3(cdof[j] + 3, d->axis[j], off);
    }
    /* --- mj_crb --- */
    for (int b = nb - 1; b >= 0; b--) {
        const int pa = m->parent[b];
        if (pa >= 0) {
            crb[pa].m += crb[b].m;
            for (int k = 0; k < 3; k++) crb[pa].h[k] += crb[b].h[k];
            for (int k = 0; k < 9; k++) crb[pa].I[k] += crb[b].I[k];
        }
    }
    for (int i = 0; i < nv; i++) for (int j = 0; j < nv; j++) d->M[i][j] = 0.0;
    for (int j = 0; j < nv; j++) {
        double buf[6];
        sinert_mul(buf, &crb[m->jnt_body[j]], cdof[j]);
        d->M[j][j] = m->armature[j] + dot6(cdof[j], buf);
        /* ancestors */
        int b = m->parent[m->jnt_body[j]];
        while (b >= 0) {
            const int i = m->body_dof[b];
            if (i >= 0) { d->M[i][j] = dot6(cdof[i], buf); d->M[j][i] = d->M[i][j]; }
            b = m->parent[b];
        }
    }
    /* --- mj_comVel + mj_rne (bias only, gravity-free base acceleration) --- */
    double cvel[MAXB][6], cacc[MAXB][6], cfrc[MAXB][6];
    for (int b = 0; b < nb; b++) {
        const int pa = m->parent[b];
        for (int k = 0; k < 6; k++) { cvel[b][k] = pa < 0 ? 0.0 : cvel[pa][k]; cacc[b][k] = pa < 0 ? 0.0 : cacc[pa][k]; }
        const int j = m->body_dof[b];
        if (j >= 0) {
            double cdofdot[6];
            cross_motion(cdofdot, cvel[b], cdof[j]);
            for (int k = 0; k < 6; k++) { cvel[b][k] += cdof[j][k] * v[j]; cacc[b][k] += cdofdot[k] * v[j]; }
        }
        double Ia[6], Iv[6], vxIv[6];
        sinert_mul(Ia, &cin[b], cacc[b]);
        sinert_mul(Iv, &cin[b], cvel[b]);
        cross_force(vxIv, cvel[b], Iv);
        for (int k = 0; k < 6; k++) cfrc[b][k] = Ia[k] + vxIv[k];
    }
    for (int b = nb - 1; b >= 0; b--) {
        const int pa = m->parent[b];
        if (pa >= 0) for (int k = 0; k < 6; k++) cfrc[pa][k] += cfrc[b][k];
    }
    for (int j = 0; j < nv; j++) d->bias[j] = dot6(cdof[j], cfrc[m->jnt_body[j]]);
    /* --- passive + actuation -> qfrc_smooth --- */
    double f[MAXV];
    for (int j = 0; j < nv; j++) {
        double c = u[j];
        if (c < m->ctrlrange[j][0]) c = m->ctrlrange[j][0];
        if (c > m->ctrlrange[j][1]) c = m->ctrlrange[j][1];
        f[j] = -m->damping[j] * v[j] - d->bias[j] + m->gear[j] * c;
    }
    /* --- constraint rows --- */
    int nr = 0;
    double J[MAXROW][MAXV], aref[MAXROW], D[MAXROW];
    for (int j = 0; j < nv; j++) {
        if (!m->jnt_limited[j]) continue;
        for (int side = -1; side <= 1; side += 2) {
            const double dist = side * (m->jnt_range[j][(side + 1) / 2] - q[j]);
            if (dist < 0.0) {
                for (int k = 0; k < nv; k++) J[nr][k] = 0.0;
                J[nr][j] = -side;
                const double imp = impedance(m->solimp, dist, 0.0);
                double R = (1.0 - imp) * m->invweight0[j] / imp;
                if (R < MJMINVAL) R = MJMINVAL;
                D[nr] = 1.0 / R;
                aref[nr] = -m->solB * (-side * v[j]) - m->solK * imp * dist;
                nr++;
            }
        }
    }
    if (m->con_radius > 0.0) {
        double c[3], t[3];
        matvec3(t, d->xmat[m->con_body], m->con_pos);
        for (int k = 0; k < 3; k++) c[k] = d->xpos[m->con_body][k] + t[k];
        const double dist = c[2] - m->con_plane_z - m->con_radius;
        if (dist < m->con_margin) {
            const double cp[3] = {c[0], c[1], c[2] - (m->con_radius + 0.5 * dist)};
            for (int k = 0; k < nv; k++) J[nr][k] = 0.0;
            int b = m->con_body;
            while (b >= 0) {
                const int j = m->body_dof[b];
                if (j >= 0) {
                    double r[3], w[3];
                    for (int k = 0; k < 3; k++) r[k] = cp[k] - d->anchor[j][k];
                    cross3(w, d->axis[j], r);
                    J[nr][j] = w[2]; /* normal = +z */
                }
                b = m->parent[b];
            }
            const double imp = impedance(m->solimp, dist, m->con_margin);
            double R = (1.0 - imp) * m->con_invweight / imp;
            if (R < MJMINVAL) R = MJMINVAL;
            D[nr] = 1.0 / R;
            double vel = 0;
            for (int k = 0; k < nv; k++) vel += J[nr][k] * v[k];
            aref[nr] = -m->solB * vel - m->solK * imp * (dist - m->con_margin);
            nr++;
        }
    }
    d->nefc = nr;
    double fc[MAXV];
    for (int j = 0; j < nv; j++) fc[j] = 0.0;
    if (nr > 0) {
        /* exact minimiser of the convex piecewise-quadratic constraint problem */
        double a[MAXV], A[MAXV][MAXV];
        for (int i = 0; i < nv; i++) { a[i] = f[i]; for (int j = 0; j < nv; j++) A[i][j] = d->M[i][j]; }
        chol_solve(nv, A, a); /* qacc_smooth */
        for (int iter = 0; iter < 100; iter++) {
            double jar[MAXROW], g[MAXV], fn = 0, gn = 0;
            int act[MAXROW];
            for (int r = 0; r < nr; r++) {
                double s = -aref[r];
                for (int k = 0; k < nv; k++) s += J[r][k] * a[k];
                jar[r] = s; act[r] = s < 0.0;
            }
            for (int i = 0; i < nv; i++) {
                double s = -f[i];
                for (int k = 0; k < nv; k++) s += d->M[i][k] * a[k];
                for (int r = 0; r < nr; r++) if (act[r]) s += D[r] * jar[r] * J[r][i];
                g[i] = s; gn += s * s; fn += f[i] * f[i];
            }
            if (sqrt(gn) <= 1e-15 * (1.0 + sqrt(fn))) break;
            double p[MAXV], Jp[MAXROW];
            for (int i = 0; i < nv; i++) {
                p[i] = -g[i];
                for (int j = 0; j < nv; j++) {
                    double s = d->M[i][j];
                    for (int r = 0; r < nr; r++) if (act[r]) s += D[r] * J[r][i] * J[r][j];
                    A[i][j] = s;
                }
            }
            chol_solve(nv, A, p);
            double g0 = 0, h0 = 0; /* smooth part of phi'(t) = g0 + t*h0 */
            for (int i = 0; i < nv; i++) {
                double s = 0, gi = -f[i];
                for (int k = 0; k < nv; k++) { s += d->M[i][k] * p[k]; gi += d->M[i][k] * a[k]; }
                g0 += p[i] * gi; h0 += p[i] * s;
            }
            for (int r = 0; r < nr; r++) { double s = 0; for (int k = 0; k < nv; k++) s += J[r][k] * p[k]; Jp[r] = s; }
            /* breakpoints of the piecewise-linear derivative */
            double bp[MAXROW + 2];
            int nbp = 0;
            bp[nbp++] = 0.0;
            for (int r = 0; r < nr; r++) if (Jp[r] != 0.0) { double t = -jar[r] / Jp[r]; if (t > 0.0) bp[nbp++] = t; }
            for (int i = 1; i < nbp; i++) { double x = bp[i]; int k = i - 1; while (k >= 0 && bp[k] > x) { bp[k + 1] = bp[k]; k--; } bp[k + 1] = x; }
            double tstar = 1.0;
            for (int s = 0; s < nbp; s++) {
                const double lo = bp[s], hi = (s + 1 < nbp) ? bp[s + 1] : INFINITY;
                const double mid = isinf(hi) ? lo + 1.0 : 0.5 * (lo + hi);
                double c0 = g0, c1 = h0;
                for (int r = 0; r < nr; r++)
                    if (jar[r] + mid * Jp[r] < 0.0) { c0 += D[r] * jar[r] * Jp[r]; c1 += D[r] * Jp[r] * Jp[r]; }
                const double t = -c0 / c1;
                if (t <= hi || s + 1 == nbp) { tstar = t < lo ? lo : t; break; }
            }
            for (int i = 0; i < nv; i++) a[i] += tstar * p[i];
        }
        for (int r = 0; r < nr; r++) {
            double s = -aref[r];
            for (int k = 0; k < nv; k++) s += J[r][k] * a[k];
            d->efc_force[r] = s < 0.0 ? -D[r] * s : 0.0;
            if (s < 0.0) for (int k = 0; k < nv; k++) fc[k] += J[r][k] * (-D[r] * s);
        }
    }
    for (int r = 0; r < nr; r++) {
        d->efc_aref[r] = aref[r]; d->efc_D[r] = D[r];
        for (int k = 0; k < nv; k++) d->efc_J[r][k] = J[r][k];
    }
    for (int k = 0; k < nv; k++) d->qfrc_constraint[k] = fc[k];
    /* --- mj_Euler: implicit in joint damping --- */
    double A[MAXV][MAXV], qa[MAXV];
    for (int i = 0; i < nv; i++) {
        qa[i] = f[i] + fc[i];
        for (int j = 0; j < nv; j++) A[i][j] = d->M[i][j];
        A[i][i] += h * m->damping[i];
    }
    chol_solve(nv, A, qa);
    for (int j = 0; j < nv; j++) { d->qacc[j] = qa[j]; v[j] += h * qa[j]; q[j] += h * v[j]; }
}

/* ------------------------------------------------------------------ C API */
ora_model* ora_model_create(int nb, int nv, const int* parent, const double* pos, const double* mass,
                            const double* ipos, const double* inertia, const int* jnt_body,
                            const double* jnt_axis, const double* jnt_range, const int* jnt_limited,
                            const double* armature, const double* damping, const double* gear,
                            const double* ctrlrange, const double* invweight0, double timestep,
                            int frame_skip, double solK, double solB, const double* solimp,
                            int hand_body, const double* hand_pos, int con_body, const double* con_pos,
                            double con_radius, double con_plane_z, double con_margin, double con_invweight) {
    if (nb > MAXB || nv > MAXV) return NULL;
    ora_model* m = (ora_model*)calloc(1, sizeof(ora_model));
    m->nb = nb; m->nv = nv;
    for (int b = 0; b < nb; b++) {
        m->parent[b] = parent[b]; m->mass[b] = mass[b]; m->body_dof[b] = -1;
        for (int k = 0; k < 3; k++) { m->pos[b][k] = pos[3 * b + k]; m->ipos[b][k] = ipos[3 * b + k]; }
        for (int k = 0; k < 9; k++) m->inertia[b][k] = inertia[9 * b + k];
    }
    for (int j = 0; j < nv; j++) {
        m->jnt_body[j] = jnt_body[j]; m->jnt_limited[j] = jnt_limited[j]; m->body_dof[jnt_body[j]] = j;
        for (int k = 0; k < 3; k++) m->jnt_axis[j][k] = jnt_axis[3 * j + k];
        m->jnt_range[j][0] = jnt_range[2 * j]; m->jnt_range[j][1] = jnt_range[2 * j + 1];
        m->armature[j] = armature[j]; m->damping[j] = damping[j]; m->gear[j] = gear[j];
        m->ctrlrange[j][0] = ctrlrange[2 * j]; m->ctrlrange[j][1] = ctrlrange[2 * j + 1];
        m->invweight0[j] = invweight0[j];
    }
    m->timestep = timestep; m->frame_skip = frame_skip; m->solK = solK; m->solB = solB;
    for (int k = 0; k < 5; k++) m->solimp[k] = solimp[k];
    m->hand_body = hand_body; m->con_body = con_body;
    for (int k = 0; k < 3; k++) { m->hand_pos[k] = hand_pos[k]; m->con_pos[k] = con_pos[k]; }
    m->con_radius = con_radius; m->con_plane_z = con_plane_z; m->con_margin = con_margin;
    m->con_invweight = con_invweight;
    return m;
}
void ora_model_destroy(ora_model* m) { free(m); }

/* M(q) and bias(q,v) for cross-checks */
void ora_mass_bias(const ora_model* m, const double* q, const double* v, double* M_out, double* bias_out,
                   double* hand_out) {
    ora_data d;
    double qq[MAXV], vv[MAXV], u[MAXV] = {0};
    for (int j = 0; j < m->nv; j++) { qq[j] = q[j]; vv[j] = v[j]; }
    ora_step(m, &d, qq, vv, u);
    for (int i = 0; i < m->nv; i++) { bias_out[i] = d.bias[i]; for (int j = 0; j < m->nv; j++) M_out[i * m->nv + j] = d.M[i][j]; }
    for (int k = 0; k < 3; k++) hand_out[k] = d.hand[k];
}

/* one mj_step, state in/out; returns number of active constraint rows */
int ora_substep(const ora_model* m, double* q, double* v, const double* u, double* qacc_out) {
    ora_data d;
    ora_step(m, &d, q, v, u);
    if (qacc_out) for (int j = 0; j < m->nv; j++) qacc_out[j] = d.qacc[j];
    return d.nefc;
}

/* one mj_step, also handing back the constraint rows: J (nefc, nv), aref, D, efc_force (nefc each; room for
 * 2 nv + 1 rows) and qfrc_constraint (nv).  Returns nefc. */
int ora_substep_efc(const ora_model* m, double* q, double* v, const double* u, double* qacc_out, double* J_out,
                    double* aref_out, double* D_out, double* force_out, double* qfrc_out) {
    ora_data d;
    ora_step(m, &d, q, v, u);
    for (int j = 0; j < m->nv; j++) { qacc_out[j] = d.qacc[j]; qfrc_out[j] = d.qfrc_constraint[j]; }
    for (int r = 0; r < d.nefc; r++) {
        aref_out[r] = d.efc_aref[r]; D_out[r] = d.efc_D[r]; force_out[r] = d.efc_force[r];
        for (int k = 0; k < m->nv; k++) J_out[r * m->nv + k] = d.efc_J[r][k];
    }
    return d.nefc;
}

/*
 * The reference rollout (gym_env_wrapper.py:125-153 around reacher_env.py:29-39).
 *   models[n_models]: particle k uses models[k / (K / n_models)] (reference: one model per worker,
 *                     contiguous particle blocks, subproc_vec_env.py:161-168)
 *   state: qpos(nv) qvel(nv) target(3)                    mean (H,nv), noise (K,H,nv) or NULL
 *   costs (K,H) = -reward; actions (K,H,nv) unclipped; qv_traj (K,H,2nv) state after each env step
 *   (NULL ok); next_obs (K,H,2nv+6) (NULL ok); ncon (K,) number of substeps with >=1 active row.
 */
typedef struct {
    ora_model* const* models; int n_models; const double *qpos, *qvel, *target; int K, H;
    const double *mean, *noise; double *costs, *actions, *qv_traj, *next_obs; int* ncon; int k0, k1;
    const double* policy_w;   /* mode="closed_loop_linear": (2nv+6+1, nv) weights; `mean` is then unused */
} ora_job;

static void* ora_rollout_range(void* arg) {
    const ora_job* J = (const ora_job*)arg;
    const int per = J->K / J->n_models, H = J->H;
    for (int k = J->k0; k < J->k1; k++) {
        const ora_model* m = J->models[k / per];
        const int nv = m->nv;
        ora_data d;
        double q[MAXV], v[MAXV], u[MAXV];
        int nc = 0;
        for (int j = 0; j < nv; j++) { q[j] = J->qpos[j]; v[j] = J->qvel[j]; }
        /* closed loop (gym_env_wrapper.py:129-136): curr_obs = get_obs() at the set state -- set_env_state ends
         * with sim.forward() (reacher_env.py:88-99), so the first hand position is fresh -- then
         * curr_obs = next_obs, whose hand position is the stale one of the step's last forward pass */
        double obs[2 * MAXV + 6];
        const int nobs = 2 * nv + 6;
        if (J->policy_w) {
            double qq[MAXV], vv[MAXV], zu[MAXV] = {0};
            for (int j = 0; j < nv; j++) { qq[j] = q[j]; vv[j] = v[j]; }
            ora_step(m, &d, qq, vv, zu);          /* d.hand: kinematics at (q, v) before the integration */
            for (int j = 0; j < nv; j++) { obs[j] = q[j]; obs[nv + j] = v[j]; }
            for (int c = 0; c < 3; c++) { obs[2 * nv + c] = d.hand[c]; obs[2 * nv + 3 + c] = d.hand[c] - J->target[c]; }
        }
        for (int t = 0; t < H; t++) {
            for (int j = 0; j < nv; j++) {
                double mu;
                if (J->policy_w) {                /* mean.T @ np.append(curr_obs, 1.0) */
                    mu = J->policy_w[nobs * nv + j];
                    for (int i = 0; i < nobs; i++) mu += J->policy_w[i * nv + j] * obs[i];
                } else mu = J->mean[t * nv + j];
                u[j] = mu + (J->noise ? J->noise[((size_t)k * H + t) * nv + j] : 0.0);
                if (J->actions) J->actions[((size_t)k * H + t) * nv + j] = u[j];
            }
            for (int s = 0; s < m->frame_skip; s++) { ora_step(m, &d, q, v, u); nc += d.nefc > 0; }
            /* site_xpos is from the last mj_forward, i.e. one integration stale (reacher_env.py:31-35) */
            const double dx = d.hand[0] - J->target[0], dy = d.hand[1] - J->target[1], dz = d.hand[2] - J->target[2];
            const double l1 = fabs(dx) + fabs(dy) + fabs(dz), l2 = sqrt(dx * dx + dy * dy + dz * dz);
            J->costs[(size_t)k * H + t] = l1 + 5.0 * l2;
            if (J->qv_traj) for (int j = 0; j < nv; j++) {
                J->qv_traj[((size_t)k * H + t) * 2 * nv + j] = q[j];
                J->qv_traj[((size_t)k * H + t) * 2 * nv + nv + j] = v[j];
            }
            if (J->next_obs) {
                double* o = J->next_obs + ((size_t)k * H + t) * (2 * nv + 6);
                for (int j = 0; j < nv; j++) { o[j] = q[j]; o[nv + j] = v[j]; }
                o[2 * nv] = d.hand[0]; o[2 * nv + 1] = d.hand[1]; o[2 * nv + 2] = d.hand[2];
                o[2 * nv + 3] = dx; o[2 * nv + 4] = dy; o[2 * nv + 5] = dz;
            }
            if (J->policy_w) {
                for (int j = 0; j < nv; j++) { obs[j] = q[j]; obs[nv + j] = v[j]; }
                obs[2 * nv] = d.hand[0]; obs[2 * nv + 1] = d.hand[1]; obs[2 * nv + 2] = d.hand[2];
                obs[2 * nv + 3] = dx; obs[2 * nv + 4] = dy; obs[2 * nv + 5] = dz;
            }
        }
        if (J->ncon) J->ncon[k] = nc;
    }
    return NULL;
}

void ora_rollout_cl(ora_model* const* models, int n_models, const double* qpos, const double* qvel,
                    const double* target, int K, int H, const double* mean, const double* noise,
                    double* costs, double* actions, double* qv_traj, double* next_obs, int* ncon, int nthreads,
                    const double* policy_w);
void ora_rollout(ora_model* const* models, int n_models, const double* qpos, const double* qvel,
                 const double* target, int K, int H, const double* mean, const double* noise,
                 double* costs, double* actions, double* qv_traj, double* next_obs, int* ncon, int nthreads) {
    ora_rollout_cl(models, n_models, qpos, qvel, target, K, H, mean, noise, costs, actions, qv_traj, next_obs, ncon,
                   nthreads, NULL);
}
void ora_rollout_cl(ora_model* const* models, int n_models, const double* qpos, const double* qvel,
                    const double* target, int K, int H, const double* mean, const double* noise,
                    double* costs, double* actions, double* qv_traj, double* next_obs, int* ncon, int nthreads,
                    const double* policy_w) {
    /* contiguous particle blocks per worker, like the reference's SubprocVecEnv (subproc_vec_env.py:161-168) */
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if (nthreads > K) nthreads = K;
    ora_job jobs[256];
    pthread_t th[256];
    for (int i = 0; i < nthreads; i++) {
        ora_job j = {models, n_models, qpos, qvel, target, K, H, mean, noise, costs, actions, qv_traj, next_obs, ncon,
                     (int)((long long)K * i / nthreads), (int)((long long)K * (i + 1) / nthreads), policy_w};
        jobs[i] = j;
    }
    if (nthreads == 1) { ora_rollout_range(&jobs[0]); return; }
    for (int i = 0; i < nthreads; i++) pthread_create(&th[i], NULL, ora_rollout_range, &jobs[i]);
    for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
}

/* Pendulum (reference: mjmpc/envs/basic/pendulum.py:33-50), K particles, H steps. */
void ora_pendulum_rollout(double th0, double thdot0, int K, int H, const double* mean, const double* noise,
                          double* costs, double* actions, double* states) {
    const double g = 10.0, mm = 1.0, l = 1.0, dt = 0.05, max_speed = 8.0, max_torque = 2.0;
    for (int k = 0; k < K; k++) {
        double th = th0, thdot = thdot0;
        for (int t = 0; t < H; t++) {
            const double a = mean[t] + (noise ? noise[(size_t)k * H + t] : 0.0);
            if (actions) actions[(size_t)k * H + t] = a;
            double u = a < -max_torque ? -max_torque : (a > max_torque ? max_torque : a);
            /* angle_normalize: python modulo has the sign of the divisor */
            double x = fmod(th + M_PI, 2.0 * M_PI);
            if (x < 0) x += 2.0 * M_PI;
            x -= M_PI;
            costs[(size_t)k * H + t] = x * x + .1 * (thdot * thdot) + .001 * (u * u);
            double nthdot = thdot + (-3 * g / (2 * l) * sin(th + M_PI) + 3. / (mm * (l * l)) * u) * dt;
            th = th + nthdot * dt;
            thdot = nthdot < -max_speed ? -max_speed : (nthdot > max_speed ? max_speed : nthdot);
            if (states) { states[((size_t)k * H + t) * 2] = th; states[((size_t)k * H + t) * 2 + 1] = thdot; }
        }
    }
}
