/*
 * oracle/tree_step.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU (FP64, plain C) restatement of MuJoCo 2.0's mj_step for a kinematic TREE of hinge / slide joints -- the
 * checker of csrc/rollout_tree.cu.  Only tests/ may load this library.
 *
 * PARITY UNPINNED AGAINST MuJoCo ITSELF (closed binary, absent from /root/reference and from this image; the
 * reference ships no golden vectors for these models).  Pinned instead against oracle/tree_ref.py -- an independent
 * numpy restatement with its own MJCF reader, Jacobian-sum mass matrix, complex-step Christoffel bias and
 * active-set enumeration (tests/test_tree_oracle.py, tests/golden/tree_pin.npz) -- and against closed forms
 * (terminal velocity of a body in the fluid, pendulum energy).
 *
 * Reference call sites:
 *   mjmpc/envs/basic/swimmer.py:7            MujocoEnv('swimmer.xml', frame_skip 4)
 *   mjmpc/envs/basic/swimmer.py:10-19        do_simulation(a, frame_skip); reward = dx/dt - 1e-4 |a|^2
 *   mjmpc/envs/basic/swimmer.py:21-24        observation = qpos[2:], qvel
 *   mjmpc/envs/basic/half_cheetah.py:10-19   same shape of step / reward (0.1 |a|^2), obs = qpos[1:], qvel
 *   mjmpc/envs/gym_env_wrapper.py:123-153    per-particle reset, u = mean[t] + noise[b,t], unclipped action recorded
 *   mjmpc/envs/assets/xml/swimmer.xml        the model (compiled by mjmpc_b200/envs/mjcf_tree.py)
 *
 * MuJoCo's own formulation is kept (world-orientation spatial vectors about the tree's centre of mass,
 * composite-rigid-body M, recursive Newton-Euler bias with the gravity trick, per-BODY loops over the body's
 * joints), whereas the CUDA kernel works in link coordinates on one link per dof -- independent derivations.
 * Passive forces: joint springs / dampers and the inertia-box fluid model of mj_passive (viscous: -3 pi d mu v,
 * -pi d^3 mu w; quadratic drag: -1/2 rho A |v| v per box face, -rho b (c^4 + d^4)/64 |w| w), evaluated in the body's
 * inertial frame.  Soft joint limits as in oracle/mjstep.c: the convex problem is solved to machine precision.
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define TMAXB 12
#define TMAXV 12
#define MJMINVAL 1e-15
#define MJPI 3.14159265358979323846

typedef struct {
    int nb, nv, nu;
    int parent[TMAXB];
    double pos[TMAXB][3], mat[TMAXB][9], mass[TMAXB], ipos[TMAXB][3], imat[TMAXB][9], inertia[TMAXB][3];
    int jtype[TMAXV], jbody[TMAXV], jlimited[TMAXV];
    double jpos[TMAXV][3], jaxis[TMAXV][3], jrange[TMAXV][2];
    double damping[TMAXV], armature[TMAXV], stiffness[TMAXV], springref[TMAXV], invweight0[TMAXV];
    double solK[TMAXV], solB[TMAXV], solimp[TMAXV][5];
    int act_dof[TMAXV];
    double gear[TMAXV], ctrlrange[TMAXV][2];
    double timestep, gravity[3], density, viscosity;
} tree_model;

typedef struct {
    double M[TMAXV][TMAXV], bias[TMAXV], passive[TMAXV], actuation[TMAXV], constraint[TMAXV], qacc[TMAXV];
    double xpos[TMAXB][3], xmat[TMAXB][9];
    int nefc;
} tree_data;

static void cross3(double* r, const double* a, const double* b) {
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    r[0] = x; r[1] = y; r[2] = z;
}
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void mv3(double* r, const double* M, const double* v) {
    double x = M[0] * v[0] + M[1] * v[1] + M[2] * v[2], y = M[3] * v[0] + M[4] * v[1] + M[5] * v[2];
    double z = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
    r[0] = x; r[1] = y; r[2] = z;
}
static void mtv3(double* r, const double* M, const double* v) { /* M' v */
    double x = M[0] * v[0] + M[3] * v[1] + M[6] * v[2], y = M[1] * v[0] + M[4] * v[1] + M[7] * v[2];
    double z = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
    r[0] = x; r[1] = y; r[2] = z;
}
static void mm3(double* R, const double* A, const double* B) {
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
typedef struct { double m, h[3], I[9]; } sinert;
static void sinert_mul(double* f, const sinert* S, const double* mv) {
    double t[3], u[3];
    mv3(t, S->I, mv);
    cross3(u, S->h, mv + 3);
    f[0] = t[0] + u[0]; f[1] = t[1] + u[1]; f[2] = t[2] + u[2];
    cross3(u, mv, S->h);
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
static void chol_solve(int n, double A[TMAXV][TMAXV], double* b) {
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
static double impedance(const double* solimp, double dist) {
    if (solimp[0] == solimp[1] || solimp[2] <= MJMINVAL) return 0.5 * (solimp[0] + solimp[1]);
    double x = fabs(dist / solimp[2]);
    if (x >= 1.0) return solimp[1];
    if (x <= 0.0) return solimp[0];
    double y;
    if (solimp[4] == 1.0) y = x;
    else if (x <= solimp[3]) y = pow(x, solimp[4]) / pow(solimp[3], solimp[4] - 1.0);
    else y = 1.0 - pow(1.0 - x, solimp[4]) / pow(1.0 - solimp[3], solimp[4] - 1.0);
    return solimp[0] + y * (solimp[1] - solimp[0]);
}

/* One mj_step at (q, v) with controls u (nu), advancing q, v in place. */
static void tree_step(const tree_model* m, tree_data* d, double* q, double* v, const double* u) {
    const int nb = m->nb, nv = m->nv;
    const double h = m->timestep;
    double anchor[TMAXV][3], axis[TMAXV][3];
    const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Z3[3] = {0, 0, 0};
    /* --- mj_kinematics: the joints of a body act in order on the body frame --- */
    for (int b = 0; b < nb; b++) {
        const int pa = m->parent[b];
        const double* Rp = pa < 0 ? I3 : d->xmat[pa];
        const double* Pp = pa < 0 ? Z3 : d->xpos[pa];
        double pos[3], R[9], t[3];
        mv3(t, Rp, m->pos[b]);
        for (int k = 0; k < 3; k++) pos[k] = Pp[k] + t[k];
        mm3(R, Rp, m->mat[b]);
        for (int j = 0; j < nv; j++) {
            if (m->jbody[j] != b) continue;
            mv3(t, R, m->jpos[j]);
            for (int k = 0; k < 3; k++) anchor[j][k] = pos[k] + t[k];
            mv3(axis[j], R, m->jaxis[j]);
            if (m->jtype[j] == 1) {
                for (int k = 0; k < 3; k++) pos[k] += axis[j][k] * q[j];
            } else {
                double Rj[9];
                rot_axis(Rj, m->jaxis[j], q[j]);
                mm3(R, R, Rj);
                mv3(t, R, m->jpos[j]);
                for (int k = 0; k < 3; k++) pos[k] = anchor[j][k] - t[k];
            }
        }
        memcpy(d->xpos[b], pos, sizeof(pos));
        memcpy(d->xmat[b], R, sizeof(R));
    }
    /* --- mj_comPos --- */
    double com[3] = {0, 0, 0}, mtot = 0, xipos[TMAXB][3], ximat[TMAXB][9];
    for (int b = 0; b < nb; b++) {
        double t[3];
        mv3(t, d->xmat[b], m->ipos[b]);
        for (int k = 0; k < 3; k++) { xipos[b][k] = d->xpos[b][k] + t[k]; com[k] += m->mass[b] * xipos[b][k]; }
        mm3(ximat[b], d->xmat[b], m->imat[b]);
        mtot += m->mass[b];
    }
    for (int k = 0; k < 3; k++) com[k] /= mtot;
    sinert cin[TMAXB], crb[TMAXB];
    for (int b = 0; b < nb; b++) {
        double Iw[9], r[3];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                double s = 0;
                for (int k = 0; k < 3; k++) s += ximat[b][3 * i + k] * m->inertia[b][k] * ximat[b][3 * j + k];
                Iw[3 * i + j] = s;
            }
        for (int k = 0; k < 3; k++) r[k] = xipos[b][k] - com[k];
        const double mm = m->mass[b], rr = dot3(r, r);
        cin[b].m = mm;
        for (int k = 0; k < 3; k++) cin[b].h[k] = mm * r[k];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) cin[b].I[3 * i + j] = Iw[3 * i + j] + mm * ((i == j ? rr : 0.0) - r[i] * r[j]);
        crb[b] = cin[b];
    }
    double cdof[TMAXV][6];
    for (int j = 0; j < nv; j++) {
        if (m->jtype[j] == 1) {
            for (int k = 0; k < 3; k++) { cdof[j][k] = 0.0; cdof[j][3 + k] = axis[j][k]; }
        } else {
            double off[3];
            for (int k = 0; k < 3; k++) { cdof[j][k] = axis[j][k]; off[k] = com[k] - anchor[j][k]; }
            cross3(cdof[j] + 3, axis[j], off);
        }
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
    /* dof i is an ancestor-or-self of dof j iff its body is an ancestor of j's body, or the same body and i <= j */
    for (int i = 0; i < nv; i++) for (int j = 0; j < nv; j++) d->M[i][j] = 0.0;
    for (int j = 0; j < nv; j++) {
        double buf[6];
        sinert_mul(buf, &crb[m->jbody[j]], cdof[j]);
        for (int i = 0; i <= j; i++) {
            int anc = 0;
            if (m->jbody[i] == m->jbody[j]) anc = 1;
            else for (int b = m->parent[m->jbody[j]]; b >= 0; b = m->parent[b]) if (b == m->jbody[i]) { anc = 1; break; }
            if (!anc) continue;
            d->M[i][j] = dot6(cdof[i], buf);
            d->M[j][i] = d->M[i][j];
        }
        d->M[j][j] += m->armature[j];
    }
    /* --- mj_comVel + mj_rne: bias = c(q, v) - gravity forces (base acceleration = -gravity) --- */
    double cvel[TMAXB][6], cacc[TMAXB][6], cfrc[TMAXB][6];
    for (int b = 0; b < nb; b++) {
        const int pa = m->parent[b];
        for (int k = 0; k < 6; k++) {
            cvel[b][k] = pa < 0 ? 0.0 : cvel[pa][k];
            cacc[b][k] = pa < 0 ? (k < 3 ? 0.0 : -m->gravity[k - 3]) : cacc[pa][k];
        }
        for (int j = 0; j < nv; j++) {
            if (m->jbody[j] != b) continue;
            double cdofdot[6];
            cross_motion(cdofdot, cvel[b], cdof[j]);      /* with the velocity accumulated so far (MuJoCo's order) */
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
    for (int j = 0; j < nv; j++) d->bias[j] = dot6(cdof[j], cfrc[m->jbody[j]]);
    /* --- mj_passive --- */
    for (int j = 0; j < nv; j++) d->passive[j] = -m->stiffness[j] * (q[j] - m->springref[j]) - m->damping[j] * v[j];
    if (m->viscosity > 0.0 || m->density > 0.0) {
        for (int b = 0; b < nb; b++) {
            if (m->mass[b] < MJMINVAL) continue;
            const double* In = m->inertia[b];
            double box[3];
            box[0] = sqrt(fmax(MJMINVAL, In[1] + In[2] - In[0]) / m->mass[b] * 6.0);
            box[1] = sqrt(fmax(MJMINVAL, In[0] + In[2] - In[1]) / m->mass[b] * 6.0);
            box[2] = sqrt(fmax(MJMINVAL, In[0] + In[1] - In[2]) / m->mass[b] * 6.0);
            /* velocity of the body's centre of mass, inertial-frame coordinates (cvel refers to the tree COM) */
            double w[3], vl[3], r[3], t[3], lw[3], lv[3], lfrc[6] = {0, 0, 0, 0, 0, 0};
            for (int k = 0; k < 3; k++) { w[k] = cvel[b][k]; r[k] = xipos[b][k] - com[k]; }
            cross3(t, w, r);
            for (int k = 0; k < 3; k++) vl[k] = cvel[b][3 + k] + t[k];
            mtv3(lw, ximat[b], w);
            mtv3(lv, ximat[b], vl);
            if (m->viscosity > 0.0) {
                const double diam = (box[0] + box[1] + box[2]) / 3.0;
                for (int k = 0; k < 3; k++) {
                    lfrc[k] = -MJPI * diam * diam * diam * m->viscosity * lw[k];
                    lfrc[3 + k] = -3.0 * MJPI * diam * m->viscosity * lv[k];
                }
            }
            if (m->density > 0.0) {
                lfrc[3] -= 0.5 * m->density * box[1] * box[2] * fabs(lv[0]) * lv[0];
                lfrc[4] -= 0.5 * m->density * box[0] * box[2] * fabs(lv[1]) * lv[1];
                lfrc[5] -= 0.5 * m->density * box[0] * box[1] * fabs(lv[2]) * lv[2];
                lfrc[0] -= m->density * box[0] * (pow(box[1], 4) + pow(box[2], 4)) * fabs(lw[0]) * lw[0] / 64.0;
                lfrc[1] -= m->density * box[1] * (pow(box[0], 4) + pow(box[2], 4)) * fabs(lw[1]) * lw[1] / 64.0;
                lfrc[2] -= m->density * box[2] * (pow(box[0], 4) + pow(box[1], 4)) * fabs(lw[2]) * lw[2] / 64.0;
            }
            double torque[3], force[3];
            mv3(torque, ximat[b], lfrc);
            mv3(force, ximat[b], lfrc + 3);
            /* mj_applyFT at xipos: qfrc += Jp' force + Jr' torque over the dofs above the body */
            for (int j = 0; j < nv; j++) {
                int anc = m->jbody[j] == b;
                for (int a = m->parent[b]; a >= 0 && !anc; a = m->parent[a]) if (a == m->jbody[j]) anc = 1;
                if (!anc) continue;
                if (m->jtype[j] == 1) d->passive[j] += dot3(axis[j], force);
                else {
                    double rr[3], jp[3];
                    for (int k = 0; k < 3; k++) rr[k] = xipos[b][k] - anchor[j][k];
                    cross3(jp, axis[j], rr);
                    d->passive[j] += dot3(jp, force) + dot3(axis[j], torque);
                }
            }
        }
    }
    /* --- mj_fwdActuation --- */
    for (int j = 0; j < nv; j++) d->actuation[j] = 0.0;
    for (int a = 0; a < m->nu; a++) {
        double c = u[a];
        if (c < m->ctrlrange[a][0]) c = m->ctrlrange[a][0];
        if (c > m->ctrlrange[a][1]) c = m->ctrlrange[a][1];
        d->actuation[m->act_dof[a]] += m->gear[a] * c;
    }
    double f[TMAXV];
    for (int j = 0; j < nv; j++) f[j] = d->passive[j] + d->actuation[j] - d->bias[j];
    /* --- joint-limit rows: J = +-e_j --- */
    int nr = 0, rdof[TMAXV];
    double rs[TMAXV], aref[TMAXV], D[TMAXV];
    for (int j = 0; j < nv; j++) {
        if (!m->jlimited[j]) continue;
        for (int side = -1; side <= 1; side += 2) {
            const double dist = side * (m->jrange[j][(side + 1) / 2] - q[j]);
            if (dist < 0.0) {
                const double imp = impedance(m->solimp[j], dist);
                double R = (1.0 - imp) * m->invweight0[j] / imp;
                if (R < MJMINVAL) R = MJMINVAL;
                rdof[nr] = j; rs[nr] = -side; D[nr] = 1.0 / R;
                aref[nr] = -m->solB[j] * (-side * v[j]) - m->solK[j] * imp * dist;
                nr++;
            }
        }
    }
    d->nefc = nr;
    double fc[TMAXV];
    for (int j = 0; j < nv; j++) fc[j] = 0.0;
    if (nr > 0) {
        double a[TMAXV], A[TMAXV][TMAXV];
        for (int i = 0; i < nv; i++) { a[i] = f[i]; for (int j = 0; j < nv; j++) A[i][j] = d->M[i][j]; }
        chol_solve(nv, A, a);
        for (int iter = 0; iter < 100; iter++) {
            double jar[TMAXV], g[TMAXV], fn = 0, gn = 0;
            int act[TMAXV];
            for (int r = 0; r < nr; r++) { jar[r] = rs[r] * a[rdof[r]] - aref[r]; act[r] = jar[r] < 0.0; }
            for (int i = 0; i < nv; i++) {
                double s = -f[i];
                for (int k = 0; k < nv; k++) s += d->M[i][k] * a[k];
                g[i] = s;
            }
            for (int r = 0; r < nr; r++) if (act[r]) g[rdof[r]] += D[r] * jar[r] * rs[r];
            for (int i = 0; i < nv; i++) { gn += g[i] * g[i]; fn += f[i] * f[i]; }
            if (sqrt(gn) <= 1e-15 * (1.0 + sqrt(fn))) break;
            double p[TMAXV], Jp[TMAXV];
            for (int i = 0; i < nv; i++) { p[i] = -g[i]; for (int j = 0; j < nv; j++) A[i][j] = d->M[i][j]; }
            for (int r = 0; r < nr; r++) if (act[r]) A[rdof[r]][rdof[r]] += D[r];
            chol_solve(nv, A, p);
            double g0 = 0, h0 = 0;
            for (int i = 0; i < nv; i++) {
                double s = 0, gi = -f[i];
                for (int k = 0; k < nv; k++) { s += d->M[i][k] * p[k]; gi += d->M[i][k] * a[k]; }
                g0 += p[i] * gi; h0 += p[i] * s;
            }
            for (int r = 0; r < nr; r++) Jp[r] = rs[r] * p[rdof[r]];
            double bp[TMAXV + 2];
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
            const double s = rs[r] * a[rdof[r]] - aref[r];
            if (s < 0.0) fc[rdof[r]] += rs[r] * (-D[r] * s);
        }
    }
    for (int k = 0; k < nv; k++) d->constraint[k] = fc[k];
    /* --- mj_Euler, implicit in joint damping --- */
    double A[TMAXV][TMAXV], qa[TMAXV];
    for (int i = 0; i < nv; i++) {
        qa[i] = f[i] + fc[i];
        for (int j = 0; j < nv; j++) A[i][j] = d->M[i][j];
        A[i][i] += h * m->damping[i];
    }
    chol_solve(nv, A, qa);
    for (int j = 0; j < nv; j++) { d->qacc[j] = qa[j]; v[j] += h * qa[j]; q[j] += h * v[j]; }
}

/* ------------------------------------------------------------------ C API */
tree_model* tree_model_create(int nb, int nv, int nu, const int* parent, const double* pos, const double* mat,
                              const double* mass, const double* ipos, const double* imat, const double* inertia,
                              const int* jtype, const int* jbody, const double* jpos, const double* jaxis,
                              const int* jlimited, const double* jrange, const double* damping, const double* armature,
                              const double* stiffness, const double* springref, const double* invweight0,
                              const double* solK, const double* solB, const double* solimp, const int* act_dof,
                              const double* gear, const double* ctrlrange, double timestep, const double* gravity,
                              double density, double viscosity) {
    if (nb > TMAXB || nv > TMAXV || nu > TMAXV) return NULL;
    tree_model* m = (tree_model*)calloc(1, sizeof(tree_model));
    m->nb = nb; m->nv = nv; m->nu = nu;
    for (int b = 0; b < nb; b++) {
        m->parent[b] = parent[b]; m->mass[b] = mass[b];
        for (int k = 0; k < 3; k++) { m->pos[b][k] = pos[3 * b + k]; m->ipos[b][k] = ipos[3 * b + k]; m->inertia[b][k] = inertia[3 * b + k]; }
        for (int k = 0; k < 9; k++) { m->mat[b][k] = mat[9 * b + k]; m->imat[b][k] = imat[9 * b + k]; }
    }
    for (int j = 0; j < nv; j++) {
        m->jtype[j] = jtype[j]; m->jbody[j] = jbody[j]; m->jlimited[j] = jlimited[j];
        for (int k = 0; k < 3; k++) { m->jpos[j][k] = jpos[3 * j + k]; m->jaxis[j][k] = jaxis[3 * j + k]; }
        m->jrange[j][0] = jrange[2 * j]; m->jrange[j][1] = jrange[2 * j + 1];
        m->damping[j] = damping[j]; m->armature[j] = armature[j]; m->stiffness[j] = stiffness[j];
        m->springref[j] = springref[j]; m->invweight0[j] = invweight0[j]; m->solK[j] = solK[j]; m->solB[j] = solB[j];
        for (int k = 0; k < 5; k++) m->solimp[j][k] = solimp[5 * j + k];
    }
    for (int a = 0; a < nu; a++) {
        m->act_dof[a] = act_dof[a]; m->gear[a] = gear[a];
        m->ctrlrange[a][0] = ctrlrange[2 * a]; m->ctrlrange[a][1] = ctrlrange[2 * a + 1];
    }
    m->timestep = timestep; m->density = density; m->viscosity = viscosity;
    for (int k = 0; k < 3; k++) m->gravity[k] = gravity[k];
    return m;
}
void tree_model_free(tree_model* m) { free(m); }

/* one substep with every intermediate exposed (tests compare them term by term with oracle/tree_ref.py) */
void tree_substep_debug(const tree_model* m, double* q, double* v, const double* u, double* M, double* bias,
                        double* passive, double* actuation, double* constraint, double* qacc, int* nefc) {
    tree_data d;
    tree_step(m, &d, q, v, u);
    for (int i = 0; i < m->nv; i++) {
        for (int j = 0; j < m->nv; j++) M[i * m->nv + j] = d.M[i][j];
        bias[i] = d.bias[i]; passive[i] = d.passive[i]; actuation[i] = d.actuation[i];
        constraint[i] = d.constraint[i]; qacc[i] = d.qacc[i];
    }
    *nefc = d.nefc;
}

/* Rollouts of the forward-progress environments (swimmer.py:10-19, half_cheetah.py:10-19):
 *   reward = w_fwd * (q[fwd_dof] after - before) / (frame_skip * timestep) - w_ctrl * |a|^2,   cost = -reward
 * state0: (K, 2 nv) or (1, 2 nv) broadcast (state_stride 0);  mean (H, nu);  noise (K, H, nu) or NULL.
 * Outputs: costs (K, H), actions (K, H, nu) unclipped, states (K, H, 2 nv) AFTER each env step (may be NULL). */
typedef struct {
    const tree_model* m; int k0, k1, H, frame_skip, fwd_dof, state_stride; double w_fwd, w_ctrl;
    const double *state0, *mean, *noise; double *costs, *actions, *states; int* nefc_total;
} tree_job;
static void* tree_worker(void* arg) {
    tree_job* J = (tree_job*)arg;
    const tree_model* m = J->m;
    const int nv = m->nv, nu = m->nu, H = J->H;
    tree_data d;
    int nefc = 0;
    for (int k = J->k0; k < J->k1; k++) {
        double q[TMAXV], v[TMAXV], u[TMAXV];
        const double* s0 = J->state0 + (size_t)k * J->state_stride;
        for (int j = 0; j < nv; j++) { q[j] = s0[j]; v[j] = s0[nv + j]; }
        for (int t = 0; t < H; t++) {
            double a2 = 0;
            for (int a = 0; a < nu; a++) {
                u[a] = J->mean[t * nu + a] + (J->noise ? J->noise[((size_t)k * H + t) * nu + a] : 0.0);
                J->actions[((size_t)k * H + t) * nu + a] = u[a];
                a2 += u[a] * u[a];
            }
            const double before = q[J->fwd_dof];
            for (int s = 0; s < J->frame_skip; s++) { tree_step(m, &d, q, v, u); nefc += d.nefc; }
            const double reward = J->w_fwd * (q[J->fwd_dof] - before) / (J->frame_skip * m->timestep) - J->w_ctrl * a2;
            J->costs[(size_t)k * H + t] = -reward;
            if (J->states) {
                double* so = J->states + ((size_t)k * H + t) * 2 * nv;
                for (int j = 0; j < nv; j++) { so[j] = q[j]; so[nv + j] = v[j]; }
            }
        }
    }
    *J->nefc_total = nefc;
    return NULL;
}
int tree_rollout(const tree_model* m, int K, int H, int frame_skip, int fwd_dof, double w_fwd, double w_ctrl,
                 const double* state0, int state_stride, const double* mean, const double* noise, double* costs,
                 double* actions, double* states, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    if (nthreads > K) nthreads = K;
    pthread_t th[64];
    tree_job jobs[64];
    int nefc[64];
    for (int i = 0; i < nthreads; i++) {
        tree_job j = {m, (int)((long)K * i / nthreads), (int)((long)K * (i + 1) / nthreads), H, frame_skip, fwd_dof,
                      state_stride, w_fwd, w_ctrl, state0, mean, noise, costs, actions, states, &nefc[i]};
        jobs[i] = j;
        pthread_create(&th[i], NULL, tree_worker, &jobs[i]);
    }
    int total = 0;
    for (int i = 0; i < nthreads; i++) { pthread_join(th[i], NULL); total += nefc[i]; }
    return total;
}
