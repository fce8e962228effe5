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
 * Contacts (tree_model_set_contacts; restated from the published engine source of MuJoCo 2.1 -- mjc_PlaneCapsule,
 * mjc_CapsuleCapsule / mjraw_SphereSphere, mju_makeFrame, mj_instantiateContact, mj_diagApprox, mj_makeImpedance -- whose
 * algorithms the 2.0 binary shares): plane against the two end spheres of a capsule (first tangent along the capsule),
 * capsule against capsule (closest points by sequential clamping, then sphere-sphere; parallel axes: ONE contact at the
 * clamped projection where MuJoCo makes up to two); condim 3, pyramidal cone: rows n +- mu t1, n +- mu t2 with
 * diagApprox = (1 + mu^2)(invweight_1 + invweight_2), R = 2 mu^2 (1 - imp)/imp diagApprox for all four (impratio 1).
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define TMAXB 12
#define TMAXV 12
#define TMAXC 24                 /* candidate contact pairs */
#define TMAXROW (TMAXV + 8 * TMAXC)
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
    /* candidate contact pairs (mjcf_tree.compile_mjcf(..., allow_contacts="model")): kind 0 capsule / sphere against
     * capsule / sphere (segments a0-a1, b0-b1 in their body frames), kind 1 world plane (point a0, normal a1) against the
     * two end spheres of segment b0-b1 */
    int ncon, ckind[TMAXC], cbody1[TMAXC], cbody2[TMAXC];
    double ca0[TMAXC][3], ca1[TMAXC][3], cra[TMAXC], cb0[TMAXC][3], cb1[TMAXC][3], crb[TMAXC];
    double cmu[TMAXC], cK[TMAXC], cB[TMAXC], csolimp[TMAXC][5], cinvw[TMAXC];
} tree_model;

typedef struct {
    double M[TMAXV][TMAXV], bias[TMAXV], passive[TMAXV], actuation[TMAXV], constraint[TMAXV], qacc[TMAXV];
    double xpos[TMAXB][3], xmat[TMAXB][9];
    int nefc, ncontact;
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
    /* --- constraint rows (dense): joint limits J = +-e_j, then four pyramid rows per contact --- */
    int nr = 0;
    static __thread double J[TMAXROW][TMAXV];
    double aref[TMAXROW], D[TMAXROW];
    for (int j = 0; j < nv; j++) {
        if (!m->jlimited[j]) continue;
        for (int side = -1; side <= 1; side += 2) {
            const double dist = side * (m->jrange[j][(side + 1) / 2] - q[j]);
            if (dist < 0.0) {
                const double imp = impedance(m->solimp[j], dist);
                double R = (1.0 - imp) * m->invweight0[j] / imp;
                if (R < MJMINVAL) R = MJMINVAL;
                for (int k = 0; k < nv; k++) J[nr][k] = 0.0;
                J[nr][j] = -side; D[nr] = 1.0 / R;
                aref[nr] = -m->solB[j] * (-side * v[j]) - m->solK[j] * imp * dist;
                nr++;
            }
        }
    }
    d->ncontact = 0;
    for (int c = 0; c < m->ncon; c++) {
        /* up to two contacts per candidate: world positions of the sphere centres that touch, normal from geom 1 to geom 2 */
        double cpos[2][3], cnrm[2][3], cdist[2], tpref[3] = {0, 0, 0};
        int nc = 0;
        const int b1 = m->cbody1[c], b2 = m->cbody2[c];
        double q0[3], q1[3], t[3];
        mv3(t, d->xmat[b2], m->cb0[c]); for (int k = 0; k < 3; k++) q0[k] = d->xpos[b2][k] + t[k];
        mv3(t, d->xmat[b2], m->cb1[c]); for (int k = 0; k < 3; k++) q1[k] = d->xpos[b2][k] + t[k];
        if (m->ckind[c] == 1) {
            /* mjc_PlaneCapsule: a sphere test at each end; the first tangent follows the capsule axis */
            const double* n = m->ca1[c];
            for (int k = 0; k < 3; k++) tpref[k] = q1[k] - q0[k];
            for (int e = 0; e < 2; e++) {
                const double* p = e ? q0 : q1;               /* MuJoCo tests pos + axis first */
                double rel[3];
                for (int k = 0; k < 3; k++) rel[k] = p[k] - m->ca0[c][k];
                const double dist = dot3(rel, n) - m->crb[c];
                if (dist < 0.0) {
                    for (int k = 0; k < 3; k++) { cnrm[nc][k] = n[k]; cpos[nc][k] = p[k] - n[k] * (m->crb[c] + 0.5 * dist); }
                    cdist[nc++] = dist;
                }
                if (m->cb0[c][0] == m->cb1[c][0] && m->cb0[c][1] == m->cb1[c][1] && m->cb0[c][2] == m->cb1[c][2]) break;  /* a sphere */
            }
        } else {
            /* mjc_CapsuleCapsule (general position) / mjc_SphereCapsule: closest points of the two segments, then sphere-sphere */
            double p0[3], p1[3];
            mv3(t, d->xmat[b1], m->ca0[c]); for (int k = 0; k < 3; k++) p0[k] = d->xpos[b1][k] + t[k];
            mv3(t, d->xmat[b1], m->ca1[c]); for (int k = 0; k < 3; k++) p1[k] = d->xpos[b1][k] + t[k];
            double c1[3], c2[3], ax1[3], ax2[3], dif[3];
            for (int k = 0; k < 3; k++) {
                c1[k] = 0.5 * (p0[k] + p1[k]); c2[k] = 0.5 * (q0[k] + q1[k]);
                ax1[k] = 0.5 * (p1[k] - p0[k]); ax2[k] = 0.5 * (q1[k] - q0[k]); dif[k] = c1[k] - c2[k];
            }
            const double ma = dot3(ax1, ax1), mb = -dot3(ax1, ax2), mc = dot3(ax2, ax2), u = -dot3(ax1, dif), w = dot3(ax2, dif);
            const double det = ma * mc - mb * mb;
            double x1 = 0.0, x2 = 0.0;
            if (fabs(det) >= MJMINVAL) {
                x1 = (mc * u - mb * w) / det; x2 = (ma * w - mb * u) / det;
                if (x1 > 1) { x1 = 1; x2 = (w - mb) / mc; } else if (x1 < -1) { x1 = -1; x2 = (w + mb) / mc; }
                if (x2 > 1) { x2 = 1; x1 = (u - mb) / ma; if (x1 > 1) x1 = 1; else if (x1 < -1) x1 = -1; }
                else if (x2 < -1) { x2 = -1; x1 = (u + mb) / ma; if (x1 > 1) x1 = 1; else if (x1 < -1) x1 = -1; }
            } else {
                /* a sphere against a segment, or parallel segments (one contact at the clamped projection of the centres) */
                if (ma > MJMINVAL) { x1 = u / ma; if (x1 > 1) x1 = 1; else if (x1 < -1) x1 = -1; }
                if (mc > MJMINVAL) { x2 = (w - mb * x1) / mc; if (x2 > 1) x2 = 1; else if (x2 < -1) x2 = -1; }
            }
            double v1[3], v2[3], dd[3];
            for (int k = 0; k < 3; k++) { v1[k] = c1[k] + ax1[k] * x1; v2[k] = c2[k] + ax2[k] * x2; dd[k] = v2[k] - v1[k]; }
            const double cd = sqrt(dot3(dd, dd)), dist = cd - m->cra[c] - m->crb[c];
            if (dist < 0.0 && cd > MJMINVAL) {
                for (int k = 0; k < 3; k++) { cnrm[0][k] = dd[k] / cd; cpos[0][k] = v1[k] + cnrm[0][k] * (m->cra[c] + 0.5 * dist); }
                cdist[0] = dist; nc = 1;
            }
        }
        for (int e = 0; e < nc; e++) {
            const double* n = cnrm[e];
            /* contact frame: normal, first tangent from the preferred axis (plane-capsule) or mju_makeFrame's default */
            double t1[3], t2[3], y[3] = {0, 0, 0};
            double pref = dot3(tpref, tpref);
            if (pref > MJMINVAL) { for (int k = 0; k < 3; k++) y[k] = tpref[k]; }
            else if (n[1] > -0.5 && n[1] < 0.5) y[1] = 1.0; else y[2] = 1.0;
            double yn = dot3(y, n);
            for (int k = 0; k < 3; k++) t1[k] = y[k] - yn * n[k];
            double tn = sqrt(dot3(t1, t1));
            if (tn < 1e-12) {            /* preferred axis along the normal: fall back to the default rule */
                y[0] = y[1] = y[2] = 0.0;
                if (n[1] > -0.5 && n[1] < 0.5) y[1] = 1.0; else y[2] = 1.0;
                yn = dot3(y, n);
                for (int k = 0; k < 3; k++) t1[k] = y[k] - yn * n[k];
                tn = sqrt(dot3(t1, t1));
            }
            for (int k = 0; k < 3; k++) t1[k] /= tn;
            cross3(t2, n, t1);
            /* Jacobian difference of the contact point: body 2 minus body 1 */
            double Jn[TMAXV], Jt1[TMAXV], Jt2[TMAXV];
            for (int j = 0; j < nv; j++) {
                double col[3] = {0, 0, 0};
                for (int s2 = 0; s2 < 2; s2++) {
                    const int body = s2 ? b2 : b1;
                    if (body < 0) continue;
                    int anc = m->jbody[j] == body;
                    for (int a = m->parent[body]; a >= 0 && !anc; a = m->parent[a]) if (a == m->jbody[j]) anc = 1;
                    if (!anc) continue;
                    double jc[3];
                    if (m->jtype[j] == 1) { for (int k = 0; k < 3; k++) jc[k] = axis[j][k]; }
                    else { double rr[3]; for (int k = 0; k < 3; k++) rr[k] = cpos[e][k] - anchor[j][k]; cross3(jc, axis[j], rr); }
                    for (int k = 0; k < 3; k++) col[k] += s2 ? jc[k] : -jc[k];
                }
                Jn[j] = dot3(n, col); Jt1[j] = dot3(t1, col); Jt2[j] = dot3(t2, col);
            }
            const double mu = m->cmu[c], imp = impedance(m->csolimp[c], cdist[e]);
            double R0 = (1.0 - imp) * ((1.0 + mu * mu) * m->cinvw[c]) / imp;
            if (R0 < MJMINVAL) R0 = MJMINVAL;
            const double Rpy = 2.0 * mu * mu * R0;
            for (int k = 0; k < 4; k++) {
                const double sg = (k & 1) ? -mu : mu;
                const double* Jt = k < 2 ? Jt1 : Jt2;
                double vel = 0.0;
                for (int j = 0; j < nv; j++) { J[nr][j] = Jn[j] + sg * Jt[j]; vel += J[nr][j] * v[j]; }
                D[nr] = 1.0 / Rpy;
                aref[nr] = -m->cB[c] * vel - m->cK[c] * imp * cdist[e];
                nr++;
            }
            d->ncontact++;
        }
    }
    d->nefc = nr;
    double fc[TMAXV];
    for (int j = 0; j < nv; j++) fc[j] = 0.0;
    if (nr > 0) {
        /* exact minimiser of the convex piecewise-quadratic constraint problem: Newton with an exact line search */
        double a[TMAXV], A[TMAXV][TMAXV];
        for (int i = 0; i < nv; i++) { a[i] = f[i]; for (int j = 0; j < nv; j++) A[i][j] = d->M[i][j]; }
        chol_solve(nv, A, a);
        for (int iter = 0; iter < 200; iter++) {
            double jar[TMAXROW], g[TMAXV], fn = 0, gn = 0;
            int act[TMAXROW];
            for (int r = 0; r < nr; r++) {
                double s2 = -aref[r];
                for (int k = 0; k < nv; k++) s2 += J[r][k] * a[k];
                jar[r] = s2; act[r] = s2 < 0.0;
            }
            for (int i = 0; i < nv; i++) {
                double s2 = -f[i];
                for (int k = 0; k < nv; k++) s2 += d->M[i][k] * a[k];
                for (int r = 0; r < nr; r++) if (act[r]) s2 += D[r] * jar[r] * J[r][i];
                g[i] = s2; gn += s2 * s2; fn += f[i] * f[i];
            }
            if (sqrt(gn) <= 1e-15 * (1.0 + sqrt(fn))) break;
            double p[TMAXV], Jp[TMAXROW];
            for (int i = 0; i < nv; i++) {
                p[i] = -g[i];
                for (int j = 0; j < nv; j++) {
                    double s2 = d->M[i][j];
                    for (int r = 0; r < nr; r++) if (act[r]) s2 += D[r] * J[r][i] * J[r][j];
                    A[i][j] = s2;
                }
            }
            chol_solve(nv, A, p);
            double g0 = 0, h0 = 0;
            for (int i = 0; i < nv; i++) {
                double s2 = 0, gi = -f[i];
                for (int k = 0; k < nv; k++) { s2 += d->M[i][k] * p[k]; gi += d->M[i][k] * a[k]; }
                g0 += p[i] * gi; h0 += p[i] * s2;
            }
            for (int r = 0; r < nr; r++) { double s2 = 0; for (int k = 0; k < nv; k++) s2 += J[r][k] * p[k]; Jp[r] = s2; }
            double bp[TMAXROW + 2];
            int nbp = 0;
            bp[nbp++] = 0.0;
            for (int r = 0; r < nr; r++) if (Jp[r] != 0.0) { double tb = -jar[r] / Jp[r]; if (tb > 0.0) bp[nbp++] = tb; }
            for (int i = 1; i < nbp; i++) { double x = bp[i]; int k = i - 1; while (k >= 0 && bp[k] > x) { bp[k + 1] = bp[k]; k--; } bp[k + 1] = x; }
            double tstar = 1.0;
            for (int s2 = 0; s2 < nbp; s2++) {
                const double lo = bp[s2], hi = (s2 + 1 < nbp) ? bp[s2 + 1] : INFINITY;
                const double mid = isinf(hi) ? lo + 1.0 : 0.5 * (lo + hi);
                double c0 = g0, c1 = h0;
                for (int r = 0; r < nr; r++)
                    if (jar[r] + mid * Jp[r] < 0.0) { c0 += D[r] * jar[r] * Jp[r]; c1 += D[r] * Jp[r] * Jp[r]; }
                const double tt = -c0 / c1;
                if (tt <= hi || s2 + 1 == nbp) { tstar = tt < lo ? lo : tt; break; }
            }
            for (int i = 0; i < nv; i++) a[i] += tstar * p[i];
        }
        for (int r = 0; r < nr; r++) {
            double s2 = -aref[r];
            for (int k = 0; k < nv; k++) s2 += J[r][k] * a[k];
            if (s2 < 0.0) for (int k = 0; k < nv; k++) fc[k] += J[r][k] * (-D[r] * s2);
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
int tree_model_set_contacts(tree_model* m, int ncon, const int* kind, const int* body1, const int* body2, const double* a0,
                            const double* a1, const double* ra, const double* b0, const double* b1, const double* rb,
                            const double* mu, const double* K, const double* B, const double* solimp, const double* invw) {
    if (ncon > TMAXC) return -1;
    m->ncon = ncon;
    for (int c = 0; c < ncon; c++) {
        m->ckind[c] = kind[c]; m->cbody1[c] = body1[c]; m->cbody2[c] = body2[c];
        for (int k = 0; k < 3; k++) { m->ca0[c][k] = a0[3 * c + k]; m->ca1[c][k] = a1[3 * c + k]; m->cb0[c][k] = b0[3 * c + k]; m->cb1[c][k] = b1[3 * c + k]; }
        m->cra[c] = ra[c]; m->crb[c] = rb[c]; m->cmu[c] = mu[c]; m->cK[c] = K[c]; m->cB[c] = B[c]; m->cinvw[c] = invw[c];
        for (int k = 0; k < 5; k++) m->csolimp[c][k] = solimp[5 * c + k];
    }
    return 0;
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
