/*
 * oracle/tg_oracle.c - CPU restatement of the tactile_gym hot path.  See tg_oracle.h for the
 * parity status of each part.  TEST INFRASTRUCTURE ONLY - never linked into the product.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no FMA contraction, so the float32
 * post-process reproduces numpy's arithmetic bit for bit).
 */
#include "tg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ small linear algebra */
typedef double v3[3];
typedef double m3[9]; /* row major */

static void v3set(v3 a, double x, double y, double z) { a[0] = x; a[1] = y; a[2] = z; }
static void v3cpy(v3 a, const v3 b) { a[0] = b[0]; a[1] = b[1]; a[2] = b[2]; }
static void v3add(v3 o, const v3 a, const v3 b) { o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2]; }
static void v3sub(v3 o, const v3 a, const v3 b) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
static void v3scale(v3 o, const v3 a, double s) { o[0] = a[0] * s; o[1] = a[1] * s; o[2] = a[2] * s; }
static void v3axpy(v3 o, double s, const v3 a) { o[0] += a[0] * s; o[1] += a[1] * s; o[2] += a[2] * s; }
static double v3dot(const v3 a, const v3 b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double v3norm(const v3 a) { return sqrt(v3dot(a, a)); }
static void v3cross(v3 o, const v3 a, const v3 b)
{
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
static void m3mulv(v3 o, const m3 M, const v3 a)
{
    double x = M[0] * a[0] + M[1] * a[1] + M[2] * a[2];
    double y = M[3] * a[0] + M[4] * a[1] + M[5] * a[2];
    double z = M[6] * a[0] + M[7] * a[1] + M[8] * a[2];
    o[0] = x; o[1] = y; o[2] = z;
}
static void m3tmulv(v3 o, const m3 M, const v3 a)
{
    double x = M[0] * a[0] + M[3] * a[1] + M[6] * a[2];
    double y = M[1] * a[0] + M[4] * a[1] + M[7] * a[2];
    double z = M[2] * a[0] + M[5] * a[1] + M[8] * a[2];
    o[0] = x; o[1] = y; o[2] = z;
}
static void m3mul(m3 o, const m3 A, const m3 B)
{
    m3 t;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    memcpy(o, t, sizeof(m3));
}
static void m3tmul(m3 o, const m3 A, const m3 B) /* A^T B */
{
    m3 t;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) t[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
    memcpy(o, t, sizeof(m3));
}
static void m3ident(m3 o) { memset(o, 0, sizeof(m3)); o[0] = o[4] = o[8] = 1.0; }

/* URDF fixed-axis rpy -> matrix, R = Rz(yaw) Ry(pitch) Rx(roll) */
static void m3rpy(m3 R, const v3 rpy)
{
    double cr = cos(rpy[0]), sr = sin(rpy[0]), cp = cos(rpy[1]), sp = sin(rpy[1]), cy = cos(rpy[2]), sy = sin(rpy[2]);
    R[0] = cy * cp; R[1] = cy * sp * sr - sy * cr; R[2] = cy * sp * cr + sy * sr;
    R[3] = sy * cp; R[4] = sy * sp * sr + cy * cr; R[5] = sy * sp * cr - cy * sr;
    R[6] = -sp;     R[7] = cp * sr;                R[8] = cp * cr;
}
/* rotation by angle about a (not necessarily unit) axis: bullet builds btQuaternion(axis, angle) */
static void m3axis_angle(m3 R, const v3 axis_in, double ang)
{
    v3 a; double n = v3norm(axis_in);
    v3scale(a, axis_in, n > 0 ? 1.0 / n : 0.0);
    double c = cos(ang), s = sin(ang), t = 1 - c;
    R[0] = t * a[0] * a[0] + c;        R[1] = t * a[0] * a[1] - s * a[2]; R[2] = t * a[0] * a[2] + s * a[1];
    R[3] = t * a[0] * a[1] + s * a[2]; R[4] = t * a[1] * a[1] + c;        R[5] = t * a[1] * a[2] - s * a[0];
    R[6] = t * a[0] * a[2] - s * a[1]; R[7] = t * a[1] * a[2] + s * a[0]; R[8] = t * a[2] * a[2] + c;
}

/* ------------------------------------------------------------------ pybullet frame helpers */
/* p.getQuaternionFromEuler: btQuaternion::setEulerZYX(yaw, pitch, roll); quaternions are [x,y,z,w] */
void or_quat_from_euler(const double rpy[3], double q[4])
{
    double hr = rpy[0] * 0.5, hp = rpy[1] * 0.5, hy = rpy[2] * 0.5;
    double cr = cos(hr), sr = sin(hr), cp = cos(hp), sp = sin(hp), cy = cos(hy), sy = sin(hy);
    q[0] = sr * cp * cy - cr * sp * sy;
    q[1] = cr * sp * cy + sr * cp * sy;
    q[2] = cr * cp * sy - sr * sp * cy;
    q[3] = cr * cp * cy + sr * sp * sy;
}
/* p.getEulerFromQuaternion (bullet getEulerZYX with the +-0.99999 gimbal guard) */
void or_euler_from_quat(const double q[4], double rpy[3])
{
    double sqx = q[0] * q[0], sqy = q[1] * q[1], sqz = q[2] * q[2], squ = q[3] * q[3];
    double sarg = -2.0 * (q[0] * q[2] - q[3] * q[1]);
    if (sarg <= -0.99999) { rpy[1] = -0.5 * M_PI; rpy[0] = 0; rpy[2] = 2 * atan2(q[0], -q[1]); }
    else if (sarg >= 0.99999) { rpy[1] = 0.5 * M_PI; rpy[0] = 0; rpy[2] = 2 * atan2(-q[0], q[1]); }
    else {
        rpy[1] = asin(sarg);
        rpy[0] = atan2(2 * (q[1] * q[2] + q[3] * q[0]), squ - sqx - sqy + sqz);
        rpy[2] = atan2(2 * (q[0] * q[1] + q[3] * q[2]), squ + sqx - sqy - sqz);
    }
}
static void quat_mul(double o[4], const double a[4], const double b[4])
{
    double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    double y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
    double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
void or_mat_from_quat(const double q[4], double R[9])
{
    double d = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    double s = 2.0 / d;
    double xs = q[0] * s, ys = q[1] * s, zs = q[2] * s;
    double wx = q[3] * xs, wy = q[3] * ys, wz = q[3] * zs;
    double xx = q[0] * xs, xy = q[0] * ys, xz = q[0] * zs;
    double yy = q[1] * ys, yz = q[1] * zs, zz = q[2] * zs;
    R[0] = 1 - (yy + zz); R[1] = xy - wz;       R[2] = xz + wy;
    R[3] = xy + wz;       R[4] = 1 - (xx + zz); R[5] = yz - wx;
    R[6] = xz - wy;       R[7] = yz + wx;       R[8] = 1 - (xx + yy);
}
static void quat_from_mat(double q[4], const m3 R) /* btMatrix3x3::getRotation */
{
    double tr = R[0] + R[4] + R[8];
    if (tr > 0) {
        double s = sqrt(tr + 1.0);
        q[3] = s * 0.5; s = 0.5 / s;
        q[0] = (R[7] - R[5]) * s; q[1] = (R[2] - R[6]) * s; q[2] = (R[3] - R[1]) * s;
    } else {
        int i = R[0] < R[4] ? (R[4] < R[8] ? 2 : 1) : (R[0] < R[8] ? 2 : 0);
        int j = (i + 1) % 3, k = (i + 2) % 3;
        double s = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
        q[i] = s * 0.5; s = 0.5 / s;
        q[3] = (R[3 * k + j] - R[3 * j + k]) * s;
        q[j] = (R[3 * j + i] + R[3 * i + j]) * s;
        q[k] = (R[3 * k + i] + R[3 * i + k]) * s;
    }
}
void or_mul_transforms(const double pa[3], const double qa[4], const double pb[3], const double qb[4], double po[3], double qo[4])
{
    m3 R; v3 t;
    or_mat_from_quat(qa, R);
    m3mulv(t, R, pb);
    v3add(po, pa, t);
    quat_mul(qo, qa, qb);
}
void or_invert_transform(const double p[3], const double q[4], double po[3], double qo[4])
{
    m3 R; v3 t;
    double qi[4] = {-q[0], -q[1], -q[2], q[3]};
    or_mat_from_quat(qi, R);
    m3mulv(t, R, p);
    po[0] = -t[0]; po[1] = -t[1]; po[2] = -t[2];
    memcpy(qo, qi, sizeof(qi));
}

/* ------------------------------------------------------------------ kinematics */
typedef struct {
    m3 Rl[OR_MAXL]; v3 pl[OR_MAXL]; /* world pose of the URDF link frames */
    m3 Rc[OR_MAXL]; v3 pc[OR_MAXL]; /* world pose of the inertial (COM) frames */
    v3 aw[OR_MAXL];                 /* joint axis, world, unit */
} Kin;

static void kin_compute(const OrModel* m, const double* q, Kin* k)
{
    for (int i = 0; i < m->nlinks; i++) {
        m3 Rj, Rq, Rp, Rin; v3 pp, t;
        int p = m->parent[i];
        if (p < 0) { m3ident(Rp); v3set(pp, 0, 0, 0); }
        else { memcpy(Rp, k->Rl[p], sizeof(m3)); v3cpy(pp, k->pl[p]); }
        m3rpy(Rj, m->joint_rpy[i]);
        m3mulv(t, Rp, m->joint_xyz[i]);
        v3add(k->pl[i], pp, t);
        m3mul(k->Rl[i], Rp, Rj);
        if (m->jtype[i] == 1) {
            m3axis_angle(Rq, m->axis[i], q[m->dof_of_link[i]]);
            m3mul(k->Rl[i], k->Rl[i], Rq);
        }
        double n = v3norm(m->axis[i]);
        v3 au; v3scale(au, m->axis[i], n > 0 ? 1 / n : 0);
        m3mulv(k->aw[i], k->Rl[i], au);
        m3rpy(Rin, m->inertial_rpy[i]);
        m3mul(k->Rc[i], k->Rl[i], Rin);
        m3mulv(t, k->Rl[i], m->inertial_xyz[i]);
        v3add(k->pc[i], k->pl[i], t);
    }
}

void or_link_states(const OrModel* m, const double* q, double pos[][3], double quat[][4])
{
    Kin k; kin_compute(m, q, &k);
    for (int i = 0; i < m->nlinks; i++) { v3cpy(pos[i], k.pc[i]); quat_from_mat(quat[i], k.Rc[i]); }
}

void or_link_frames(const OrModel* m, const double* q, double pos[][3], double R[][9])
{
    Kin k; kin_compute(m, q, &k);
    for (int i = 0; i < m->nlinks; i++) { v3cpy(pos[i], k.pl[i]); memcpy(R[i], k.Rl[i], sizeof(m3)); }
}

void or_jacobian(const OrModel* m, const double* q, int link, double J[6][OR_MAXD])
{
    Kin k; kin_compute(m, q, &k);
    for (int r = 0; r < 6; r++) for (int c = 0; c < OR_MAXD; c++) J[r][c] = 0;
    for (int i = link; i >= 0; i = m->parent[i]) {
        if (m->jtype[i] != 1) continue;
        int d = m->dof_of_link[i];
        v3 r, lin; v3sub(r, k.pc[link], k.pl[i]); v3cross(lin, k.aw[i], r);
        for (int c = 0; c < 3; c++) { J[c][d] = lin[c]; J[3 + c][d] = k.aw[i][c]; }
    }
}

void or_link_velocity(const OrModel* m, const double* q, const double* qd, int link, double lin[3], double ang[3])
{
    double J[6][OR_MAXD];
    or_jacobian(m, q, link, J);
    for (int c = 0; c < 3; c++) {
        lin[c] = ang[c] = 0;
        for (int d = 0; d < m->ndof; d++) { lin[c] += J[c][d] * qd[d]; ang[c] += J[3 + c][d] * qd[d]; }
    }
}

/* pb.calculateInverseDynamics: recursive Newton-Euler in the world frame, gravity as a base acceleration.
 * No damping terms (bullet's inverse-dynamics tree does not model them). */
void or_inverse_dynamics(const OrModel* m, const double* q, const double* qd, const double* qdd, double* tau)
{
    Kin k; kin_compute(m, q, &k);
    v3 w[OR_MAXL], al[OR_MAXL], ap[OR_MAXL]; /* ang vel, ang acc, lin acc of the link-frame origin */
    v3 F[OR_MAXL], N[OR_MAXL];               /* accumulated force, and moment about the link-frame origin */
    for (int i = 0; i < m->nlinks; i++) {
        int p = m->parent[i];
        v3 wp, alp, app, pp;
        if (p < 0) { v3set(wp, 0, 0, 0); v3set(alp, 0, 0, 0); v3scale(app, m->gravity, -1.0); v3set(pp, 0, 0, 0); }
        else { v3cpy(wp, w[p]); v3cpy(alp, al[p]); v3cpy(app, ap[p]); v3cpy(pp, k.pl[p]); }
        v3 r, t, t2;
        v3sub(r, k.pl[i], pp);
        /* a_pivot = a_p + alpha_p x r + w_p x (w_p x r) */
        v3cross(t, alp, r); v3add(ap[i], app, t);
        v3cross(t, wp, r); v3cross(t2, wp, t); v3add(ap[i], ap[i], t2);
        v3cpy(w[i], wp); v3cpy(al[i], alp);
        if (m->jtype[i] == 1) {
            int d = m->dof_of_link[i];
            v3 jv; v3scale(jv, k.aw[i], qd[d]);
            v3add(w[i], wp, jv);
            v3axpy(al[i], qdd ? qdd[d] : 0.0, k.aw[i]);
            v3cross(t, wp, jv); v3add(al[i], al[i], t);
        }
        /* COM acceleration */
        v3 rc, ac; v3sub(rc, k.pc[i], k.pl[i]);
        v3cross(t, al[i], rc); v3add(ac, ap[i], t);
        v3cross(t, w[i], rc); v3cross(t2, w[i], t); v3add(ac, ac, t2);
        v3scale(F[i], ac, m->mass[i]);
        /* N_com = I alpha + w x I w, I = Rc diag Rc^T */
        v3 wl, all, Iw, Ial, nl, nw;
        m3tmulv(wl, k.Rc[i], w[i]); m3tmulv(all, k.Rc[i], al[i]);
        for (int c = 0; c < 3; c++) { Iw[c] = m->inertia[i][c] * wl[c]; Ial[c] = m->inertia[i][c] * all[c]; }
        v3cross(nl, wl, Iw); v3add(nl, nl, Ial);
        m3mulv(nw, k.Rc[i], nl);
        v3cross(t, rc, F[i]); v3add(N[i], nw, t);
    }
    for (int i = m->nlinks - 1; i >= 0; i--) {
        if (m->jtype[i] == 1) tau[m->dof_of_link[i]] = v3dot(k.aw[i], N[i]);
        int p = m->parent[i];
        if (p >= 0) {
            v3 r, t; v3sub(r, k.pl[i], k.pl[p]);
            v3cross(t, r, F[i]);
            v3add(N[p], N[p], N[i]); v3add(N[p], N[p], t);
            v3add(F[p], F[p], F[i]);
        }
    }
}

/* ------------------------------------------------------------------ spatial algebra (COM-local frames) */
/* spatial vectors: [angular(3), linear(3)]; forces: [torque(3), force(3)] */
typedef double sv[6];
typedef double sm[36];

typedef struct {
    m3 E[OR_MAXL]; /* rotation parent COM frame -> this COM frame */
    v3 r[OR_MAXL]; /* this COM minus parent COM, in PARENT COM coordinates */
    sv S[OR_MAXL]; /* motion subspace in this COM frame */
    m3 Rc[OR_MAXL];
    sv v[OR_MAXL], c[OR_MAXL];
    sm IA[OR_MAXL];
    sv pA[OR_MAXL], U[OR_MAXL];
    double Dinv[OR_MAXL], u[OR_MAXL];
} Aba;

static void x_motion(sv o, const m3 E, const v3 r, const sv a) /* parent -> child */
{
    v3 t, lin;
    m3mulv(o, E, a);
    v3cross(t, a, r);       /* w x r */
    v3add(lin, a + 3, t);   /* v + w x r */
    m3mulv(o + 3, E, lin);
}
static void xt_force_acc(sv acc, const m3 E, const v3 r, const sv f) /* child -> parent, accumulate */
{
    v3 fp, np, t;
    m3tmulv(fp, E, f + 3);
    m3tmulv(np, E, f);
    v3cross(t, r, fp);
    for (int c = 0; c < 3; c++) { acc[c] += np[c] + t[c]; acc[3 + c] += fp[c]; }
}
static void x_matrix(sm X, const m3 E, const v3 r) /* 6x6 motion transform parent -> child */
{
    /* [ E 0 ; -E rx  E ] with (w x r) = -rx w */
    m3 rx = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0}, Erx;
    m3mul(Erx, E, rx);
    memset(X, 0, sizeof(sm));
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            X[6 * i + j] = E[3 * i + j];
            X[6 * (i + 3) + j + 3] = E[3 * i + j];
            X[6 * (i + 3) + j] = -Erx[3 * i + j];
        }
}
static void sm_mulv(sv o, const sm M, const sv a)
{
    sv t;
    for (int i = 0; i < 6; i++) { t[i] = 0; for (int j = 0; j < 6; j++) t[i] += M[6 * i + j] * a[j]; }
    memcpy(o, t, sizeof(sv));
}
static double sv_dot(const sv a, const sv b) { double s = 0; for (int i = 0; i < 6; i++) s += a[i] * b[i]; return s; }
static void cross_motion(sv o, const sv v, const sv mvec) /* v x m */
{
    v3 a, b, c;
    v3cross(a, v, mvec);
    v3cross(b, v, mvec + 3);
    v3cross(c, v + 3, mvec);
    v3cpy(o, a); v3add(o + 3, b, c);
}

/* First ABA sweep + articulated inertias.  Restates btMultiBody::computeAccelerationsArticulatedBodyAlgorithmMultiDof:
 * link quantities live in each link's COM frame, gravity enters as a link force, per-link velocity damping
 * f = -m v (k + k|v|), n = -I w (k + k|w|) is folded into the zero-acceleration force. */
static void aba_setup(const OrModel* m, const double* q, const double* qd, const double* tau, int with_vel_terms, Aba* A)
{
    Kin k; kin_compute(m, q, &k);
    for (int i = 0; i < m->nlinks; i++) {
        int p = m->parent[i];
        m3 Rp; v3 pp, dw, Rin_axis, d;
        if (p < 0) { m3ident(Rp); v3set(pp, 0, 0, 0); } else { memcpy(Rp, k.Rc[p], sizeof(m3)); v3cpy(pp, k.pc[p]); }
        memcpy(A->Rc[i], k.Rc[i], sizeof(m3));
        m3tmul(A->E[i], k.Rc[i], Rp);
        v3sub(dw, k.pc[i], pp);
        m3tmulv(A->r[i], Rp, dw);
        memset(A->S[i], 0, sizeof(sv));
        if (m->jtype[i] == 1) {
            m3tmulv(Rin_axis, k.Rc[i], k.aw[i]);            /* axis in COM frame */
            v3sub(dw, k.pc[i], k.pl[i]); m3tmulv(d, k.Rc[i], dw); /* pivot -> COM in COM frame */
            v3cpy(A->S[i], Rin_axis);
            v3cross(A->S[i] + 3, Rin_axis, d);
        }
        /* velocities */
        sv vp = {0, 0, 0, 0, 0, 0};
        if (p >= 0) memcpy(vp, A->v[p], sizeof(sv));
        x_motion(A->v[i], A->E[i], A->r[i], vp);
        sv vj = {0, 0, 0, 0, 0, 0};
        if (m->jtype[i] == 1 && with_vel_terms) for (int c = 0; c < 6; c++) vj[c] = A->S[i][c] * qd[m->dof_of_link[i]];
        if (!with_vel_terms) memset(A->v[i], 0, sizeof(sv));
        for (int c = 0; c < 6; c++) A->v[i][c] += vj[c];
        cross_motion(A->c[i], A->v[i], vj);
        /* isolated spatial inertia at COM */
        memset(A->IA[i], 0, sizeof(sm));
        for (int c = 0; c < 3; c++) { A->IA[i][7 * c] = m->inertia[i][c]; A->IA[i][7 * (c + 3)] = m->mass[i]; }
        /* zero-acceleration force */
        v3 gl, Iw, t;
        m3tmulv(gl, k.Rc[i], m->gravity);
        for (int c = 0; c < 3; c++) { A->pA[i][c] = 0; A->pA[i][3 + c] = -m->mass[i] * gl[c]; }
        if (with_vel_terms) {
            const double* w = A->v[i]; const double* vl = A->v[i] + 3;
            for (int c = 0; c < 3; c++) Iw[c] = m->inertia[i][c] * w[c];
            v3cross(t, w, Iw); for (int c = 0; c < 3; c++) A->pA[i][c] += t[c];
            v3cross(t, w, vl); for (int c = 0; c < 3; c++) A->pA[i][3 + c] += m->mass[i] * t[c];
            double ka = m->ang_damping * (1.0 + v3norm(w)), kl = m->lin_damping * (1.0 + v3norm(vl));
            for (int c = 0; c < 3; c++) { A->pA[i][c] += Iw[c] * ka; A->pA[i][3 + c] += m->mass[i] * vl[c] * kl; }
        } else {
            memset(A->pA[i], 0, sizeof(sv)); /* delta solves: no bias */
        }
    }
    /* tips -> base */
    for (int i = m->nlinks - 1; i >= 0; i--) {
        int p = m->parent[i];
        sm Ia; sv pa;
        memcpy(Ia, A->IA[i], sizeof(sm)); memcpy(pa, A->pA[i], sizeof(sv));
        if (m->jtype[i] == 1) {
            int d = m->dof_of_link[i];
            sm_mulv(A->U[i], A->IA[i], A->S[i]);
            double D = sv_dot(A->S[i], A->U[i]);
            A->Dinv[i] = 1.0 / D;
            A->u[i] = (tau ? tau[d] : 0.0) - sv_dot(A->S[i], A->pA[i]);
            for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) Ia[6 * a + b] -= A->U[i][a] * A->Dinv[i] * A->U[i][b];
            sv Iac; sm_mulv(Iac, Ia, A->c[i]);
            for (int a = 0; a < 6; a++) pa[a] += Iac[a] + A->U[i][a] * A->Dinv[i] * A->u[i];
        } else {
            sv Iac; sm_mulv(Iac, Ia, A->c[i]);
            for (int a = 0; a < 6; a++) pa[a] += Iac[a];
        }
        if (p >= 0) {
            sm X, T;
            x_matrix(X, A->E[i], A->r[i]);
            /* IA[p] += X^T Ia X */
            for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) { double s = 0; for (int c = 0; c < 6; c++) s += Ia[6 * a + c] * X[6 * c + b]; T[6 * a + b] = s; }
            for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) { double s = 0; for (int c = 0; c < 6; c++) s += X[6 * c + a] * T[6 * c + b]; A->IA[p][6 * a + b] += s; }
            xt_force_acc(A->pA[p], A->E[i], A->r[i], pa);
        }
    }
}

static void aba_accel(const OrModel* m, const Aba* A, double* qdd)
{
    sv a[OR_MAXL];
    for (int i = 0; i < m->nlinks; i++) {
        int p = m->parent[i];
        sv ap = {0, 0, 0, 0, 0, 0};
        if (p >= 0) memcpy(ap, a[p], sizeof(sv));
        x_motion(a[i], A->E[i], A->r[i], ap);
        for (int c = 0; c < 6; c++) a[i][c] += A->c[i][c];
        if (m->jtype[i] == 1) {
            double qa = A->Dinv[i] * (A->u[i] - sv_dot(A->U[i], a[i]));
            qdd[m->dof_of_link[i]] = qa;
            for (int c = 0; c < 6; c++) a[i][c] += A->S[i][c] * qa;
        }
    }
}

/* btMultiBody::calcAccelerationDeltasMultiDof: response qdd = M^-1 f to a generalized force, reusing U, D */
static void aba_delta(const OrModel* m, const Aba* A, const double* f, double* out)
{
    sv p[OR_MAXL], a[OR_MAXL]; double u[OR_MAXL];
    memset(p, 0, sizeof(p));
    for (int i = m->nlinks - 1; i >= 0; i--) {
        sv pa; memcpy(pa, p[i], sizeof(sv));
        if (m->jtype[i] == 1) {
            u[i] = f[m->dof_of_link[i]] - sv_dot(A->S[i], p[i]);
            for (int c = 0; c < 6; c++) pa[c] += A->U[i][c] * A->Dinv[i] * u[i];
        }
        if (m->parent[i] >= 0) xt_force_acc(p[m->parent[i]], A->E[i], A->r[i], pa);
    }
    for (int i = 0; i < m->nlinks; i++) {
        int pi = m->parent[i];
        sv ap = {0, 0, 0, 0, 0, 0};
        if (pi >= 0) memcpy(ap, a[pi], sizeof(sv));
        x_motion(a[i], A->E[i], A->r[i], ap);
        if (m->jtype[i] == 1) {
            double qa = A->Dinv[i] * (u[i] - sv_dot(A->U[i], a[i]));
            out[m->dof_of_link[i]] = qa;
            for (int c = 0; c < 6; c++) a[i][c] += A->S[i][c] * qa;
        }
    }
}

void or_mass_matrix_inverse(const OrModel* m, const double* q, double Minv[OR_MAXD][OR_MAXD])
{
    static double zero[OR_MAXD];
    Aba A; aba_setup(m, q, zero, NULL, 0, &A);
    for (int j = 0; j < m->ndof; j++) {
        double f[OR_MAXD] = {0}, o[OR_MAXD];
        f[j] = 1.0;
        aba_delta(m, &A, f, o);
        for (int i = 0; i < m->ndof; i++) Minv[i][j] = o[i];
    }
}

void or_forward_dynamics(const OrModel* m, const double* q, const double* qd, const double* tau, int with_damping, double* qdd)
{
    OrModel mm = *m;
    if (!with_damping) { mm.lin_damping = 0; mm.ang_damping = 0; }
    Aba A; aba_setup(&mm, q, qd, tau, 1, &A);
    aba_accel(&mm, &A, qdd);
}

/* ------------------------------------------------------------------ stepSimulation
 * [EXT] btMultiBodyDynamicsWorld::internalSingleStepSimulation for one fixed-base multibody with
 * joint motors only:
 *   1. joint damping torque -= jointDamping * qd  (PhysicsServer applies it before each step)
 *   2. ABA (gravity, gyroscopic, link damping, applied torques); qd += dt * qdd
 *   3. one constraint row per motor: J = e_i, response = M^-1 e_i, target velocity
 *        rhs_v = kp * erp(=1) * (pos_target - q)/dt + qd + kd * (vel_target - qd),  |impulse| <= force * dt
 *   4. projected Gauss-Seidel, at most numSolverIterations sweeps; the sweep direction alternates
 *      (even iterations back to front); after each sweep the solver stops if the largest squared
 *      velocity-level change of any row, (deltaImpulse / jacDiagABInv)^2, is <= solverResidualThreshold
 *      (pybullet's default 1e-7, PyBullet Quickstart Guide, setPhysicsEngineParameter; the reference
 *      does not override it)
 *   5. qd += delta; q += dt * qd
 * Call site: robots/arms/robot.py:141; parameters rl_envs/base_tactile_env.py:127-130.
 */
/* diagnostics (tools/sweep_stats.py): PGS sweeps executed and calls, summed since the library was loaded */
long long or_dbg_sweeps = 0, or_dbg_calls = 0;

void or_step_simulation(const OrModel* m, OrState* s, const double* tau_applied)
{
    int n = m->ndof;
    double tau[OR_MAXD] = {0}, qdd[OR_MAXD];
    for (int i = 0; i < n; i++) tau[i] = (tau_applied ? tau_applied[i] : 0.0) - m->joint_damping * s->qd[i];
    Aba A; aba_setup(m, s->q, s->qd, tau, 1, &A);
    aba_accel(m, &A, qdd);
    for (int i = 0; i < n; i++) s->qd[i] += m->dt * qdd[i];

    /* constraint rows */
    double resp[OR_MAXD][OR_MAXD], diaginv[OR_MAXD], rhs[OR_MAXD], lim[OR_MAXD], applied[OR_MAXD], dv[OR_MAXD];
    int rows[OR_MAXD], nrows = 0;
    for (int i = 0; i < n; i++) {
        double maximp = s->max_force[i] * m->dt;
        if (maximp == 0) continue;
        double f[OR_MAXD] = {0}; f[i] = 1.0;
        aba_delta(m, &A, f, resp[nrows]);
        double denom = resp[nrows][i];
        diaginv[nrows] = denom > 2.2204460492503131e-16 ? 1.0 / denom : 0.0;
        double v = s->qd[i];
        double kp = s->motor_mode[i] == 1 ? s->kp[i] : 0.0;
        double tp = s->motor_mode[i] == 1 ? s->target_pos[i] : 0.0;
        double pos_stab = 1.0 * (tp - s->q[i]) / m->dt;
        double rhs_v = kp * pos_stab + v + s->kd[i] * (s->target_vel[i] - v);
        rhs[nrows] = (rhs_v - v) * diaginv[nrows];
        lim[nrows] = maximp; applied[nrows] = 0; rows[nrows] = i;
        nrows++;
    }
    for (int i = 0; i < n; i++) dv[i] = 0;
    for (int it = 0; it < m->solver_iters; it++) {
        double resid = 0;
        for (int jj = 0; jj < nrows; jj++) {
            int r = (it & 1) ? jj : nrows - 1 - jj;
            int d = rows[r];
            double delta = rhs[r] - dv[d] * diaginv[r];
            double sum = applied[r] + delta;
            if (sum < -lim[r]) { delta = -lim[r] - applied[r]; applied[r] = -lim[r]; }
            else if (sum > lim[r]) { delta = lim[r] - applied[r]; applied[r] = lim[r]; }
            else applied[r] = sum;
            for (int i = 0; i < n; i++) dv[i] += resp[r][i] * delta;
            double dvel = diaginv[r] != 0 ? delta / diaginv[r] : 0.0;
            if (dvel * dvel > resid) resid = dvel * dvel;
        }
        or_dbg_sweeps++;
        if (resid <= m->solver_residual_threshold) break;
    }
    or_dbg_calls++;
    for (int i = 0; i < n; i++) { s->qd[i] += dv[i]; s->q[i] += m->dt * s->qd[i]; }
}

/* ------------------------------------------------------------------ stepSimulation with a free body + P2P constraint
 * Same pipeline as or_step_simulation; the solver now sees 6 + 3 rows.  Rows keep creation order (motors were created
 * with the robot, the constraint later, object_balance_env.py:110), sweeps alternate direction, same residual exit. */
static void quat_rotate(const double q[4], const double v[3], double o[3]) { m3 R; or_mat_from_quat(q, R); m3mulv(o, R, v); }

void or_step_sim_obj(const OrModel* m, OrState* s, OrObject* o)
{
    int n = m->ndof;
    double tau_gc[OR_MAXD], tau[OR_MAXD] = {0}, qdd[OR_MAXD];
    or_inverse_dynamics(m, s->q, s->qd, NULL, tau_gc);
    for (int i = 0; i < n; i++) tau[i] = tau_gc[i] - m->joint_damping * s->qd[i];
    Aba A; aba_setup(m, s->q, s->qd, tau, 1, &A);
    aba_accel(m, &A, qdd);
    for (int i = 0; i < n; i++) s->qd[i] += m->dt * qdd[i];

    /* object: unconstrained update about the composite COM (gravity, one-step external force, gyroscopic torque; no damping) */
    m3 Rb, Iw; v3 dw, cw, vc;
    double Iinv[3];
    or_mat_from_quat(o->quat, Rb);
    m3mulv(dw, Rb, o->com_off);
    v3add(cw, o->pos, dw);
    { v3 t; v3cross(t, o->omg, dw); v3add(vc, o->vel, t); }
    for (int c = 0; c < 3; c++) Iinv[c] = 1.0 / o->inertia[c];
    {
        v3 F = {m->gravity[0] * o->mass, m->gravity[1] * o->mass, m->gravity[2] * o->mass}, T = {0, 0, 0};
        if (o->ext_pending) {
            v3 r, t;
            v3sub(r, o->ext_pos, cw); v3cross(t, r, o->ext_force);
            v3add(F, F, o->ext_force); v3add(T, T, t);
            o->ext_pending = 0;
        }
        v3 wl, Iwv, gy, Tl, al, aw;
        m3tmulv(wl, Rb, o->omg);
        for (int c = 0; c < 3; c++) Iwv[c] = o->inertia[c] * wl[c];
        v3cross(gy, wl, Iwv);
        m3tmulv(Tl, Rb, T);
        for (int c = 0; c < 3; c++) al[c] = (Tl[c] - gy[c]) * Iinv[c];
        m3mulv(aw, Rb, al);
        for (int c = 0; c < 3; c++) { vc[c] += m->dt * F[c] / o->mass; o->omg[c] += m->dt * aw[c]; }
        (void)Iw;
    }

    /* rows */
    enum { MAXR = OR_MAXD + 3 };
    double jr[MAXR][OR_MAXD], ur[MAXR][OR_MAXD], jbl[MAXR][3], jba[MAXR][3], ubl[MAXR][3], uba[MAXR][3];
    double diaginv[MAXR], rhs[MAXR], lim[MAXR], applied[MAXR];
    int nrows = 0;
    for (int i = 0; i < n; i++) {
        double maximp = s->max_force[i] * m->dt;
        if (maximp == 0) continue;
        double f[OR_MAXD] = {0}; f[i] = 1.0;
        for (int d = 0; d < n; d++) jr[nrows][d] = f[d];
        aba_delta(m, &A, f, ur[nrows]);
        for (int c = 0; c < 3; c++) { jbl[nrows][c] = jba[nrows][c] = ubl[nrows][c] = uba[nrows][c] = 0; }
        double denom = ur[nrows][i];
        diaginv[nrows] = denom > 2.2204460492503131e-16 ? 1.0 / denom : 0.0;
        double v = s->qd[i];
        double kp = s->motor_mode[i] == 1 ? s->kp[i] : 0.0, tp = s->motor_mode[i] == 1 ? s->target_pos[i] : 0.0;
        double rhs_v = kp * ((tp - s->q[i]) / m->dt) + v + s->kd[i] * (s->target_vel[i] - v);
        rhs[nrows] = (rhs_v - v) * diaginv[nrows];
        lim[nrows] = maximp; applied[nrows] = 0;
        nrows++;
    }
    if (o->p2p_enabled) {
        double P[OR_MAXL][3], Q[OR_MAXL][4], J[6][OR_MAXD];
        or_link_states(m, s->q, P, Q);
        or_jacobian(m, s->q, m->tcp_link, J);
        v3 pa, pb, rb, t;
        v3cpy(pa, P[m->tcp_link]);
        quat_rotate(o->quat, o->pivot_b, t); v3add(pb, o->pos, t);
        v3sub(rb, pb, cw);
        for (int i = 0; i < 3; i++) {
            int r = nrows;
            v3 nA = {0, 0, 0}; nA[i] = -1.0;     /* on the arm: -e_i; on the object: +e_i */
            double f[OR_MAXD];
            for (int d = 0; d < n; d++) { jr[r][d] = -J[i][d]; f[d] = jr[r][d]; }
            aba_delta(m, &A, f, ur[r]);
            v3 nB = {0, 0, 0}; nB[i] = 1.0;
            v3cpy(jbl[r], nB); v3cross(jba[r], rb, nB);
            for (int c = 0; c < 3; c++) ubl[r][c] = jbl[r][c] / o->mass;
            { v3 jl, ul; m3tmulv(jl, Rb, jba[r]); for (int c = 0; c < 3; c++) ul[c] = jl[c] * Iinv[c]; m3mulv(uba[r], Rb, ul); }
            double denom = 0;
            for (int d = 0; d < n; d++) denom += jr[r][d] * ur[r][d];
            denom += v3dot(jbl[r], ubl[r]) + v3dot(jba[r], uba[r]);
            diaginv[r] = denom > 2.2204460492503131e-16 ? 1.0 / denom : 0.0;
            double rel = 0;
            for (int d = 0; d < n; d++) rel += jr[r][d] * s->qd[d];
            rel += v3dot(jbl[r], vc) + v3dot(jba[r], o->omg);
            double pos_error = (pa[i] - pb[i]) * nA[i];              /* (pivotA - pivotB) . normal */
            double positional = -pos_error * o->erp / m->dt;
            rhs[r] = (positional + (0.0 - rel)) * diaginv[r];
            lim[r] = o->max_impulse; applied[r] = 0;
            nrows++;
        }
    }
    double dv[OR_MAXD] = {0}, dvl[3] = {0, 0, 0}, dva[3] = {0, 0, 0};
    for (int it = 0; it < m->solver_iters; it++) {
        double resid = 0;
        for (int jj = 0; jj < nrows; jj++) {
            int r = (it & 1) ? jj : nrows - 1 - jj;
            double dot = 0;
            for (int d = 0; d < n; d++) dot += jr[r][d] * dv[d];
            dot += v3dot(jbl[r], dvl) + v3dot(jba[r], dva);
            double delta = rhs[r] - dot * diaginv[r];
            double sum = applied[r] + delta;
            if (sum < -lim[r]) { delta = -lim[r] - applied[r]; applied[r] = -lim[r]; }
            else if (sum > lim[r]) { delta = lim[r] - applied[r]; applied[r] = lim[r]; }
            else applied[r] = sum;
            for (int d = 0; d < n; d++) dv[d] += ur[r][d] * delta;
            for (int c = 0; c < 3; c++) { dvl[c] += ubl[r][c] * delta; dva[c] += uba[r][c] * delta; }
            double dvel = diaginv[r] != 0 ? delta / diaginv[r] : 0.0;
            if (dvel * dvel > resid) resid = dvel * dvel;
        }
        if (resid <= m->solver_residual_threshold) break;
    }
    for (int i = 0; i < n; i++) { s->qd[i] += dv[i]; s->q[i] += m->dt * s->qd[i]; }
    /* object: apply, integrate (exponential map on the orientation), convert back to the base-link COM */
    for (int c = 0; c < 3; c++) { vc[c] += dvl[c]; o->omg[c] += dva[c]; }
    v3 cnew;
    for (int c = 0; c < 3; c++) cnew[c] = cw[c] + m->dt * vc[c];
    {
        double wn = v3norm(o->omg), ang = wn * m->dt;
        double dq[4] = {0, 0, 0, 1};
        if (wn > 1e-300) { double sn = sin(0.5 * ang) / wn; dq[0] = o->omg[0] * sn; dq[1] = o->omg[1] * sn; dq[2] = o->omg[2] * sn; dq[3] = cos(0.5 * ang); }
        double qn[4]; quat_mul(qn, dq, o->quat);
        double nn = sqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
        for (int c = 0; c < 4; c++) o->quat[c] = qn[c] / nn;
    }
    quat_rotate(o->quat, o->com_off, dw);
    v3sub(o->pos, cnew, dw);
    { v3 t; v3cross(t, o->omg, dw); v3sub(o->vel, vc, t); }
}

/* Robot.step_sim (robot.py:131-141): gravity compensation (base_robot_arm.py:174-189) then stepSimulation */
void or_step_sim(const OrModel* m, OrState* s)
{
    double tau[OR_MAXD];
    or_inverse_dynamics(m, s->q, s->qd, NULL, tau);
    or_step_simulation(m, s, tau);
}

/* ------------------------------------------------------------------ control */
static void workframe_quat(const OrModel* m, double q[4]) { or_quat_from_euler(m->workframe_rpy, q); }

/* get_current_TCP_pos_vel_workframe pose part (base_robot_arm.py:153-172, worldframe_to_workframe :62-74) */
void or_tcp_pose_workframe(const OrModel* m, const double* q, double pos[3], double rpy[3])
{
    double P[OR_MAXL][3], Q[OR_MAXL][4], rpy_w[3], qw[4], wq[4], ip[3], iq[4], oq[4];
    or_link_states(m, q, P, Q);
    or_euler_from_quat(Q[m->tcp_link], rpy_w);
    or_quat_from_euler(rpy_w, qw);
    workframe_quat(m, wq);
    or_invert_transform(m->workframe_pos, wq, ip, iq);
    or_mul_transforms(ip, iq, P[m->tcp_link], qw, pos, oq);
    or_euler_from_quat(oq, rpy);
}

/* solve A x = b (n x n) by LU with partial pivoting; returns smallest |pivot| / largest |pivot| */
static double lu_solve(int n, double A[][OR_MAXD], double* b, double* x)
{
    double M[OR_MAXD][OR_MAXD + 1];
    double pmin = 1e300, pmax = 0;
    for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) M[i][j] = A[i][j]; M[i][n] = b[i]; }
    for (int c = 0; c < n; c++) {
        int piv = c;
        for (int r = c + 1; r < n; r++) if (fabs(M[r][c]) > fabs(M[piv][c])) piv = r;
        if (piv != c) for (int j = 0; j <= n; j++) { double t = M[c][j]; M[c][j] = M[piv][j]; M[piv][j] = t; }
        double p = fabs(M[c][c]);
        if (p < pmin) pmin = p;
        if (p > pmax) pmax = p;
        if (p == 0) continue;
        for (int r = c + 1; r < n; r++) {
            double f = M[r][c] / M[c][c];
            for (int j = c; j <= n; j++) M[r][j] -= f * M[c][j];
        }
    }
    for (int r = n - 1; r >= 0; r--) {
        double sacc = M[r][n];
        for (int j = r + 1; j < n; j++) sacc -= M[r][j] * x[j];
        x[r] = M[r][r] != 0 ? sacc / M[r][r] : 0;
    }
    return pmax > 0 ? pmin / pmax : 0;
}

/* x = pinv(J) v for a 6 x n Jacobian via one-sided Jacobi SVD of J^T (n x 6), rcond = 1e-15 like numpy */
static void pinv_apply(int n, double J[6][OR_MAXD], const double v[6], double* x)
{
    /* work on W = J^T (n rows, 6 cols); orthogonalise columns: W V = U S */
    double W[OR_MAXD][6], V[6][6];
    for (int i = 0; i < n; i++) for (int j = 0; j < 6; j++) W[i][j] = J[j][i];
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) V[i][j] = i == j;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0;
        for (int p = 0; p < 5; p++) for (int qq = p + 1; qq < 6; qq++) {
            double a = 0, b = 0, c = 0;
            for (int i = 0; i < n; i++) { a += W[i][p] * W[i][p]; b += W[i][qq] * W[i][qq]; c += W[i][p] * W[i][qq]; }
            if (fabs(c) <= 1e-300 || fabs(c) <= 1e-17 * sqrt(a * b)) continue;
            off += fabs(c);
            double zeta = (b - a) / (2 * c);
            double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1 + zeta * zeta));
            double cs = 1 / sqrt(1 + t * t), sn = cs * t;
            for (int i = 0; i < n; i++) { double wp = W[i][p], wq = W[i][qq]; W[i][p] = cs * wp - sn * wq; W[i][qq] = sn * wp + cs * wq; }
            for (int i = 0; i < 6; i++) { double vp = V[i][p], vq = V[i][qq]; V[i][p] = cs * vp - sn * vq; V[i][qq] = sn * vp + cs * vq; }
        }
        if (off == 0) break;
    }
    /* J^T = U S V^T  =>  J = V S U^T  => pinv(J) = U S^-1 V^T ; x = sum_k U_k (V_k . v) / s_k */
    double sig[6], smax = 0;
    for (int k = 0; k < 6; k++) { double a = 0; for (int i = 0; i < n; i++) a += W[i][k] * W[i][k]; sig[k] = sqrt(a); if (sig[k] > smax) smax = sig[k]; }
    for (int i = 0; i < n; i++) x[i] = 0;
    for (int k = 0; k < 6; k++) {
        if (sig[k] <= 1e-15 * smax) continue;
        double vk = 0; for (int j = 0; j < 6; j++) vk += V[j][k] * v[j];
        for (int i = 0; i < n; i++) x[i] += (W[i][k] / sig[k]) * vk / sig[k];
    }
}

void or_tcp_velocity_control(const OrModel* m, OrState* s, const double vels_work[6])
{
    /* check_TCP_vel_lims (base_robot_arm.py:357-380) */
    double pos[3], rpy[3], v[6];
    or_tcp_pose_workframe(m, s->q, pos, rpy);
    for (int i = 0; i < 6; i++) {
        double cur = i < 3 ? pos[i] : rpy[i - 3];
        int ex = (cur < m->tcp_lims[i][0] && vels_work[i] < 0) || (cur > m->tcp_lims[i][1] && vels_work[i] > 0);
        v[i] = ex ? 0.0 : vels_work[i];
    }
    /* workvel_to_worldvel (:96-105) */
    double wq[4], R[9], vw[6];
    workframe_quat(m, wq); or_mat_from_quat(wq, R);
    m3mulv(vw, R, v); m3mulv(vw + 3, R, v + 3);
    /* Jacobian at the TCP link, inverse or pseudo-inverse (:296-322) */
    double J[6][OR_MAXD], qd_t[OR_MAXD];
    or_jacobian(m, s->q, m->tcp_link, J);
    int n = m->ndof, use_pinv = (n != 6) || m->mg400_slave;
    if (!use_pinv) {
        double A[OR_MAXD][OR_MAXD];
        for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) A[i][j] = J[i][j];
        double rc = lu_solve(6, A, vw, qd_t);
        if (rc < 1e-13) use_pinv = 1; /* numpy matrix_rank < 6  ->  pinv branch */
    }
    if (use_pinv) pinv_apply(n, J, vw, qd_t);
    if (m->mg400_slave) { /* mg400.py:111-120 */
        qd_t[n - 3] = qd_t[1]; qd_t[n - 2] = -qd_t[1]; qd_t[n - 1] = qd_t[1] + qd_t[2];
    }
    for (int i = 0; i < n; i++) {
        s->motor_mode[i] = 0; s->target_vel[i] = qd_t[i]; s->kd[i] = m->vel_gain; s->kp[i] = 0; s->max_force[i] = m->max_force;
    }
}

void or_apply_action(const OrModel* m, OrState* s, const double vels_work[6], int repeat)
{
    or_tcp_velocity_control(m, s, vels_work);
    for (int i = 0; i < repeat; i++) or_step_sim(m, s);
}

/* [EXT] pybullet IK, IK2_VEL_DLS_WITH_ORIENTATION without null space: repeat up to 100 times
 *   dq = (J^T J + diag(0.5))^-1 J^T e,  clamped so max|dq| <= 45 deg,
 * e = [target_pos - pos ; axis-angle of target_orn * orn^-1], until |e_pos| < residualThreshold.
 * End-effector frame: link origin + inertial-frame orientation (SURVEY 8(c) kinematics KAT).
 * Call site: base_robot_arm.py:201-209. */
void or_inverse_kinematics(const OrModel* m, const double* q0, const double target_pos[3], const double target_quat[4], double* q)
{
    int n = m->ndof;
    for (int i = 0; i < n; i++) q[i] = q0[i];
    for (int it = 0; it < 100; it++) {
        double P[OR_MAXL][3], Q[OR_MAXL][4], J[6][OR_MAXD], e[6];
        or_link_states(m, q, P, Q);
        or_jacobian(m, q, m->tcp_link, J);
        for (int c = 0; c < 3; c++) e[c] = target_pos[c] - P[m->tcp_link][c];
        double res = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
        if (it > 0 && res < 1e-8) break;
        double qi[4] = {-Q[m->tcp_link][0], -Q[m->tcp_link][1], -Q[m->tcp_link][2], Q[m->tcp_link][3]}, dq[4];
        quat_mul(dq, target_quat, qi);
        double wv = dq[3] < -1 ? -1 : (dq[3] > 1 ? 1 : dq[3]);
        double ang = 2 * acos(wv);
        double sn = sqrt(dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2]);
        if (ang > M_PI) ang -= 2 * M_PI;
        for (int c = 0; c < 3; c++) e[3 + c] = sn > 1e-300 ? ang * dq[c] / sn : 0.0;
        double A[OR_MAXD][OR_MAXD], b[OR_MAXD], d[OR_MAXD];
        for (int i = 0; i < n; i++) {
            b[i] = 0; for (int r = 0; r < 6; r++) b[i] += J[r][i] * e[r];
            for (int j = 0; j < n; j++) { A[i][j] = i == j ? 0.5 : 0.0; for (int r = 0; r < 6; r++) A[i][j] += J[r][i] * J[r][j]; }
        }
        lu_solve(n, A, b, d);
        double mx = 0; for (int i = 0; i < n; i++) if (fabs(d[i]) > mx) mx = fabs(d[i]);
        double sc = mx > M_PI / 4 ? (M_PI / 4) / mx : 1.0;
        for (int i = 0; i < n; i++) q[i] += sc * d[i];
    }
}

/* blocking_move's constant-velocity retargeting (robot.py:221-246, from google-ravens): step_j = cur + unit(targ - cur) * cv,
 * and cv halves (after it was used) once every joint is closer than cv to its target */
void or_blocking_retarget(int n, const double* q, const double* targ_j, double* cv, double* step_j)
{
    double diff[OR_MAXD], nrm = 0;
    int all_small = 1;
    for (int i = 0; i < n; i++) { diff[i] = targ_j[i] - q[i]; nrm += diff[i] * diff[i]; }
    nrm = sqrt(nrm);
    for (int i = 0; i < n; i++) {
        double vdir = nrm > 0 ? diff[i] / nrm : 0.0;
        step_j[i] = q[i] + vdir * *cv;
        if (!(fabs(diff[i]) < *cv)) all_small = 0;
    }
    if (all_small) *cv /= 2;
}

/* blocking_move's exit test (robot.py:248-259) on the TCP pose and joint speeds read BEFORE the step:
 * sum |pos error| < 2e-4, quaternion angle < 1e-3, sum |qd| < 0.1 */
int or_blocking_reached(const double tpos[3], const double targ_orn[4], const double tcp_pos[3], const double tcp_quat[4], const double* qd, int n)
{
    double tot = 0, pe = 0, ip = 0;
    for (int i = 0; i < n; i++) tot += fabs(qd[i]);
    for (int c = 0; c < 3; c++) pe += fabs(tpos[c] - tcp_pos[c]);
    for (int c = 0; c < 4; c++) ip += targ_orn[c] * tcp_quat[c];
    double ca = 2 * ip * ip - 1; ca = ca < -1 ? -1 : (ca > 1 ? 1 : ca);
    return pe < 2e-4 && acos(ca) < 1e-3 && tot < 0.1;
}

/* Robot.reset (robot.py:114-125): arm.reset (base_robot_arm.py:17-37) -> tcp_direct_workframe_move (:191-226)
 * -> blocking_move(max_steps=1000, constant_vel=0.001) (robot.py:188-260).  Position motors set inside
 * blocking_move pass no `forces`, so pybullet's default (1e5) applies [EXT]. */
int or_robot_reset(const OrModel* m, OrState* s, const double* rest_q, const double tcp_pos_work[3], const double tcp_rpy_work[3])
{
    int n = m->ndof;
    for (int i = 0; i < n; i++) {
        s->q[i] = rest_q[i]; s->qd[i] = 0;
        s->motor_mode[i] = 1; s->target_pos[i] = rest_q[i]; s->target_vel[i] = 0; s->kp[i] = m->pos_gain; s->kd[i] = m->vel_gain; s->max_force[i] = m->max_force;
    }
    /* workframe_to_worldframe (:47-60) */
    double wq[4], tq[4], tpos[3], tquat[4], trpy[3], targ_orn[4], targ_j[OR_MAXD];
    workframe_quat(m, wq);
    or_quat_from_euler(tcp_rpy_work, tq);
    or_mul_transforms(m->workframe_pos, wq, tcp_pos_work, tq, tpos, tquat);
    or_euler_from_quat(tquat, trpy);
    or_quat_from_euler(trpy, targ_orn);
    or_inverse_kinematics(m, s->q, tpos, targ_orn, targ_j);
    for (int i = 0; i < n; i++) { s->target_pos[i] = targ_j[i]; }
    double cv = 0.001;
    int steps = 0;
    for (int it = 0; it < 1000; it++) {
        double P[OR_MAXL][3], Q[OR_MAXL][4], curqd[OR_MAXD], step_j[OR_MAXD];
        or_link_states(m, s->q, P, Q);
        for (int i = 0; i < n; i++) curqd[i] = s->qd[i];
        or_blocking_retarget(n, s->q, targ_j, &cv, step_j);
        for (int i = 0; i < n; i++) {
            s->motor_mode[i] = 1; s->target_pos[i] = step_j[i]; s->target_vel[i] = 0;
            s->kp[i] = m->pos_gain; s->kd[i] = m->vel_gain; s->max_force[i] = 100000.0;
        }
        or_step_sim(m, s);
        steps++;
        if (or_blocking_reached(tpos, targ_orn, P[m->tcp_link], Q[m->tcp_link], curqd, n)) break;
    }
    return steps;
}

/* the IK target of tcp_position_control (base_robot_arm.py:233-252): current work-frame pose + delta, check_TCP_pos_lims,
 * workframe_to_worldframe, getQuaternionFromEuler of the world rpy */
void or_tcp_position_target(const OrModel* m, const double* q, const double delta_work[6], double tpos[3], double targ_orn[4])
{
    double pos[3], rpy[3], tp[3], tr[3];
    or_tcp_pose_workframe(m, q, pos, rpy);
    for (int c = 0; c < 3; c++) {
        tp[c] = pos[c] + delta_work[c]; tr[c] = rpy[c] + delta_work[3 + c];
        tp[c] = tp[c] < m->tcp_lims[c][0] ? m->tcp_lims[c][0] : (tp[c] > m->tcp_lims[c][1] ? m->tcp_lims[c][1] : tp[c]);
        tr[c] = tr[c] < m->tcp_lims[3 + c][0] ? m->tcp_lims[3 + c][0] : (tr[c] > m->tcp_lims[3 + c][1] ? m->tcp_lims[3 + c][1] : tr[c]);
    }
    double wq[4], tq[4], tquat[4], trpy[3];
    workframe_quat(m, wq);
    or_quat_from_euler(tr, tq);
    or_mul_transforms(m->workframe_pos, wq, tp, tq, tpos, tquat);
    or_euler_from_quat(tquat, trpy);
    or_quat_from_euler(trpy, targ_orn);
}

/* Robot.apply_action(control_mode="TCP_position_control") (robot.py:156-186): tcp_position_control (base_robot_arm.py:228-279) then
 * blocking_move(max_steps, constant_vel=None) (robot.py:188-260).  The world stepped inside the blocking move is the env's:
 * arm only (o == NULL), arm + constrained object (object_balance: P == NULL), arm + cube / marble with contacts (P != NULL). */
int or_tcp_position_control_world(const OrModel* m, OrState* s, OrObject* o, OrPush* P, const double delta_work[6], int max_steps)
{
    int n = m->ndof;
    double tpos[3], targ_orn[4], targ_j[OR_MAXD];
    or_tcp_position_target(m, s->q, delta_work, tpos, targ_orn);
    or_inverse_kinematics(m, s->q, tpos, targ_orn, targ_j);
    if (m->mg400_slave) { /* MG400.tcp_position_control (mg400.py:167-172) */
        targ_j[n - 3] = targ_j[1]; targ_j[n - 2] = -targ_j[1]; targ_j[n - 1] = targ_j[1] + targ_j[2];
    }
    for (int i = 0; i < n; i++) {
        s->motor_mode[i] = 1; s->target_pos[i] = targ_j[i]; s->target_vel[i] = 0;
        s->kp[i] = m->pos_gain; s->kd[i] = m->vel_gain; s->max_force[i] = m->max_force;
    }
    int steps = 0;
    for (int it = 0; it < max_steps; it++) {
        double Pl[OR_MAXL][3], Q[OR_MAXL][4], curqd[OR_MAXD];
        or_link_states(m, s->q, Pl, Q);
        for (int i = 0; i < n; i++) curqd[i] = s->qd[i];
        if (o && P) or_step_sim_push(m, s, o, P);
        else if (o) or_step_sim_obj(m, s, o);
        else or_step_sim(m, s);
        steps++;
        if (or_blocking_reached(tpos, targ_orn, Pl[m->tcp_link], Q[m->tcp_link], curqd, n)) break;
    }
    return steps;
}

int or_tcp_position_control(const OrModel* m, OrState* s, const double delta_work[6], int max_steps)
{
    return or_tcp_position_control_world(m, s, NULL, NULL, delta_work, max_steps);
}

/* ------------------------------------------------------------------ tactile raster */
/* TactileSensor.update_cam_frame + get_imgs camera vectors (tactile_sensor.py:150-229) */
void or_camera_frame(const OrModel* m, const double* q, double eye[3], double fwd[3], double up[3], double right[3])
{
    double P[OR_MAXL][3], Q[OR_MAXL][4], cq[4], cp[3], co[4], R[9];
    or_link_states(m, q, P, Q);
    or_quat_from_euler(m->cam_rpy, cq);
    or_mul_transforms(P[m->body_link], Q[m->body_link], m->cam_pos, cq, cp, co);
    or_mat_from_quat(co, R);
    v3 ex = {1, 0, 0}, ez = {0, 0, 1}, f, u, sdir;
    m3mulv(f, R, ex); m3mulv(u, R, ez);
    /* computeViewMatrix(eye, eye + focal*f, u): f = normalize(target-eye), s = normalize(f x up), u = s x f */
    double fn = v3norm(f); v3scale(f, f, 1 / fn);
    double un = v3norm(u); v3scale(u, u, 1 / un);
    v3cross(sdir, f, u); double sn = v3norm(sdir); v3scale(sdir, sdir, 1 / sn);
    v3cross(u, sdir, f);
    v3cpy(eye, cp); v3cpy(fwd, f); v3cpy(up, u); v3cpy(right, sdir);
}

/* z-buffer by per-pixel ray casting (pixel centres; row 0 = top), window depth d = f/(f-n) (1 - n/z) */
static void raster_tris_d(const double eye[3], const double fwd[3], const double up[3], const double right[3],
                          double fov_deg, double near_, double far_, int S, const double* tris, int ntri, float* depth)
{
    double th = tan(fov_deg * (M_PI / 180.0) / 2.0);
    for (int t = 0; t < ntri; t++) {
        const double* T = tris + 9 * t;
        /* eye-space vertices (x right, y up, z forward) */
        double ve[3][3]; int all_front = 1;
        for (int k = 0; k < 3; k++) {
            v3 d; v3sub(d, T + 3 * k, eye);
            ve[k][0] = v3dot(d, right); ve[k][1] = v3dot(d, up); ve[k][2] = v3dot(d, fwd);
            if (ve[k][2] <= 1e-6) all_front = 0;
        }
        int c0 = 0, c1 = S - 1, r0 = 0, r1 = S - 1;
        if (all_front) {
            double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
            for (int k = 0; k < 3; k++) {
                double x = ve[k][0] / (ve[k][2] * th), y = ve[k][1] / (ve[k][2] * th);
                if (x < xmin) xmin = x; if (x > xmax) xmax = x; if (y < ymin) ymin = y; if (y > ymax) ymax = y;
            }
            c0 = (int)floor((xmin + 1) * 0.5 * S - 0.5) - 1; c1 = (int)ceil((xmax + 1) * 0.5 * S - 0.5) + 1;
            r0 = (int)floor((1 - ymax) * 0.5 * S - 0.5) - 1; r1 = (int)ceil((1 - ymin) * 0.5 * S - 0.5) + 1;
            if (c0 < 0) c0 = 0; if (r0 < 0) r0 = 0; if (c1 > S - 1) c1 = S - 1; if (r1 > S - 1) r1 = S - 1;
        }
        v3 e1, e2; v3sub(e1, ve[1], ve[0]); v3sub(e2, ve[2], ve[0]);
        for (int r = r0; r <= r1; r++)
            for (int c = c0; c <= c1; c++) {
                double xn = (c + 0.5) / S * 2 - 1, yn = 1 - (r + 0.5) / S * 2;
                v3 dir = {xn * th, yn * th, 1.0}, pv, tv, qv;
                /* Moller-Trumbore, origin at 0 */
                v3cross(pv, dir, e2);
                double det = v3dot(e1, pv);
                if (fabs(det) < 1e-300) continue;
                double inv = 1.0 / det;
                v3scale(tv, ve[0], -1.0);
                double u = v3dot(tv, pv) * inv;
                if (u < -1e-12 || u > 1 + 1e-12) continue;
                v3cross(qv, tv, e1);
                double v = v3dot(dir, qv) * inv;
                if (v < -1e-12 || u + v > 1 + 1e-12) continue;
                double z = v3dot(e2, qv) * inv; /* dir.z = 1 -> t = z_eye */
                if (z < near_ || z > far_) continue;
                float d = (float)(far_ / (far_ - near_) * (1.0 - near_ / z));
                if (d < depth[r * S + c]) depth[r * S + c] = d;
            }
    }
}

void or_depth_image(const double eye[3], const double fwd[3], const double up[3], const double right[3],
                    double fov_deg, double near_, double far_, int S, const float* tris, int ntri, float* depth_out)
{
    double* td = (double*)malloc(sizeof(double) * 9 * (size_t)ntri);
    for (long i = 0; i < 9L * ntri; i++) td[i] = tris[i];
    for (int i = 0; i < S * S; i++) depth_out[i] = 1.0f;
    raster_tris_d(eye, fwd, up, right, fov_deg, near_, far_, S, td, ntri, depth_out);
    free(td);
}

/* TactileSensor.t_s_camera (tactile_sensor.py:261-294), float32 arithmetic like numpy */
/* TactileSensor.t_s_camera's arithmetic (tactile_sensor.py:268-292) on a given depth image, float32 as numpy does it:
 * diff = cur - nodef; |diff| <= 1e-4 -> 0; uint8(clip(|diff|, 0, 0.05) / 0.05 * 255); border pixels take the baked grey. */
/* `pen_img[full_mask] = 0` (tactile_sensor.py:283-288: pixels whose segmentation id is the sensor body) has no line here on
 * purpose: `cur` is nodef_dep with the stimulus composited in, so a pixel where the body is the nearest surface has
 * cur == nodef_dep -> diff 0 -> already 0, and where the stimulus is nearer the segmentation id is the stimulus' and the
 * reference keeps the value as well.  tests/test_oracle_raster.py::test_body_mask_zeroing_is_an_identity_for_the_composited_depth
 * applies the reference's line with a segmentation built from the sensor meshes and finds no pixel changed. */
void or_postprocess(int S, const float* cur, const float* nodef_dep, const float* nodef_gray, const unsigned char* border_mask,
                    int border_on, unsigned char* img_out)
{
    const float eps = (float)1e-4, maxpen = (float)0.05;
    for (int i = 0; i < S * S; i++) {
        float diff = cur[i] - nodef_dep[i];
        if (diff >= -eps && diff <= eps) diff = 0.0f;
        float pen = fabsf(diff);
        float cl = pen < 0.0f ? 0.0f : (pen > maxpen ? maxpen : pen);
        float val = (cl / maxpen) * 255.0f;
        unsigned char o = (unsigned char)val;
        if (border_on && border_mask[i] == 1) o = (unsigned char)nodef_gray[i];
        img_out[i] = o;
    }
}

void or_tactile_image(const OrModel* m, const double* q, int S, const double* tris_world, int ntri,
                      const float* nodef_dep, const float* nodef_gray, const unsigned char* border_mask,
                      int border_on, unsigned char* img_out, float* depth_out)
{
    double eye[3], fwd[3], up[3], right[3];
    or_camera_frame(m, q, eye, fwd, up, right);
    float* cur = (float*)malloc(sizeof(float) * (size_t)S * S);
    memcpy(cur, nodef_dep, sizeof(float) * (size_t)S * S);
    raster_tris_d(eye, fwd, up, right, m->fov_deg, m->near_, m->far_, S, tris_world, ntri, cur);
    or_postprocess(S, cur, nodef_dep, nodef_gray, border_mask, border_on, img_out);
    if (depth_out) memcpy(depth_out, cur, sizeof(float) * (size_t)S * S);
    free(cur);
}


/* ================================================================ surface_follow: OpenSimplex heightfield
 * base_surface_env.py:443-448 `OpenSimplex(seed=self.np_random.randint(1e8))`, :311-327 noise2 over the 64x64 grid.
 * [EXT] opensimplex package (unpinned): _init / _noise2 restated from the published algorithm. */
void or_opensimplex_init(long long seed_in, short perm[256])
{
    short source[256];
    unsigned long long seed = (unsigned long long)seed_in; /* int64 wrap-around arithmetic */
    for (int i = 0; i < 256; i++) source[i] = (short)i;
    for (int k = 0; k < 3; k++) seed = seed * 6364136223846793005ULL + 1442695040888963407ULL;
    for (int i = 255; i >= 0; i--) {
        seed = seed * 6364136223846793005ULL + 1442695040888963407ULL;
        /* r = int((seed + 31) % (i + 1)) with Python's floor modulo on the signed 64-bit seed */
        const long long sseed = (long long)seed;
        const long long n = i + 1;
        long long r = ((sseed % n) + n) % n;
        r = (r + 31 % n) % n;
        perm[i] = source[r];
        source[r] = source[i];
    }
}

static double os_extrapolate2(const short* perm, long long xsb, long long ysb, double dx, double dy)
{
    static const double G2[16] = {5, 2, 2, 5, -5, 2, -2, 5, 5, -2, 2, -5, -5, -2, -2, -5};
    const int index = perm[(perm[xsb & 0xFF] + ysb) & 0xFF] & 0x0E;
    return G2[index] * dx + G2[index + 1] * dy;
}

double or_opensimplex_noise2(const short perm[256], double x, double y)
{
    const double STRETCH = -0.211324865405187, SQUISH = 0.366025403784439, NORM = 47.0;
    const double stretch_offset = (x + y) * STRETCH;
    const double xs = x + stretch_offset, ys = y + stretch_offset;
    long long xsb = (long long)floor(xs), ysb = (long long)floor(ys);
    const double squish_offset = (double)(xsb + ysb) * SQUISH;
    const double xb = (double)xsb + squish_offset, yb = (double)ysb + squish_offset;
    const double xins = xs - (double)xsb, yins = ys - (double)ysb;
    const double in_sum = xins + yins;
    double dx0 = x - xb, dy0 = y - yb;
    double value = 0.0;
    double dx_ext, dy_ext;
    long long xsv_ext, ysv_ext;

    const double dx1 = dx0 - 1 - SQUISH, dy1 = dy0 - 0 - SQUISH;
    double attn1 = 2 - dx1 * dx1 - dy1 * dy1;
    if (attn1 > 0) { attn1 *= attn1; value += attn1 * attn1 * os_extrapolate2(perm, xsb + 1, ysb + 0, dx1, dy1); }
    const double dx2 = dx0 - 0 - SQUISH, dy2 = dy0 - 1 - SQUISH;
    double attn2 = 2 - dx2 * dx2 - dy2 * dy2;
    if (attn2 > 0) { attn2 *= attn2; value += attn2 * attn2 * os_extrapolate2(perm, xsb + 0, ysb + 1, dx2, dy2); }

    if (in_sum <= 1) {
        const double zins = 1 - in_sum;
        if (zins > xins || zins > yins) {
            if (xins > yins) { xsv_ext = xsb + 1; ysv_ext = ysb - 1; dx_ext = dx0 - 1; dy_ext = dy0 + 1; }
            else { xsv_ext = xsb - 1; ysv_ext = ysb + 1; dx_ext = dx0 + 1; dy_ext = dy0 - 1; }
        } else { xsv_ext = xsb + 1; ysv_ext = ysb + 1; dx_ext = dx0 - 1 - 2 * SQUISH; dy_ext = dy0 - 1 - 2 * SQUISH; }
    } else {
        const double zins = 2 - in_sum;
        if (zins < xins || zins < yins) {
            if (xins > yins) { xsv_ext = xsb + 2; ysv_ext = ysb + 0; dx_ext = dx0 - 2 - 2 * SQUISH; dy_ext = dy0 + 0 - 2 * SQUISH; }
            else { xsv_ext = xsb + 0; ysv_ext = ysb + 2; dx_ext = dx0 + 0 - 2 * SQUISH; dy_ext = dy0 - 2 - 2 * SQUISH; }
        } else { dx_ext = dx0; dy_ext = dy0; xsv_ext = xsb; ysv_ext = ysb; }
        xsb += 1; ysb += 1;
        dx0 = dx0 - 1 - 2 * SQUISH; dy0 = dy0 - 1 - 2 * SQUISH;
    }
    double attn0 = 2 - dx0 * dx0 - dy0 * dy0;
    if (attn0 > 0) { attn0 *= attn0; value += attn0 * attn0 * os_extrapolate2(perm, xsb, ysb, dx0, dy0); }
    double attn_ext = 2 - dx_ext * dx_ext - dy_ext * dy_ext;
    if (attn_ext > 0) { attn_ext *= attn_ext; value += attn_ext * attn_ext * os_extrapolate2(perm, xsv_ext, ysv_ext, dx_ext, dy_ext); }
    return value / NORM;
}

void or_surface_heights(long long seed, int rows, int cols, double interp, double range, double* out)
{
    short perm[256];
    or_opensimplex_init(seed, perm);
    for (int x = 0; x < rows; x++)
        for (int y = 0; y < cols; y++) out[x * cols + y] = or_opensimplex_noise2(perm, x * interp, y * interp) * range;
}

/* ================================================================ object_push: stepSimulation with contact rows
 * See the block comment in tg_oracle.h for what is restated and what is simplified.  Call sites: robots/arms/robot.py:141
 * (stepSimulation), object_push_env.py:204-229 (cube dynamics), sensors/tactile_sensor.py:314-332 (tip dynamics). */
typedef struct {
    int on_arm;          /* 1: tip core <-> cube, 0: cube <-> table */
    double n[3];         /* unit normal: out of the cube face towards the tip / up from the table */
    double pa[3], pb[3]; /* contact point on the arm (tip contacts only); contact point on the cube */
    double dist, mu, cfm, erp;
    int feature;         /* identity across substeps: cube vertex v -> v, hull vertex i -> 8 + i (warm starting) */
} PushContact;

static long long qkey(double x) { return llrint(x * 1e9); } /* comparisons on a 1 nm grid: ties break by index */

/* signed distance of a cube-local point to the cube surface (negative inside) and the face it belongs to */
static double cube_sd(const double half[3], const double l[3], int* axis, int* sign)
{
    double best = -1e300; int a = 0;
    for (int c = 0; c < 3; c++) { const double d = fabs(l[c]) - half[c]; if (d > best) { best = d; a = c; } }
    *axis = a; *sign = l[a] >= 0 ? 1 : -1;
    return best;
}

static void plane_space1(const double n[3], double p[3], double q[3]) /* [EXT] btPlaneSpace1 */
{
    if (fabs(n[2]) > 0.7071067811865475244008443621048490) {
        const double a = n[1] * n[1] + n[2] * n[2], k = 1.0 / sqrt(a);
        p[0] = 0; p[1] = -n[2] * k; p[2] = n[1] * k;
        q[0] = a * k; q[1] = -n[0] * p[2]; q[2] = n[0] * p[1];
    } else {
        const double a = n[0] * n[0] + n[1] * n[1], k = 1.0 / sqrt(a);
        p[0] = -n[1] * k; p[1] = n[0] * k; p[2] = 0;
        q[0] = -n[2] * p[1]; q[1] = n[2] * p[0]; q[2] = a * k;
    }
}

static int push_contacts(const OrModel* m, const Kin* k, const OrObject* o, const OrPush* P, const m3 Rb, PushContact* C)
{
    int nc = 0;
    if (P->shape == 1) {
        /* object_roll: sphere <-> table, sphere <-> cap of the tip core's cylinder */
        const double dtk0 = m->dt * P->tip_k + P->tip_d, dtk = dtk0 < 2.2204460492503131e-16 ? 2.2204460492503131e-16 : dtk0;
        const double dist = o->pos[2] - P->radius - P->table_z;
        if (dist <= P->slop) {
            PushContact* c = &C[nc++];
            c->on_arm = 0; v3set(c->n, 0, 0, 1);
            v3set(c->pb, o->pos[0], o->pos[1], o->pos[2] - P->radius); v3cpy(c->pa, c->pb);
            c->dist = dist; c->mu = P->mu_table; c->cfm = 0.0; c->erp = P->erp; c->feature = 0;
        }
        const int tl = P->tip_link;
        v3 cc, ax, d, t;
        m3mulv(t, k->Rl[tl], P->cyl_pos); v3add(cc, k->pl[tl], t);
        m3mulv(ax, k->Rl[tl], P->cyl_axis);
        v3sub(d, o->pos, cc);
        double h = v3dot(d, ax);
        if (h < 0) { v3scale(ax, ax, -1.0); h = -h; }            /* the cap that faces the sphere */
        const v3 lat = {d[0] - h * ax[0], d[1] - h * ax[1], d[2] - h * ax[2]};
        const double sd = h - P->cyl_half_len - P->radius;
        if (sd <= P->slop && v3dot(lat, lat) <= P->cyl_radius * P->cyl_radius) {
            PushContact* c = &C[nc++];
            c->on_arm = 1;
            v3scale(c->n, ax, -1.0);                               /* out of the sphere, towards the tip */
            for (int q = 0; q < 3; q++) { c->pb[q] = o->pos[q] - P->radius * ax[q]; c->pa[q] = c->pb[q] - sd * ax[q]; }
            c->dist = sd; c->mu = P->mu_tip; c->feature = 8;
            c->cfm = (1.0 / dtk) / m->dt; c->erp = (m->dt * P->tip_k) / dtk;
        }
        return nc;
    }
    /* cube <-> table: cube vertices at or below the table top */
    for (int v = 0; v < 8 && nc < 4; v++) {
        v3 l = {(v & 1) ? P->half[0] : -P->half[0], (v & 2) ? P->half[1] : -P->half[1], (v & 4) ? P->half[2] : -P->half[2]}, w;
        m3mulv(w, Rb, l); v3add(w, w, o->pos);
        const double dist = w[2] - P->table_z;
        if (dist > P->slop) continue;
        PushContact* c = &C[nc++];
        c->on_arm = 0; v3set(c->n, 0, 0, 1); v3cpy(c->pb, w); v3cpy(c->pa, w);
        c->dist = dist; c->mu = P->mu_table; c->cfm = 0.0; c->erp = P->erp; c->feature = v;
    }
    /* tip core hull <-> cube */
    const int tl = P->tip_link;
    m3 M; v3 t, d;
    m3tmul(M, Rb, k->Rl[tl]);
    v3sub(d, k->pl[tl], o->pos); m3tmulv(t, Rb, d);
    long long deep_key = 0, ext_key[3][2]; int deep = -1, ext[3][2], ncand = 0;
    for (int i = 0; i < P->n_hull; i++) {
        v3 l; int ax, sg;
        m3mulv(l, M, P->hull + 3 * i); v3add(l, l, t);
        const double sd = cube_sd(P->half, l, &ax, &sg);
        if (sd > P->slop) continue;
        const long long key = qkey(sd);
        if (ncand == 0 || key < deep_key) { deep_key = key; deep = i; }
        for (int c = 0; c < 3; c++) {
            const long long lk = qkey(l[c]);
            if (ncand == 0 || lk < ext_key[c][0]) { ext_key[c][0] = lk; ext[c][0] = i; }
            if (ncand == 0 || lk > ext_key[c][1]) { ext_key[c][1] = lk; ext[c][1] = i; }
        }
        ncand++;
    }
    if (ncand > 0) {
        v3 l; int A, sg;
        m3mulv(l, M, P->hull + 3 * deep); v3add(l, l, t);
        cube_sd(P->half, l, &A, &sg);
        const int U = (A + 1) % 3, V = (A + 2) % 3;
        const long long dv = qkey(l[V]);
        const long long a0 = llabs(ext_key[V][0] - dv), a1 = llabs(ext_key[V][1] - dv);
        int sel[4] = {deep, ext[U][0], ext[U][1], a0 >= a1 ? ext[V][0] : ext[V][1]};
        const double dtk = m->dt * P->tip_k + P->tip_d;
        for (int j = 0; j < 4; j++) {
            int dup = 0;
            for (int i = 0; i < j; i++) if (sel[i] == sel[j]) dup = 1;
            if (dup) continue;
            int ax; v3 ln = {0, 0, 0}, w;
            m3mulv(l, M, P->hull + 3 * sel[j]); v3add(l, l, t);
            const double sd = cube_sd(P->half, l, &ax, &sg);
            ln[ax] = sg;
            PushContact* c = &C[nc++];
            c->on_arm = 1;
            m3mulv(c->n, Rb, ln);
            m3mulv(w, k->Rl[tl], P->hull + 3 * sel[j]); v3add(c->pa, k->pl[tl], w);
            for (int q = 0; q < 3; q++) c->pb[q] = c->pa[q] - sd * c->n[q];
            c->dist = sd; c->mu = P->mu_tip; c->feature = 8 + sel[j];
            c->cfm = (1.0 / (dtk < 2.2204460492503131e-16 ? 2.2204460492503131e-16 : dtk)) / m->dt;
            c->erp = (m->dt * P->tip_k) / (dtk < 2.2204460492503131e-16 ? 2.2204460492503131e-16 : dtk);
        }
    }
    return nc;
}

void or_step_sim_push(const OrModel* m, OrState* s, OrObject* o, OrPush* P)
{
    const int n = m->ndof;
    double tau_gc[OR_MAXD], tau[OR_MAXD] = {0}, qdd[OR_MAXD];
    or_inverse_dynamics(m, s->q, s->qd, NULL, tau_gc);
    for (int i = 0; i < n; i++) tau[i] = tau_gc[i] - m->joint_damping * s->qd[i];
    Aba A; aba_setup(m, s->q, s->qd, tau, 1, &A);
    aba_accel(m, &A, qdd);
    for (int i = 0; i < n; i++) s->qd[i] += m->dt * qdd[i];

    /* narrow phase on the poses at the start of the step */
    Kin k; kin_compute(m, s->q, &k);
    m3 Rb; or_mat_from_quat(o->quat, Rb);
    PushContact C[OR_MAXC];
    const int nc = push_contacts(m, &k, o, P, Rb, C);

    /* cube: unconstrained velocity update about its COM (gravity, [EXT] multibody base damping, gyroscopic term) */
    double Iinv[3];
    for (int c = 0; c < 3; c++) Iinv[c] = 1.0 / o->inertia[c];
    {
        v3 wl, Iw, gy, al, aw;
        m3tmulv(wl, Rb, o->omg);
        for (int c = 0; c < 3; c++) Iw[c] = o->inertia[c] * wl[c];
        v3cross(gy, wl, Iw);
        const double ka = P->ang_damping * (1.0 + v3norm(o->omg)), kl = P->lin_damping * (1.0 + v3norm(o->vel));
        for (int c = 0; c < 3; c++) al[c] = (-Iw[c] * ka - gy[c]) * Iinv[c];
        m3mulv(aw, Rb, al);
        for (int c = 0; c < 3; c++) {
            const double f = m->gravity[c] * o->mass - o->mass * o->vel[c] * kl;
            o->vel[c] += m->dt * f / o->mass; o->omg[c] += m->dt * aw[c];
        }
    }

    /* rows: motors, then per contact (normal, friction 1, friction 2) */
    enum { MAXR = 3 * OR_MAXC };
    double jr[MAXR][OR_MAXD], ur[MAXR][OR_MAXD], jl[MAXR][3], ja[MAXR][3], ul[MAXR][3], ua[MAXR][3];
    double diag[MAXR], dinv[MAXR], rhs[MAXR], cfmr[MAXR], applied[MAXR];
    double mresp[OR_MAXD][OR_MAXD], mdinv[OR_MAXD], mrhs[OR_MAXD], mlim[OR_MAXD], mapplied[OR_MAXD];
    int mrow[OR_MAXD], nm = 0;
    for (int i = 0; i < n; i++) {
        const double maximp = s->max_force[i] * m->dt;
        if (maximp == 0) continue;
        double f[OR_MAXD] = {0}; f[i] = 1.0;
        aba_delta(m, &A, f, mresp[nm]);
        const double denom = mresp[nm][i];
        mdinv[nm] = denom > 2.2204460492503131e-16 ? 1.0 / denom : 0.0;
        const double v = s->qd[i];
        const double kp = s->motor_mode[i] == 1 ? s->kp[i] : 0.0, tp = s->motor_mode[i] == 1 ? s->target_pos[i] : 0.0;
        const double rhs_v = kp * ((tp - s->q[i]) / m->dt) + v + s->kd[i] * (s->target_vel[i] - v);
        mrhs[nm] = (rhs_v - v) * mdinv[nm];
        mlim[nm] = maximp; mapplied[nm] = 0; mrow[nm] = i;
        nm++;
    }
    for (int c = 0; c < nc; c++) {
        const PushContact* ct = &C[c];
        /* relative velocity of the two contact points (after the unconstrained update) */
        v3 va = {0, 0, 0}, vb, rb, t, vrel;
        v3sub(rb, ct->pb, o->pos);
        v3cross(t, o->omg, rb); v3add(vb, o->vel, t);
        if (ct->on_arm)
            for (int i = P->tip_link; i >= 0; i = m->parent[i]) {
                if (m->jtype[i] != 1) continue;
                v3 r, lin; v3sub(r, ct->pa, k.pl[i]); v3cross(lin, k.aw[i], r);
                v3axpy(va, s->qd[m->dof_of_link[i]], lin);
            }
        if (ct->on_arm) v3sub(vrel, va, vb); else v3cpy(vrel, vb);
        /* friction directions: a fixed orthonormal basis of the contact plane ([EXT] btPlaneSpace1).  Bullet's default
         * aligns the first direction with the lateral relative velocity; with two directions and an isotropic cone that
         * does not change the converged solution, but it makes a solve truncated by the residual exit chaotic at
         * resting contacts (a unit vector built from a ~1e-5 m/s velocity), so it is not restated. */
        v3 dir[3];
        v3cpy(dir[0], ct->n);
        plane_space1(ct->n, dir[1], dir[2]);
        const double sc = ct->on_arm ? -1.0 : 1.0; /* the cube is the second body of a tip contact */
        for (int q = 0; q < 3; q++) {
            const int r = 3 * c + q;
            double f[OR_MAXD] = {0};
            for (int d2 = 0; d2 < n; d2++) { jr[r][d2] = 0; ur[r][d2] = 0; }
            if (ct->on_arm) {
                for (int i = P->tip_link; i >= 0; i = m->parent[i]) {
                    if (m->jtype[i] != 1) continue;
                    v3 rr, lin; v3sub(rr, ct->pa, k.pl[i]); v3cross(lin, k.aw[i], rr);
                    jr[r][m->dof_of_link[i]] = v3dot(dir[q], lin);
                }
                for (int d2 = 0; d2 < n; d2++) f[d2] = jr[r][d2];
                aba_delta(m, &A, f, ur[r]);
            }
            v3scale(jl[r], dir[q], sc);
            v3cross(ja[r], rb, dir[q]); v3scale(ja[r], ja[r], sc);
            v3scale(ul[r], jl[r], 1.0 / o->mass);
            { v3 a, b; m3tmulv(a, Rb, ja[r]); for (int c2 = 0; c2 < 3; c2++) b[c2] = a[c2] * Iinv[c2]; m3mulv(ua[r], Rb, b); }
            double den = v3dot(jl[r], ul[r]) + v3dot(ja[r], ua[r]);
            for (int d2 = 0; d2 < n; d2++) den += jr[r][d2] * ur[r][d2];
            diag[r] = den;
            const double rel = v3dot(dir[q], vrel);
            if (q == 0) {
                dinv[r] = 1.0 / (den + ct->cfm);
                const double positional = ct->dist > 0 ? 0.0 : -ct->dist * ct->erp / m->dt;
                const double velerr = -rel - (ct->dist > 0 ? ct->dist / m->dt : 0.0);
                rhs[r] = (positional + velerr) * dinv[r];
                cfmr[r] = ct->cfm * dinv[r];
            } else {
                dinv[r] = den > 2.2204460492503131e-16 ? 1.0 / den : 0.0;
                rhs[r] = -rel * dinv[r];
                cfmr[r] = 0.0;
            }
            /* [EXT] warm starting: the impulse the same feature ended the previous stepSimulation with, times
             * m_warmstartingFactor (0.85) */
            applied[r] = 0.0;
            for (int w = 0; w < P->ws_n; w++) if (P->ws_feature[w] == ct->feature) applied[r] = P->warmstart * P->ws_impulse[w][q];
        }
    }
    double dv[OR_MAXD] = {0}, dvl[3] = {0, 0, 0}, dva[3] = {0, 0, 0};
    for (int r = 0; r < 3 * nc; r++) {
        if (applied[r] == 0.0) continue;
        for (int d = 0; d < n; d++) dv[d] += ur[r][d] * applied[r];
        for (int q = 0; q < 3; q++) { dvl[q] += ul[r][q] * applied[r]; dva[q] += ua[r][q] * applied[r]; }
    }
    int it = 0;
    for (; it < m->solver_iters; it++) {
        double resid = 0;
        for (int jj = 0; jj < nm; jj++) {
            const int r = (it & 1) ? jj : nm - 1 - jj, d = mrow[r];
            double delta = mrhs[r] - dv[d] * mdinv[r];
            const double sum = mapplied[r] + delta;
            if (sum < -mlim[r]) { delta = -mlim[r] - mapplied[r]; mapplied[r] = -mlim[r]; }
            else if (sum > mlim[r]) { delta = mlim[r] - mapplied[r]; mapplied[r] = mlim[r]; }
            else mapplied[r] = sum;
            for (int i = 0; i < n; i++) dv[i] += mresp[r][i] * delta;
            const double dvel = mdinv[r] != 0 ? delta / mdinv[r] : 0.0;
            if (dvel * dvel > resid) resid = dvel * dvel;
        }
        for (int c = 0; c < nc; c++) { /* normal rows: impulse >= 0 */
            const int r = 3 * c;
            double dot = v3dot(jl[r], dvl) + v3dot(ja[r], dva);
            for (int d = 0; d < n; d++) dot += jr[r][d] * dv[d];
            double delta = rhs[r] - applied[r] * cfmr[r] - dot * dinv[r];
            const double sum = applied[r] + delta;
            if (sum < 0.0) { delta = -applied[r]; applied[r] = 0.0; } else applied[r] = sum;
            for (int d = 0; d < n; d++) dv[d] += ur[r][d] * delta;
            for (int q = 0; q < 3; q++) { dvl[q] += ul[r][q] * delta; dva[q] += ua[r][q] * delta; }
            const double dvel = delta * (diag[r] + C[c].cfm);
            if (dvel * dvel > resid) resid = dvel * dvel;
        }
        for (int c = 0; c < nc; c++) { /* friction pairs inside the cone mu * normal impulse */
            const double lim = C[c].mu * applied[3 * c];
            if (!(applied[3 * c] > 0.0)) continue;
            double sum[2], delta[2];
            for (int q = 0; q < 2; q++) {
                const int r = 3 * c + 1 + q;
                double dot = v3dot(jl[r], dvl) + v3dot(ja[r], dva);
                for (int d = 0; d < n; d++) dot += jr[r][d] * dv[d];
                delta[q] = rhs[r] - dot * dinv[r];
                sum[q] = applied[r] + delta[q];
            }
            const double nrm2 = sum[0] * sum[0] + sum[1] * sum[1];
            if (nrm2 > lim * lim) { const double sc = lim / sqrt(nrm2); sum[0] *= sc; sum[1] *= sc; }
            for (int q = 0; q < 2; q++) {
                const int r = 3 * c + 1 + q;
                delta[q] = sum[q] - applied[r]; applied[r] = sum[q];
                for (int d = 0; d < n; d++) dv[d] += ur[r][d] * delta[q];
                for (int q2 = 0; q2 < 3; q2++) { dvl[q2] += ul[r][q2] * delta[q]; dva[q2] += ua[r][q2] * delta[q]; }
                const double dvel = delta[q] * diag[r];
                if (dvel * dvel > resid) resid = dvel * dvel;
            }
        }
        if (resid <= m->solver_residual_threshold) { it++; break; }
    }
    P->ws_n = nc;
    for (int c = 0; c < nc; c++) { P->ws_feature[c] = C[c].feature; for (int q = 0; q < 3; q++) P->ws_impulse[c][q] = applied[3 * c + q]; }
    P->n_contacts = nc; P->n_iters = it;
    for (int c = 0; c < OR_MAXC; c++) {
        P->normal_impulse[c] = c < nc ? applied[3 * c] : 0.0;
        for (int q = 0; q < 3; q++) P->contact_pos[c][q] = c < nc ? C[c].pb[q] : 0.0;
    }
    for (int i = 0; i < n; i++) { s->qd[i] += dv[i]; s->q[i] += m->dt * s->qd[i]; }
    for (int c = 0; c < 3; c++) { o->vel[c] += dvl[c]; o->omg[c] += dva[c]; o->pos[c] += m->dt * o->vel[c]; }
    {
        const double wn = v3norm(o->omg), ang = wn * m->dt;
        double dq[4] = {0, 0, 0, 1};
        if (wn > 1e-300) { const double sn = sin(0.5 * ang) / wn; dq[0] = o->omg[0] * sn; dq[1] = o->omg[1] * sn; dq[2] = o->omg[2] * sn; dq[3] = cos(0.5 * ang); }
        double qn[4]; quat_mul(qn, dq, o->quat);
        const double nn = sqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
        for (int c = 0; c < 4; c++) o->quat[c] = qn[c] / nn;
    }
}

/* ---- object_roll tactile image: the stimulus is a sphere.  [EXT] pybullet draws a <sphere> visual as a tessellated
 * mesh whose resolution is not recoverable here; the analytic sphere is rendered instead (a 32-segment tessellation
 * differs from it by < 0.5 % of the radius = 0.2 output LSB at the flat TacTip).  Depth per pixel = nearest ray / sphere
 * intersection, then TactileSensor.t_s_camera (tactile_sensor.py:261-294) as in or_tactile_image. */
void or_tactile_image_sphere(const OrModel* m, const double* q, int S, const double centre[3], double radius,
                             const float* nodef_dep, const float* nodef_gray, const unsigned char* border_mask,
                             int border_on, unsigned char* img_out)
{
    double eye[3], fwd[3], up[3], right[3];
    or_camera_frame(m, q, eye, fwd, up, right);
    const double th = tan(m->fov_deg * (M_PI / 180.0) / 2.0);
    v3 d0; v3sub(d0, centre, eye);
    const double s[3] = {v3dot(d0, right), v3dot(d0, up), v3dot(d0, fwd)};
    const double ss = v3dot(s, s) - radius * radius;
    const float eps = (float)1e-4, maxpen = (float)0.05;
    for (int r = 0; r < S; r++)
        for (int c = 0; c < S; c++) {
            const int i = r * S + c;
            float cur = nodef_dep[i];
            const double xn = (c + 0.5) / S * 2 - 1, yn = 1 - (r + 0.5) / S * 2;
            const double d[3] = {xn * th, yn * th, 1.0};
            const double dd = v3dot(d, d), ds = v3dot(d, s), disc = ds * ds - dd * ss;
            if (disc >= 0) {
                const double z = (ds - sqrt(disc)) / dd;
                if (z >= m->near_ && z <= m->far_) {
                    const float dep = (float)(m->far_ / (m->far_ - m->near_) * (1.0 - m->near_ / z));
                    if (dep < cur) cur = dep;
                }
            }
            float diff = cur - nodef_dep[i];
            if (diff >= -eps && diff <= eps) diff = 0.0f;
            const float pen = fabsf(diff);
            const float cl = pen < 0.0f ? 0.0f : (pen > maxpen ? maxpen : pen);
            const float val = (cl / maxpen) * 255.0f;
            unsigned char o = (unsigned char)val;
            if (border_on && border_mask[i] == 1) o = (unsigned char)nodef_gray[i];
            img_out[i] = o;
        }
}
