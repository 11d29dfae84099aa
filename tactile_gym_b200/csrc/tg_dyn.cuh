// tg_dyn.cuh - per-env articulated-body dynamics + constraint solve, one thread per env, fp64.
//
// Replaces, per physics substep, what the reference does with three pybullet calls
// (robots/arms/robot.py:131-141):  calculateInverseDynamics (gravity compensation,
// robots/arms/base_robot_arm.py:174-189) + setJointMotorControlArray(TORQUE_CONTROL) + stepSimulation.
//
// Design (B200-first, not a port of btMultiBody):
//   * fixed joints are merged away at asset-compile time, so a UR5 is 6 bodies, not 11 links;
//   * all spatial quantities are expressed in world axes about the WORLD ORIGIN, which makes composite
//     inertias and wrenches plain sums over the subtree (no 6x6 transforms): joint-space inertia by CRBA,
//     6x6 Cholesky inverse -> the constraint response matrix A = M^-1 that bullet builds column by column
//     with calcAccelerationDeltasMultiDof;
//   * the arm's topology is a template parameter: every loop unrolls, every array lives in registers;
//   * the motor rows (J = e_i) are solved by the same projected Gauss-Seidel sweep bullet runs
//     (numSolverIterations = 150, alternating sweep direction, impulse clamp force*dt, pybullet's
//     solverResidualThreshold exit), staged in registers rather than shared memory because a row is 6 numbers.
// The CPU oracle (oracle/tg_oracle.c) restates bullet's own formulation (ABA in link-COM frames over all
// 11 links); agreement between the two is the parity test.
#pragma once
#include <math.h>

#include "../../include/tactile_gym_b200.h"

#define TGD __device__ __forceinline__

struct TopoChain6 {
    static constexpr int NB = 6;
    __host__ __device__ static constexpr int parent(int i) { return i - 1; }
};
struct TopoMG400 {
    static constexpr int NB = 8;
    __host__ __device__ static constexpr int parent(int i) { return i == 0 ? -1 : (i == 5 ? 0 : i - 1); }
};

// ---------------------------------------------------------------- small vector helpers
TGD void v3cross(double* o, const double* a, const double* b)
{
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
TGD double v3dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
TGD void m3mulv(double* o, const double* M, const double* a)
{
    double x = M[0] * a[0] + M[1] * a[1] + M[2] * a[2];
    double y = M[3] * a[0] + M[4] * a[1] + M[5] * a[2];
    double z = M[6] * a[0] + M[7] * a[1] + M[8] * a[2];
    o[0] = x; o[1] = y; o[2] = z;
}
TGD void m3tmulv(double* o, const double* M, const double* a)
{
    double x = M[0] * a[0] + M[3] * a[1] + M[6] * a[2];
    double y = M[1] * a[0] + M[4] * a[1] + M[7] * a[2];
    double z = M[2] * a[0] + M[5] * a[1] + M[8] * a[2];
    o[0] = x; o[1] = y; o[2] = z;
}
TGD void m3mul(double* o, const double* A, const double* B)
{
    double t[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
#pragma unroll
    for (int i = 0; i < 9; i++) o[i] = t[i];
}

// pybullet frame helpers (same formulas as the oracle; used for the TCP limit check and rewards)
TGD void quat_from_euler(const double* rpy, double* q)
{
    double sr, cr, sp, cp, sy, cy;
    sincos(rpy[0] * 0.5, &sr, &cr); sincos(rpy[1] * 0.5, &sp, &cp); sincos(rpy[2] * 0.5, &sy, &cy);
    q[0] = sr * cp * cy - cr * sp * sy;
    q[1] = cr * sp * cy + sr * cp * sy;
    q[2] = cr * cp * sy - sr * sp * cy;
    q[3] = cr * cp * cy + sr * sp * sy;
}
TGD void euler_from_quat(const double* q, double* rpy)
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
TGD void quat_mul(double* o, const double* a, const double* b)
{
    double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    double y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
    double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
TGD void mat_from_quat(const double* q, double* R)
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
TGD void quat_from_mat(double* q, const double* R)
{
    double tr = R[0] + R[4] + R[8];
    if (tr > 0) {
        double s = sqrt(tr + 1.0);
        q[3] = s * 0.5; s = 0.5 / s;
        q[0] = (R[7] - R[5]) * s; q[1] = (R[2] - R[6]) * s; q[2] = (R[3] - R[1]) * s;
    } else if (R[0] >= R[4] && R[0] >= R[8]) {
        double s = sqrt(R[0] - R[4] - R[8] + 1.0);
        q[0] = s * 0.5; s = 0.5 / s;
        q[3] = (R[7] - R[5]) * s; q[1] = (R[3] + R[1]) * s; q[2] = (R[6] + R[2]) * s;
    } else if (R[4] >= R[8]) {
        double s = sqrt(R[4] - R[8] - R[0] + 1.0);
        q[1] = s * 0.5; s = 0.5 / s;
        q[3] = (R[2] - R[6]) * s; q[2] = (R[7] + R[5]) * s; q[0] = (R[1] + R[3]) * s;
    } else {
        double s = sqrt(R[8] - R[0] - R[4] + 1.0);
        q[2] = s * 0.5; s = 0.5 / s;
        q[3] = (R[3] - R[1]) * s; q[0] = (R[2] + R[6]) * s; q[1] = (R[5] + R[7]) * s;
    }
}

// ---------------------------------------------------------------- kinematics
template <int NB>
struct Kin {
    double R[NB][9]; // world rotation of each body frame
    double p[NB][3]; // world position of each joint origin (= body frame origin)
    double a[NB][3]; // world joint axis
};

// Body frames are world-aligned at q = 0 (scene.py:_align_body_frames), so R_i = R_parent . Rot(axis_i, q_i).
// sc[i] = (sin q_i, cos q_i).
template <class T>
TGD void fk_sc(const TgArm& arm, const double (&sc)[T::NB][2], Kin<T::NB>& k)
{
#pragma unroll
    for (int i = 0; i < T::NB; i++) {
        const int p = T::parent(i);
        const double s = sc[i][0], c = sc[i][1];
        const double ax = arm.axis[i][0], ay = arm.axis[i][1], az = arm.axis[i][2], t1 = 1.0 - c;
        double Rq[9] = {t1 * ax * ax + c,      t1 * ax * ay - s * az, t1 * ax * az + s * ay,
                        t1 * ax * ay + s * az, t1 * ay * ay + c,      t1 * ay * az - s * ax,
                        t1 * ax * az - s * ay, t1 * ay * az + s * ax, t1 * az * az + c};
        if (p < 0) {
#pragma unroll
            for (int c2 = 0; c2 < 3; c2++) k.p[i][c2] = arm.jpos[i][c2];
#pragma unroll
            for (int c2 = 0; c2 < 9; c2++) k.R[i][c2] = Rq[c2];
        } else {
            double t[3];
            m3mulv(t, k.R[p], arm.jpos[i]);
#pragma unroll
            for (int c2 = 0; c2 < 3; c2++) k.p[i][c2] = k.p[p][c2] + t[c2];
            m3mul(k.R[i], k.R[p], Rq);
        }
        m3mulv(k.a[i], k.R[i], arm.axis[i]);
    }
}

template <class T>
TGD void fk(const TgArm& arm, const double* q, Kin<T::NB>& k)
{
    double sc[T::NB][2];
#pragma unroll
    for (int i = 0; i < T::NB; i++) sincos(q[i], &sc[i][0], &sc[i][1]);
    fk_sc<T>(arm, sc, k);
}

// advance (sin q, cos q) by the small angle d = dt * qd: exact trig identity with a 6th-order Taylor of (sin d, cos d)
// (|d| <= 2e-3 -> truncation < 1e-19); larger moves fall back to sincos
TGD void sc_advance(double (&sc)[2], double q_new, double d)
{
    if (fabs(d) > 2e-3) { sincos(q_new, &sc[0], &sc[1]); return; }
    const double d2 = d * d;
    const double sd = d * (1.0 - d2 * (1.0 / 6.0) * (1.0 - d2 * (1.0 / 20.0)));
    const double cd = 1.0 - d2 * 0.5 * (1.0 - d2 * (1.0 / 12.0) * (1.0 - d2 * (1.0 / 30.0)));
    const double s = sc[0], c = sc[1];
    sc[0] = s * cd + c * sd;
    sc[1] = c * cd - s * sd;
}

// world pose of a frame rigidly attached to body b
template <int NB>
TGD void frame_pose(const Kin<NB>& k, int b, const double* lpos, const double* lrot, double* pos, double* R)
{
    double t[3];
    m3mulv(t, k.R[b], lpos);
    pos[0] = k.p[b][0] + t[0]; pos[1] = k.p[b][1] + t[1]; pos[2] = k.p[b][2] + t[2];
    m3mul(R, k.R[b], lrot);
}

// ---------------------------------------------------------------- dynamics
// Spatial inertia about the world origin, world axes: mass m, first moment h = m c, I_O (sym 6: xx xy xz yy yz zz)
struct SpI {
    double m, h[3], I[6];
};
TGD void spi_apply(const SpI& s, const double* w, const double* v, double* n, double* f)
{   // momentum-like map: n = I_O w + h x v ; f = m v + w x h
    double hv[3], wh[3];
    v3cross(hv, s.h, v);
    v3cross(wh, w, s.h);
    n[0] = s.I[0] * w[0] + s.I[1] * w[1] + s.I[2] * w[2] + hv[0];
    n[1] = s.I[1] * w[0] + s.I[3] * w[1] + s.I[4] * w[2] + hv[1];
    n[2] = s.I[2] * w[0] + s.I[4] * w[1] + s.I[5] * w[2] + hv[2];
    f[0] = s.m * v[0] + wh[0]; f[1] = s.m * v[1] + wh[1]; f[2] = s.m * v[2] + wh[2];
}

template <class T>
TGD void body_inertias(const TgArm& arm, const Kin<T::NB>& k, SpI (&sp)[T::NB])
{
#pragma unroll
    for (int i = 0; i < T::NB; i++) {
        double cw[3], t[3];
        m3mulv(t, k.R[i], arm.com[i]);
        cw[0] = k.p[i][0] + t[0]; cw[1] = k.p[i][1] + t[1]; cw[2] = k.p[i][2] + t[2];
        const double m = arm.mass[i];
        // Iw = R Ic R^T
        const double* I = arm.inertia[i];
        double Ic[9] = {I[0], I[1], I[2], I[1], I[3], I[4], I[2], I[4], I[5]}, RI[9];
        m3mul(RI, k.R[i], Ic);
        const double* R = k.R[i];
        double xx = RI[0] * R[0] + RI[1] * R[1] + RI[2] * R[2];
        double xy = RI[0] * R[3] + RI[1] * R[4] + RI[2] * R[5];
        double xz = RI[0] * R[6] + RI[1] * R[7] + RI[2] * R[8];
        double yy = RI[3] * R[3] + RI[4] * R[4] + RI[5] * R[5];
        double yz = RI[3] * R[6] + RI[4] * R[7] + RI[5] * R[8];
        double zz = RI[6] * R[6] + RI[7] * R[7] + RI[8] * R[8];
        const double c2 = v3dot(cw, cw);
        sp[i].m = m;
        sp[i].h[0] = m * cw[0]; sp[i].h[1] = m * cw[1]; sp[i].h[2] = m * cw[2];
        sp[i].I[0] = xx + m * (c2 - cw[0] * cw[0]);
        sp[i].I[1] = xy - m * cw[0] * cw[1];
        sp[i].I[2] = xz - m * cw[0] * cw[2];
        sp[i].I[3] = yy + m * (c2 - cw[1] * cw[1]);
        sp[i].I[4] = yz - m * cw[1] * cw[2];
        sp[i].I[5] = zz + m * (c2 - cw[2] * cw[2]);
    }
}

// Joint-space inertia by the composite-rigid-body algorithm; M is full symmetric NB x NB.
template <class T>
TGD void crba(const Kin<T::NB>& k, SpI (&sp)[T::NB], double (&M)[T::NB][T::NB])
{
    constexpr int NB = T::NB;
    // composite inertias: leaves -> root (plain sums about the common origin)
#pragma unroll
    for (int i = NB - 1; i >= 0; i--) {
        const int p = T::parent(i);
        if (p >= 0) {
            sp[p].m += sp[i].m;
#pragma unroll
            for (int c = 0; c < 3; c++) sp[p].h[c] += sp[i].h[c];
#pragma unroll
            for (int c = 0; c < 6; c++) sp[p].I[c] += sp[i].I[c];
        }
    }
#pragma unroll
    for (int i = 0; i < NB; i++)
#pragma unroll
        for (int j = 0; j < NB; j++) M[i][j] = 0.0;
#pragma unroll
    for (int i = 0; i < NB; i++) {
        double lin[3], n[3], f[3];
        v3cross(lin, k.p[i], k.a[i]); // velocity of the origin-coincident point for unit joint rate
        spi_apply(sp[i], k.a[i], lin, n, f);
#pragma unroll
        for (int j = i; j >= 0; j--) {
            // j runs over ancestors-or-self of i
            bool anc = false;
            {
                int a = i;
#pragma unroll
                for (int s = 0; s < NB; s++) { if (a == j) anc = true; if (a >= 0) a = T::parent(a); }
            }
            if (anc) {
                double lj[3];
                v3cross(lj, k.p[j], k.a[j]);
                double mij = v3dot(k.a[j], n) + v3dot(lj, f);
                M[i][j] = mij; M[j][i] = mij;
            }
        }
    }
}

// Cholesky inverse of an SPD NB x NB matrix (in registers).  One rsqrt per column, no divisions.
template <int NB>
TGD void spd_inverse(const double (&M)[NB][NB], double (&Minv)[NB][NB])
{
    double L[NB][NB], Li[NB][NB], dinv[NB];
#pragma unroll
    for (int j = 0; j < NB; j++) {
        double d = M[j][j];
#pragma unroll
        for (int k2 = 0; k2 < j; k2++) d -= L[j][k2] * L[j][k2];
        const double inv = rsqrt(d);
        dinv[j] = inv;
        L[j][j] = d * inv;
#pragma unroll
        for (int i = j + 1; i < NB; i++) {
            double s = M[i][j];
#pragma unroll
            for (int k2 = 0; k2 < j; k2++) s -= L[i][k2] * L[j][k2];
            L[i][j] = s * inv;
        }
    }
    // Li = L^-1 (lower triangular)
#pragma unroll
    for (int j = 0; j < NB; j++) {
        Li[j][j] = dinv[j];
#pragma unroll
        for (int i = j + 1; i < NB; i++) {
            double s = 0;
#pragma unroll
            for (int k2 = j; k2 < i; k2++) s -= L[i][k2] * Li[k2][j];
            Li[i][j] = s * dinv[i];
        }
    }
#pragma unroll
    for (int i = 0; i < NB; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) {
            double s = 0;
#pragma unroll
            for (int k2 = i; k2 < NB; k2++) s += Li[k2][i] * Li[k2][j];
            Minv[i][j] = s; Minv[j][i] = s;
        }
}

// body velocities about the world origin: w[i], vO[i] (velocity of the body-fixed point at the origin)
template <class T>
TGD void velocities(const Kin<T::NB>& k, const double* qd, double (&w)[T::NB][3], double (&vO)[T::NB][3])
{
#pragma unroll
    for (int i = 0; i < T::NB; i++) {
        const int p = T::parent(i);
        double lin[3];
        v3cross(lin, k.p[i], k.a[i]);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            w[i][c] = (p >= 0 ? w[p][c] : 0.0) + k.a[i][c] * qd[i];
            vO[i][c] = (p >= 0 ? vO[p][c] : 0.0) + lin[c] * qd[i];
        }
    }
}

// Generalised force of bullet's per-link velocity damping (btMultiBody ABA: f = -m v (k + k|v|) at each
// link COM, n = -I w (k + k|w|) in the link's inertial frame), accumulated over each joint's subtree.
template <class T>
TGD void damping_forces(const TgArm& arm, const TgPhysics& ph, const Kin<T::NB>& k, const double (&w)[T::NB][3],
                        const double (&vO)[T::NB][3], double* Q)
{
    constexpr int NB = T::NB;
    double Nw[NB][3], Fw[NB][3];
#pragma unroll
    for (int b = 0; b < NB; b++) {
        // |w| is the same in every frame: one norm per body.  The norms only scale a ~1e-4 N force by (1 + |v|), so a
        // float square root (relative error 6e-8) is exact for every purpose here.
        const double ka = ph.ang_damping * (1.0 + (double)sqrtf((float)v3dot(w[b], w[b])));
        double wb[3], Nb[3] = {0, 0, 0}, Na[3] = {0, 0, 0}, Fa[3] = {0, 0, 0};
        m3tmulv(wb, k.R[b], w[b]); // angular velocity in the body frame
        // mass-carrying URDF links merged into body b (sorted by body at scene-compile time)
#pragma unroll 1
        for (int s = arm.sub_start[b]; s < arm.sub_start[b + 1]; s++) {
            double t[3], x[3], v[3], wl[3], nl[3], nb[3], xf[3];
            m3mulv(t, k.R[b], arm.sub_com[s]);
            x[0] = k.p[b][0] + t[0]; x[1] = k.p[b][1] + t[1]; x[2] = k.p[b][2] + t[2];
            v3cross(v, w[b], x);
            v[0] += vO[b][0]; v[1] += vO[b][1]; v[2] += vO[b][2];
            const double kl = ph.lin_damping * (1.0 + (double)sqrtf((float)v3dot(v, v)));
            m3tmulv(wl, arm.sub_rot[s], wb);            // ... in the link's inertial frame
#pragma unroll
            for (int c = 0; c < 3; c++) nl[c] = -arm.sub_inertia[s][c] * wl[c] * ka;
            m3mulv(nb, arm.sub_rot[s], nl);             // torque back in the body frame
            const double ms = arm.sub_mass[s];
            double f[3] = {-ms * v[0] * kl, -ms * v[1] * kl, -ms * v[2] * kl};
            v3cross(xf, x, f);
#pragma unroll
            for (int c = 0; c < 3; c++) { Nb[c] += nb[c]; Na[c] += xf[c]; Fa[c] += f[c]; }
        }
        double nwv[3];
        m3mulv(nwv, k.R[b], Nb);
#pragma unroll
        for (int c = 0; c < 3; c++) { Nw[b][c] = nwv[c] + Na[c]; Fw[b][c] = Fa[c]; }
    }
#pragma unroll
    for (int i = NB - 1; i >= 0; i--) {
        double lin[3];
        v3cross(lin, k.p[i], k.a[i]);
        Q[i] = v3dot(k.a[i], Nw[i]) + v3dot(lin, Fw[i]);
        const int p = T::parent(i);
        if (p >= 0) {
#pragma unroll
            for (int c = 0; c < 3; c++) { Nw[p][c] += Nw[i][c]; Fw[p][c] += Fw[i][c]; }
        }
    }
}

// Recursive Newton-Euler, tau = M qdd + C(q,qd) + G   (pb.calculateInverseDynamics, base_robot_arm.py:174-179)
// `sp` must hold the ISOLATED body inertias (before crba() makes them composite).
template <class T>
TGD void rnea(const TgPhysics& ph, const Kin<T::NB>& k, const SpI (&sp)[T::NB], const double* qd, const double* qdd, double* tau)
{
    constexpr int NB = T::NB;
    double w[NB][3], v[NB][3], aw[NB][3], av[NB][3], N[NB][3], F[NB][3];
#pragma unroll
    for (int i = 0; i < NB; i++) {
        const int p = T::parent(i);
        double lin[3], jw[3], jv[3], c1[3], c2[3], c3[3];
        v3cross(lin, k.p[i], k.a[i]);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            jw[c] = k.a[i][c] * qd[i]; jv[c] = lin[c] * qd[i];
            w[i][c] = (p >= 0 ? w[p][c] : 0.0) + jw[c];
            v[i][c] = (p >= 0 ? v[p][c] : 0.0) + jv[c];
        }
        // spatial motion cross  v_i x (S qd) = [w x jw ; w x jv + v x jw]
        v3cross(c1, w[i], jw); v3cross(c2, w[i], jv); v3cross(c3, v[i], jw);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            aw[i][c] = (p >= 0 ? aw[p][c] : 0.0) + k.a[i][c] * (qdd ? qdd[i] : 0.0) + c1[c];
            av[i][c] = (p >= 0 ? av[p][c] : -ph.gravity[c]) + lin[c] * (qdd ? qdd[i] : 0.0) + c2[c] + c3[c];
        }
        // f = I a + v x* (I v)
        double n1[3], f1[3], n2[3], f2[3], t1[3], t2[3], t3[3];
        spi_apply(sp[i], aw[i], av[i], n1, f1);
        spi_apply(sp[i], w[i], v[i], n2, f2);
        v3cross(t1, w[i], n2); v3cross(t2, v[i], f2); v3cross(t3, w[i], f2);
#pragma unroll
        for (int c = 0; c < 3; c++) { N[i][c] = n1[c] + t1[c] + t2[c]; F[i][c] = f1[c] + t3[c]; }
    }
#pragma unroll
    for (int i = NB - 1; i >= 0; i--) {
        double lin[3];
        v3cross(lin, k.p[i], k.a[i]);
        tau[i] = v3dot(k.a[i], N[i]) + v3dot(lin, F[i]);
        const int p = T::parent(i);
        if (p >= 0) {
#pragma unroll
            for (int c = 0; c < 3; c++) { N[p][c] += N[i][c]; F[p][c] += F[i][c]; }
        }
    }
}

// Joint motors of one env: all control joints share mode / gains / force limit (that is how the
// reference drives them: base_robot_arm.py:27-37, 325-332, robot.py:228-236)
template <int NB>
struct Motors {
    int mode;            // 0 velocity, 1 position
    double kp, kd, max_force;
    double target_pos[NB], target_vel[NB];
};

// First half of one Robot.step_sim(): gravity compensation + unconstrained velocity update; leaves A = M^-1.
template <class T>
TGD void robot_pre(const TgArm& arm, const TgPhysics& ph, const double* q, double* qd, const double (&sc)[T::NB][2], double (&A)[T::NB][T::NB])
{
    constexpr int NB = T::NB;
    double qdd[NB];
    {
        Kin<NB> k;
        fk_sc<T>(arm, sc, k);
        SpI sp[NB];
        body_inertias<T>(arm, k, sp);
        double tau[NB];
        {
            double w[NB][3], vO[NB][3];
            velocities<T>(k, qd, w, vO);
            damping_forces<T>(arm, ph, k, w, vO, tau);
        }
        if (!ph.gravity_comp) {
            // no compensation torque: the bias forces act.  (With compensation on - the only mode the
            // reference uses, robot.py:138 - tau_gc = RNEA(q, qd, 0) cancels the bias identically, so
            // neither is evaluated.)
            double bias[NB];
            rnea<T>(ph, k, sp, qd, nullptr, bias);
#pragma unroll
            for (int i = 0; i < NB; i++) tau[i] -= bias[i];
        }
#pragma unroll
        for (int i = 0; i < NB; i++) tau[i] -= ph.joint_damping * qd[i];
        double M[NB][NB];
        crba<T>(k, sp, M);
        spd_inverse<NB>(M, A);
#pragma unroll
        for (int i = 0; i < NB; i++) {
            double s = 0;
#pragma unroll
            for (int j = 0; j < NB; j++) s += A[i][j] * tau[j];
            qdd[i] = s;
        }
    }
#pragma unroll
    for (int i = 0; i < NB; i++) qd[i] += ph.dt * qdd[i];
}

// One Robot.step_sim(): gravity compensation + stepSimulation with motor rows only.
// Returns the number of PGS sweeps executed (diagnostics).
template <class T>
TGD int substep(const TgArm& arm, const TgPhysics& ph, double* q, double* qd, double (&sc)[T::NB][2], const Motors<T::NB>& mot)
{
    constexpr int NB = T::NB;
    double A[NB][NB]; // M^-1
    robot_pre<T>(arm, ph, q, qd, sc, A);

    // motor rows: J = e_i, response column A[:, i], |impulse| <= force * dt
    double rhs[NB], dinv[NB], applied[NB], dv[NB];
    const double lim = mot.max_force * ph.dt;
#pragma unroll
    for (int i = 0; i < NB; i++) {
        const double denom = A[i][i];
        dinv[i] = denom > 2.2204460492503131e-16 ? 1.0 / denom : 0.0;
        const double v = qd[i];
        const double pos_stab = mot.mode == 1 ? mot.kp * ((mot.target_pos[i] - q[i]) / ph.dt) : 0.0;
        const double rhs_v = pos_stab + v + mot.kd * (mot.target_vel[i] - v);
        rhs[i] = (rhs_v - v) * dinv[i];
        applied[i] = 0; dv[i] = 0;
    }
    // Projected Gauss-Seidel: sweeps alternate direction (even: back to front); stop when the largest
    // squared velocity-level change of a sweep is <= solverResidualThreshold (pybullet default 1e-7) [EXT].
    // The clamp is written with selects: same values as bullet's if/else chain, no divergence.
    auto row = [&](int r, double& resid) {
        double delta = rhs[r] - dv[r] * dinv[r];
        const double sum = applied[r] + delta;
        const bool lo = sum < -lim, hi = sum > lim;
        delta = lo ? (-lim - applied[r]) : (hi ? (lim - applied[r]) : delta);
        applied[r] = lo ? -lim : (hi ? lim : sum);
#pragma unroll
        for (int i = 0; i < NB; i++) dv[i] += A[r][i] * delta;
        const double dvel = delta * A[r][r];
        resid = fmax(resid, dvel * dvel);
    };
    int it = 0;
    if (lim != 0.0) {
#pragma unroll 1
        for (; it < ph.solver_iters; it++) {
            double resid = 0;
            if (it & 1) {
#pragma unroll
                for (int r = 0; r < NB; r++) row(r, resid);
            } else {
#pragma unroll
                for (int r = NB - 1; r >= 0; r--) row(r, resid);
            }
            if (resid <= ph.solver_residual_threshold) { it++; break; }
        }
    }
#pragma unroll
    for (int i = 0; i < NB; i++) {
        qd[i] += dv[i];
        const double d = ph.dt * qd[i];
        q[i] += d;
        sc_advance(sc[i], q[i], d);
    }
    return it;
}

template <class T> TGD void tcp_world(const TgArm& arm, const Kin<T::NB>& k, double* pos, double* quat);
template <class T> TGD void tcp_jacobian(const TgArm& arm, const Kin<T::NB>& k, const double* tcp_pos, double (&J)[6][T::NB]);

// Free rigid body hanging on the TCP by a point-to-point constraint (object_balance).  pos/quat: base-link COM
// pose, vel: world velocity of that point, omg: world angular velocity.
struct ObjState {
    double pos[3], quat[4], vel[3], omg[3];
    double ext_pos[3]; // world point where the one-step external force applies
    int ext_pending;
    double grav_z, pivot_z; // per-episode gravity; constraint pivot z in the base-link COM frame
    double mass;            // object_push: the cube's mass this episode (rand_obj_mass)
};

// Robot.step_sim() with the object in the world: 6 motor rows + 3 point-to-point rows
// ([EXT] btMultiBodyPoint2Point: rows along -x,-y,-z on the arm / +x,+y,+z on the object, Baumgarte erp * gap / dt).
template <class T>
TGD int substep_obj(const TgArm& arm, const TgPhysics& ph, const TgTask& task, double* q, double* qd, double (&sc)[T::NB][2],
                    const Motors<T::NB>& mot, ObjState& o)
{
    constexpr int NB = T::NB;
    constexpr int NR = NB + 3;
    double A[NB][NB];
    robot_pre<T>(arm, ph, q, qd, sc, A);

    // object: unconstrained update about the composite COM (gravity, one-step external force, gyroscopic torque)
    double Rb[9], dw[3], cw[3], vc[3], Iinv[3];
    mat_from_quat(o.quat, Rb);
    m3mulv(dw, Rb, task.obj_com_off);
#pragma unroll
    for (int c = 0; c < 3; c++) { cw[c] = o.pos[c] + dw[c]; Iinv[c] = 1.0 / task.obj_inertia[c]; }
    {
        double t[3];
        v3cross(t, o.omg, dw);
#pragma unroll
        for (int c = 0; c < 3; c++) vc[c] = o.vel[c] + t[c];
        double F[3] = {0.0, 0.0, o.grav_z * task.obj_mass}, Tq[3] = {0, 0, 0};
        if (o.ext_pending) {
            const double ef[3] = {0.0, 0.0, -task.obj_force};
            double r[3] = {o.ext_pos[0] - cw[0], o.ext_pos[1] - cw[1], o.ext_pos[2] - cw[2]};
            v3cross(Tq, r, ef);
            F[2] += ef[2];
            o.ext_pending = 0;
        }
        double wl[3], Iwv[3], gy[3], Tl[3], al[3], aw[3];
        m3tmulv(wl, Rb, o.omg);
#pragma unroll
        for (int c = 0; c < 3; c++) Iwv[c] = task.obj_inertia[c] * wl[c];
        v3cross(gy, wl, Iwv);
        m3tmulv(Tl, Rb, Tq);
#pragma unroll
        for (int c = 0; c < 3; c++) al[c] = (Tl[c] - gy[c]) * Iinv[c];
        m3mulv(aw, Rb, al);
#pragma unroll
        for (int c = 0; c < 3; c++) { vc[c] += ph.dt * F[c] / task.obj_mass; o.omg[c] += ph.dt * aw[c]; }
    }

    // rows 0..NB-1: motors (J = e_i); rows NB..NB+2: point-to-point
    double ur[3][NB], jr[3][NB], jba[3][3], uba[3][3]; // p2p rows: arm jacobian / response, object angular jacobian / response
    double rhs[NR], dinv[NR], diag[NR], applied[NR];
    const double lim_m = mot.max_force * ph.dt;
#pragma unroll
    for (int i = 0; i < NB; i++) {
        diag[i] = A[i][i];
        dinv[i] = diag[i] > 2.2204460492503131e-16 ? 1.0 / diag[i] : 0.0;
        const double v = qd[i];
        const double pos_stab = mot.mode == 1 ? mot.kp * ((mot.target_pos[i] - q[i]) / ph.dt) : 0.0;
        const double rhs_v = pos_stab + v + mot.kd * (mot.target_vel[i] - v);
        rhs[i] = (rhs_v - v) * dinv[i];
        applied[i] = 0;
    }
    {
        Kin<NB> k;
        fk_sc<T>(arm, sc, k);
        double pa[3], tq[4], J[6][NB], pb[3], rb[3], t[3];
        tcp_world<T>(arm, k, pa, tq);
        tcp_jacobian<T>(arm, k, pa, J);
        const double pl[3] = {0.0, 0.0, o.pivot_z};
        m3mulv(t, Rb, pl);
#pragma unroll
        for (int c = 0; c < 3; c++) { pb[c] = o.pos[c] + t[c]; rb[c] = pb[c] - cw[c]; }
#pragma unroll
        for (int i = 0; i < 3; i++) {
            double nB[3] = {0, 0, 0};
            nB[i] = 1.0;
            double denom = 0, rel = 0;
#pragma unroll
            for (int d = 0; d < NB; d++) jr[i][d] = -J[i][d];
#pragma unroll
            for (int d = 0; d < NB; d++) {
                double u = 0;
#pragma unroll
                for (int e = 0; e < NB; e++) u += A[d][e] * jr[i][e];
                ur[i][d] = u;
                denom += jr[i][d] * u;
                rel += jr[i][d] * qd[d];
            }
            double jl[3], ul[3];
            v3cross(jba[i], rb, nB);
            m3tmulv(jl, Rb, jba[i]);
#pragma unroll
            for (int c = 0; c < 3; c++) ul[c] = jl[c] * Iinv[c];
            m3mulv(uba[i], Rb, ul);
            denom += 1.0 / task.obj_mass + v3dot(jba[i], uba[i]);
            rel += vc[i] + v3dot(jba[i], o.omg);
            diag[NB + i] = denom;
            dinv[NB + i] = denom > 2.2204460492503131e-16 ? 1.0 / denom : 0.0;
            const double pos_error = -(pa[i] - pb[i]);               // (pivotA - pivotB) . (-e_i)
            const double positional = -pos_error * task.p2p_erp / ph.dt;
            rhs[NB + i] = (positional - rel) * dinv[NB + i];
            applied[NB + i] = 0;
        }
    }
    double dv[NB], dvl[3] = {0, 0, 0}, dva[3] = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < NB; i++) dv[i] = 0;
    auto row_m = [&](int r, double& resid) {
        double delta = rhs[r] - dv[r] * dinv[r];
        const double sum = applied[r] + delta;
        const bool lo = sum < -lim_m, hi = sum > lim_m;
        delta = lo ? (-lim_m - applied[r]) : (hi ? (lim_m - applied[r]) : delta);
        applied[r] = lo ? -lim_m : (hi ? lim_m : sum);
#pragma unroll
        for (int i = 0; i < NB; i++) dv[i] += A[r][i] * delta;
        const double dvel = delta * diag[r];
        resid = fmax(resid, dvel * dvel);
    };
    auto row_p = [&](int i, double& resid) {
        const int r = NB + i;
        const double lim = task.p2p_max_impulse;
        double dot = dvl[i] + v3dot(jba[i], dva);
#pragma unroll
        for (int d = 0; d < NB; d++) dot += jr[i][d] * dv[d];
        double delta = rhs[r] - dot * dinv[r];
        const double sum = applied[r] + delta;
        const bool lo = sum < -lim, hi = sum > lim;
        delta = lo ? (-lim - applied[r]) : (hi ? (lim - applied[r]) : delta);
        applied[r] = lo ? -lim : (hi ? lim : sum);
#pragma unroll
        for (int d = 0; d < NB; d++) dv[d] += ur[i][d] * delta;
        dvl[i] += delta / task.obj_mass;
#pragma unroll
        for (int c = 0; c < 3; c++) dva[c] += uba[i][c] * delta;
        const double dvel = delta * diag[r];
        resid = fmax(resid, dvel * dvel);
    };
    int it = 0;
#pragma unroll 1
    for (; it < ph.solver_iters; it++) {
        double resid = 0;
        if (it & 1) {
            if (lim_m != 0.0) {
#pragma unroll
                for (int r = 0; r < NB; r++) row_m(r, resid);
            }
#pragma unroll
            for (int i = 0; i < 3; i++) row_p(i, resid);
        } else {
#pragma unroll
            for (int i = 2; i >= 0; i--) row_p(i, resid);
            if (lim_m != 0.0) {
#pragma unroll
                for (int r = NB - 1; r >= 0; r--) row_m(r, resid);
            }
        }
        if (resid <= ph.solver_residual_threshold) { it++; break; }
    }
#pragma unroll
    for (int i = 0; i < NB; i++) {
        qd[i] += dv[i];
        const double d = ph.dt * qd[i];
        q[i] += d;
        sc_advance(sc[i], q[i], d);
    }
    // object: apply the impulses, integrate (exponential map), back to the base-link COM
    double cnew[3];
#pragma unroll
    for (int c = 0; c < 3; c++) { vc[c] += dvl[c]; o.omg[c] += dva[c]; cnew[c] = cw[c] + ph.dt * vc[c]; }
    {
        const double wn = sqrt(v3dot(o.omg, o.omg)), ang = wn * ph.dt;
        double dq[4] = {0, 0, 0, 1};
        if (wn > 1e-300) {
            double sn, cs;
            sincos(0.5 * ang, &sn, &cs);
            sn /= wn;
            dq[0] = o.omg[0] * sn; dq[1] = o.omg[1] * sn; dq[2] = o.omg[2] * sn; dq[3] = cs;
        }
        double qn[4];
        quat_mul(qn, dq, o.quat);
        const double nn = 1.0 / sqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
#pragma unroll
        for (int c = 0; c < 4; c++) o.quat[c] = qn[c] * nn;
    }
    mat_from_quat(o.quat, Rb);
    m3mulv(dw, Rb, task.obj_com_off);
    {
        double t[3];
        v3cross(t, o.omg, dw);
#pragma unroll
        for (int c = 0; c < 3; c++) { o.pos[c] = cnew[c] - dw[c]; o.vel[c] = vc[c] - t[c]; }
    }
    return it;
}

// ---------------------------------------------------------------- TCP helpers
// getLinkState(tcp)[0:2] and the workframe pose built from it (base_robot_arm.py:136-172)
template <class T>
TGD void tcp_world(const TgArm& arm, const Kin<T::NB>& k, double* pos, double* quat)
{
    double R[9];
#pragma unroll
    for (int b = 0; b < T::NB; b++)
        if (arm.tcp_body == b) { frame_pose<T::NB>(k, b, arm.tcp_pos, arm.tcp_rot, pos, R); }
    quat_from_mat(quat, R);
}

TGD void world_to_work_at(const TgTask& task, const double* wf_pos, const double* pos, const double* quat, double* wpos, double* wrpy);
TGD void world_to_work(const TgTask& task, const double* pos, const double* quat, double* wpos, double* wrpy)
{
    world_to_work_at(task, task.workframe_pos, pos, quat, wpos, wrpy);
}
// same with an explicit workframe origin (object_roll moves it every episode, object_roll_env.py:197-202)
TGD void world_to_work_at(const TgTask& task, const double* wf_pos, const double* pos, const double* quat, double* wpos, double* wrpy)
{
    // worldframe_to_workframe (base_robot_arm.py:62-74): rpy -> quat -> inverse workframe -> rpy
    double rpy_w[3], qw[4], wq[4], wqi[4], R[9], d[3], oq[4];
    euler_from_quat(quat, rpy_w);
    quat_from_euler(rpy_w, qw);
    quat_from_euler(task.workframe_rpy, wq);
    wqi[0] = -wq[0]; wqi[1] = -wq[1]; wqi[2] = -wq[2]; wqi[3] = wq[3];
    mat_from_quat(wqi, R);
    // inverse transform: p' = R^-1 (p - t) computed as bullet does: inv_pos = -(R^-1 t); out = inv_pos + R^-1 p
    double it[3], ip[3];
    m3mulv(it, R, wf_pos);
    m3mulv(ip, R, pos);
    d[0] = -it[0] + ip[0]; d[1] = -it[1] + ip[1]; d[2] = -it[2] + ip[2];
    wpos[0] = d[0]; wpos[1] = d[1]; wpos[2] = d[2];
    quat_mul(oq, wqi, qw);
    euler_from_quat(oq, wrpy);
}

// solve the N x N system A x = b by LU with partial pivoting, all in registers (Mx is the augmented matrix);
// returns min|pivot| / max|pivot|
template <int N>
TGD double solveN(double (&Mx)[N][N + 1], double* x)
{
    double pmin = 1e300, pmax = 0;
#pragma unroll
    for (int c = 0; c < N; c++) {
        // pivot search + swap by value
#pragma unroll
        for (int r = c + 1; r < N; r++) {
            if (fabs(Mx[r][c]) > fabs(Mx[c][c])) {
#pragma unroll
                for (int j = 0; j < N + 1; j++) { double t = Mx[c][j]; Mx[c][j] = Mx[r][j]; Mx[r][j] = t; }
            }
        }
        const double pv = fabs(Mx[c][c]);
        pmin = fmin(pmin, pv); pmax = fmax(pmax, pv);
        const double inv = pv > 0 ? 1.0 / Mx[c][c] : 0.0;
#pragma unroll
        for (int r = c + 1; r < N; r++) {
            const double f = Mx[r][c] * inv;
#pragma unroll
            for (int j = c; j < N + 1; j++) Mx[r][j] -= f * Mx[c][j];
        }
    }
#pragma unroll
    for (int r = N - 1; r >= 0; r--) {
        double s = Mx[r][N];
#pragma unroll
        for (int j = r + 1; j < N; j++) s -= Mx[r][j] * x[j];
        x[r] = Mx[r][r] != 0.0 ? s / Mx[r][r] : 0.0;
    }
    return pmax > 0 ? pmin / pmax : 0.0;
}
TGD double solve6(double (&Mx)[6][7], double* x) { return solveN<6>(Mx, x); }

// x = pinv(J) v for a 6 x NB Jacobian: one-sided Jacobi SVD of J^T, singular values below 1e-15 * max dropped
// (np.linalg.pinv's default rcond; base_robot_arm.py:316-319, mg400.py:109)
template <int NB>
TGD void pinv_apply(const double (&J)[6][NB], const double* v, double* x)
{
    double W[NB][6], V[6][6];
#pragma unroll
    for (int i = 0; i < NB; i++)
#pragma unroll
        for (int j = 0; j < 6; j++) W[i][j] = J[j][i];
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j < 6; j++) V[i][j] = i == j ? 1.0 : 0.0;
#pragma unroll 1
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0;
#pragma unroll
        for (int p = 0; p < 5; p++)
#pragma unroll
            for (int q = p + 1; q < 6; q++) {
                double a = 0, b = 0, c = 0;
#pragma unroll
                for (int i = 0; i < NB; i++) { a += W[i][p] * W[i][p]; b += W[i][q] * W[i][q]; c += W[i][p] * W[i][q]; }
                if (!(fabs(c) <= 1e-300 || fabs(c) <= 1e-17 * sqrt(a * b))) {
                    off += fabs(c);
                    const double zeta = (b - a) / (2 * c);
                    const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1 + zeta * zeta));
                    const double cs = 1 / sqrt(1 + t * t), sn = cs * t;
#pragma unroll
                    for (int i = 0; i < NB; i++) { const double wp = W[i][p], wq = W[i][q]; W[i][p] = cs * wp - sn * wq; W[i][q] = sn * wp + cs * wq; }
#pragma unroll
                    for (int i = 0; i < 6; i++) { const double vp = V[i][p], vq = V[i][q]; V[i][p] = cs * vp - sn * vq; V[i][q] = sn * vp + cs * vq; }
                }
            }
        if (off == 0) break;
    }
    double sig[6], smax = 0;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        double a = 0;
#pragma unroll
        for (int i = 0; i < NB; i++) a += W[i][k] * W[i][k];
        sig[k] = sqrt(a); smax = fmax(smax, sig[k]);
    }
#pragma unroll
    for (int i = 0; i < NB; i++) x[i] = 0;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        if (sig[k] > 1e-15 * smax) {
            double vk = 0;
#pragma unroll
            for (int j = 0; j < 6; j++) vk += V[j][k] * v[j];
#pragma unroll
            for (int i = 0; i < NB; i++) x[i] += (W[i][k] / sig[k]) * vk / sig[k];
        }
    }
}

// geometric Jacobian of the TCP point, world frame: rows 0-2 linear, 3-5 angular (pb.calculateJacobian)
template <class T>
TGD void tcp_jacobian(const TgArm& arm, const Kin<T::NB>& k, const double* tcp_pos, double (&J)[6][T::NB])
{
#pragma unroll
    for (int j = 0; j < T::NB; j++) {
        bool anc = false;
        {
            int a = arm.tcp_body;
#pragma unroll
            for (int s = 0; s < T::NB; s++) { if (a == j) anc = true; if (a >= 0) { int pa = -1;
#pragma unroll
                for (int b = 0; b < T::NB; b++) if (a == b) pa = T::parent(b);
                a = pa; } }
        }
        double r[3] = {tcp_pos[0] - k.p[j][0], tcp_pos[1] - k.p[j][1], tcp_pos[2] - k.p[j][2]}, lin[3];
        v3cross(lin, k.a[j], r);
#pragma unroll
        for (int c = 0; c < 3; c++) { J[c][j] = anc ? lin[c] : 0.0; J[3 + c][j] = anc ? k.a[j][c] : 0.0; }
    }
}
