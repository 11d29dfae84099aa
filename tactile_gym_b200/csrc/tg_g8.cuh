// tg_g8.cuh - the physics substep with one env per GROUP OF 8 LANES (four envs per warp), one body / joint per lane.
//
// north_star's mapping ("one warp-group per env instance, ... warp-shuffle reductions") for the motor-row tasks.  The
// one-thread-per-env substep (tg_dyn.cuh) keeps every per-joint quantity in the registers of a single thread: ~5.2 k
// instructions per substep in one dependent fp64 chain, 255 registers, ~80 KB of unrolled code - at N = 4096 the SMs hold one
// warp per scheduler with 8 active lanes and wait on instruction fetch and fp64 latency (profiles/r01_step_kernel_final.md).
// Here lane j of a group owns body j:
//   * forward kinematics and body velocities are PREFIX operations over the kinematic tree: pointer jumping with shuffles,
//     3 rounds for any tree of depth <= 8;
//   * the per-body work (spatial inertia, bullet's per-link damping wrench) is lane-local;
//   * composite inertias and subtree wrenches are SUFFIX sums over the tree (shuffle-down rounds, specialised per topology);
//   * CRBA: lane j walks up its ancestors and holds column j of M; the matrix is then all-gathered, every lane runs the same
//     small Cholesky factorisation and solves for its OWN column of A = M^-1 (the constraint response of motor row j);
//   * projected Gauss-Seidel over the motor rows: the lane that owns row r computes its impulse from lane-local data, one
//     broadcast, one FMA per lane (dv_i += A[r][i] delta) - same row order, same clamp, same residual exit as tg_dyn.cuh.
// ~1.4 k instructions per substep for four envs, a loop body that fits the instruction cache, and two warps per scheduler at
// N = 4096 to hide the fp64 latency.  Arithmetic per entry is the same as in tg_dyn.cuh; only the order of a few sums differs
// (1e-16), the oracle tolerance (1e-10) is unchanged.
#pragma once
#include "tg_env.cuh"

#define G8 8
#define G8_FULL 0xffffffffu

TGD double g8_get(double v, int src) { return __shfl_sync(G8_FULL, v, src, G8); }
TGD int g8_geti(int v, int src) { return __shfl_sync(G8_FULL, v, src, G8); }
TGD double g8_down(double v, int d) { return __shfl_down_sync(G8_FULL, v, d, G8); }

// ---------------------------------------------------------------- topology helpers (lane = body index)
template <class T> struct G8Topo;
template <> struct G8Topo<TopoChain6> {
    static constexpr int MAXDEPTH = 5;
    __host__ __device__ static constexpr int depth(int i) { return i; }
    __host__ __device__ static constexpr bool is_anc(int k, int j) { return k <= j; } // k ancestor-or-self of j
    // v <- sum of v over the subtree of each lane: suffix sums along the chain.  Lanes 6, 7 are leaves below lane 5 when they
    // lend a hand with its sub-links (their own "subtree" sums are never read), zero otherwise.
    template <int CNT> TGD static void subtree_sum(double (&v)[CNT], int j)
    {
#pragma unroll
        for (int d = 1; d < G8; d <<= 1) {
            const bool in = j + d < G8;
#pragma unroll
            for (int c = 0; c < CNT; c++) {
                const double t = g8_down(v[c], d);
                v[c] += in ? t : 0.0;
            }
        }
    }
};
template <> struct G8Topo<TopoMG400> {
    static constexpr int MAXDEPTH = 4;
    __host__ __device__ static constexpr int depth(int i) { return i < 5 ? i : i - 4; }
    __host__ __device__ static constexpr bool is_anc(int k, int j)
    {
        return k == j || k == 0 || (k < j && ((k >= 1 && j <= 4) || (k >= 5 && j >= 5)));
    }
    // tree 0 - {1-2-3-4, 5-6-7}: suffix sums inside the two chains, then the root takes both chain heads
    template <int CNT> TGD static void subtree_sum(double (&v)[CNT], int j)
    {
        const bool a1 = (j >= 1 && j <= 3) || (j >= 5 && j <= 6);
        const bool a2 = (j >= 1 && j <= 2) || j == 5;
#pragma unroll
        for (int c = 0; c < CNT; c++) {
            double t = g8_down(v[c], 1);
            v[c] += a1 ? t : 0.0;
            t = g8_down(v[c], 2);
            v[c] += a2 ? t : 0.0;
            const double h1 = g8_get(v[c], 1), h5 = g8_get(v[c], 5);
            v[c] += j == 0 ? h1 + h5 : 0.0;
        }
    }
};

// per-lane constants of body j (read once per kernel)
struct G8Body {
    double jpos[3], axis[3], mass, com[3], inertia[6];
    int par;        // parent lane, -1 for the root and for idle lanes
    int sub0, sub1; // this body's mass-carrying URDF links (per-link damping)
};

// sub-link table in shared memory: [TG_MAXSUB][G8_SUBW] doubles: com(3) rot(9) inertia(3) mass(1) (+1 pad: conflict-free rows)
#define G8_SUBW 17

template <class T>
TGD void g8_load_body(const TgArm& arm, int j, G8Body& c)
{
    const bool live = j < T::NB;
    const int jj = live ? j : 0;
#pragma unroll
    for (int k = 0; k < 3; k++) { c.jpos[k] = live ? arm.jpos[jj][k] : 0.0; c.axis[k] = live ? arm.axis[jj][k] : 0.0; c.com[k] = live ? arm.com[jj][k] : 0.0; }
#pragma unroll
    for (int k = 0; k < 6; k++) c.inertia[k] = live ? arm.inertia[jj][k] : 0.0;
    c.mass = live ? arm.mass[jj] : 0.0;
    c.par = live ? T::parent(jj) : -1;
    c.sub0 = live ? arm.sub_start[jj] : 0;
    c.sub1 = live ? arm.sub_start[jj + 1] : 0;
    if (T::NB < G8) {
        // Idle lanes lend a hand with the per-link damping: a body's sub-links beyond its first move (last first) to idle
        // lanes, which sit in the tree as massless leaves rigidly below that body (zero axis: same frame, same velocity, no
        // joint), so the scans hand them the body's kinematics and the suffix sums hand the body their wrench.  UR5: wrist 3
        // carries three links (wrist, sensor body, tip) and lanes 6, 7 are idle - no lane loops over sub-links at all.
        int extra = 0, taken = 0; // extras handed out so far; those of THIS lane's body
#pragma unroll
        for (int bdy = 0; bdy < T::NB; bdy++) {
            const int s0 = arm.sub_start[bdy], s1 = arm.sub_start[bdy + 1];
            for (int sx = s1 - 1; sx > s0; sx--) {
                const int lane_x = T::NB + extra;
                if (lane_x >= G8) break;
                if (j == lane_x) { c.par = bdy; c.sub0 = sx; c.sub1 = sx + 1; }
                if (j == bdy) taken++;
                extra++;
            }
        }
        if (live) c.sub1 -= taken; // the body keeps the head of its range
    }
}

// the longest sub-link range any lane of the warp carries (trip count of the damping loop)
TGD int g8_max_sub(const G8Body& c)
{
    int m = c.sub1 - c.sub0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) m = max(m, __shfl_xor_sync(G8_FULL, m, d));
    return m;
}

TGD void g8_fill_sub_table(const TgArm& arm, double* s_sub)
{
    for (int t = threadIdx.x; t < TG_MAXSUB * 16; t += blockDim.x) {
        const int s = t >> 4, c = t & 15;
        double v;
        if (c < 3) v = arm.sub_com[s][c];
        else if (c < 12) v = arm.sub_rot[s][c - 3];
        else if (c < 15) v = arm.sub_inertia[s][c - 12];
        else v = arm.sub_mass[s];
        s_sub[s * G8_SUBW + c] = v;
    }
}

// Cholesky factor of the SPD matrix M (upper part given as Mf[k][j], k <= j): L[i][k] for i > k, dinv[k] = 1 / L[k][k].
// Same operation order as spd_inverse (tg_dyn.cuh).
template <int NB>
TGD void g8_cholesky(const double (&Mf)[NB][NB], double (&L)[NB][NB], double (&dinv)[NB])
{
#pragma unroll
    for (int j = 0; j < NB; j++) {
        double d = Mf[j][j];
#pragma unroll
        for (int k2 = 0; k2 < j; k2++) d -= L[j][k2] * L[j][k2];
        const double inv = rsqrt(d);
        dinv[j] = inv;
        L[j][j] = d * inv;
#pragma unroll
        for (int i = j + 1; i < NB; i++) {
            double s = Mf[j][i];
#pragma unroll
            for (int k2 = 0; k2 < j; k2++) s -= L[i][k2] * L[j][k2];
            L[i][j] = s * inv;
        }
    }
}

// forward kinematics of the group: T_j = T_parent o (Rot(axis, q), jpos), composed by pointer jumping (3 rounds: depth <= 8).
// R, p: world rotation of body j's frame and world position of its joint origin (Kin<>::R, Kin<>::p of tg_dyn.cuh).
TGD void g8_fk(const G8Body& bc, int j, double sn, double cs, double (&R)[9], double (&p)[3])
{
    {
        const double s = sn, c = cs, ax = bc.axis[0], ay = bc.axis[1], az = bc.axis[2], t1 = 1.0 - c;
        R[0] = t1 * ax * ax + c;      R[1] = t1 * ax * ay - s * az; R[2] = t1 * ax * az + s * ay;
        R[3] = t1 * ax * ay + s * az; R[4] = t1 * ay * ay + c;      R[5] = t1 * ay * az - s * ax;
        R[6] = t1 * ax * az - s * ay; R[7] = t1 * ay * az + s * ax; R[8] = t1 * az * az + c;
        p[0] = bc.jpos[0]; p[1] = bc.jpos[1]; p[2] = bc.jpos[2];
    }
    int anc = bc.par;
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const int src = anc < 0 ? j : anc;
        double Ra[9], pa[3];
#pragma unroll
        for (int k = 0; k < 9; k++) Ra[k] = g8_get(R[k], src);
#pragma unroll
        for (int k = 0; k < 3; k++) pa[k] = g8_get(p[k], src);
        const int anc2 = g8_geti(anc, src);
        if (anc >= 0) {
            double t[3];
            m3mulv(t, Ra, p);
            p[0] = pa[0] + t[0]; p[1] = pa[1] + t[1]; p[2] = pa[2] + t[2];
            m3mul(R, Ra, R);
            anc = anc2;
        }
    }
}

// TCP pose (getLinkState(tcp)[0:2]) and camera frame of the group from the lanes' body frames: computed on the lane that carries
// the body, handed to every lane.  tp[3], tq[4], cam[12] as tcp_world / write_camera give them.
TGD void g8_tcp_and_camera(const TgArm& arm, const double (&R)[9], const double (&p)[3], double* tp, double* tq, double* cam)
{
    double pos[3], Rt[9], t[3], q4[4], c12[12];
    m3mulv(t, R, arm.tcp_pos);
    pos[0] = p[0] + t[0]; pos[1] = p[1] + t[1]; pos[2] = p[2] + t[2];
    m3mul(Rt, R, arm.tcp_rot);
    quat_from_mat(q4, Rt);
#pragma unroll
    for (int k = 0; k < 3; k++) tp[k] = g8_get(pos[k], arm.tcp_body);
#pragma unroll
    for (int k = 0; k < 4; k++) tq[k] = g8_get(q4[k], arm.tcp_body);
    if (cam) {
        m3mulv(t, R, arm.cam_pos);
        pos[0] = p[0] + t[0]; pos[1] = p[1] + t[1]; pos[2] = p[2] + t[2];
        m3mul(Rt, R, arm.cam_rot);
        camera_from_frame(pos, Rt, c12);
#pragma unroll
        for (int k = 0; k < 12; k++) cam[k] = g8_get(c12[k], arm.cam_body);
    }
}

// state of one lane between substeps
struct G8Lane {
    double q, qd, s, c; // joint angle, rate, sin, cos
};

// One Robot.step_sim() (gravity compensation + stepSimulation with motor rows only) for the four envs of a warp.
// mode / kp / kd / max_force as in Motors<>; tpos / tvel = this lane's motor target.  `live` = the group carries an env.
template <class T>
TGD void g8_substep(const TgPhysics& ph, const G8Body& bc, const double* __restrict__ s_sub, int max_sub, int j, bool live, G8Lane& st,
                    int mode, double kp, double kd, double max_force, double tpos, double tvel)
{
    constexpr int NB = T::NB;
    using Tp = G8Topo<T>;
    double R[9], p[3];
    g8_fk(bc, j, st.s, st.c, R, p);
    double a[3], lin[3];
    m3mulv(a, R, bc.axis);
    v3cross(lin, p, a); // velocity of the origin-coincident point for unit joint rate

    // ---- body velocities about the world origin: prefix sums of a qd, lin qd over ancestors-or-self
    double w[3], vO[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { w[k] = a[k] * st.qd; vO[k] = lin[k] * st.qd; }
    {
        int anc = bc.par;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const int src = anc < 0 ? j : anc;
            double wa[3], va[3];
#pragma unroll
            for (int k = 0; k < 3; k++) { wa[k] = g8_get(w[k], src); va[k] = g8_get(vO[k], src); }
            const int anc2 = g8_geti(anc, src);
            if (anc >= 0) {
#pragma unroll
                for (int k = 0; k < 3; k++) { w[k] += wa[k]; vO[k] += va[k]; }
                anc = anc2;
            }
        }
    }

    // ---- lane-local: spatial inertia about the world origin (body_inertias) and bullet's per-link damping wrench (damping_forces)
    double acc[16]; // [0..5] damping wrench (N, F) about the origin, [6..15] spatial inertia m, h(3), I(6): summed over subtrees below
    {
        double cw[3], t[3];
        m3mulv(t, R, bc.com);
        cw[0] = p[0] + t[0]; cw[1] = p[1] + t[1]; cw[2] = p[2] + t[2];
        const double m = bc.mass;
        const double* I = bc.inertia;
        double Ic[9] = {I[0], I[1], I[2], I[1], I[3], I[4], I[2], I[4], I[5]}, RI[9];
        m3mul(RI, R, Ic);
        const double xx = RI[0] * R[0] + RI[1] * R[1] + RI[2] * R[2];
        const double xy = RI[0] * R[3] + RI[1] * R[4] + RI[2] * R[5];
        const double xz = RI[0] * R[6] + RI[1] * R[7] + RI[2] * R[8];
        const double yy = RI[3] * R[3] + RI[4] * R[4] + RI[5] * R[5];
        const double yz = RI[3] * R[6] + RI[4] * R[7] + RI[5] * R[8];
        const double zz = RI[6] * R[6] + RI[7] * R[7] + RI[8] * R[8];
        const double c2 = v3dot(cw, cw);
        acc[6] = m;
        acc[7] = m * cw[0]; acc[8] = m * cw[1]; acc[9] = m * cw[2];
        acc[10] = xx + m * (c2 - cw[0] * cw[0]);
        acc[11] = xy - m * cw[0] * cw[1];
        acc[12] = xz - m * cw[0] * cw[2];
        acc[13] = yy + m * (c2 - cw[1] * cw[1]);
        acc[14] = yz - m * cw[1] * cw[2];
        acc[15] = zz + m * (c2 - cw[2] * cw[2]);
    }
    {
        const double ka = ph.ang_damping * (1.0 + (double)sqrtf((float)v3dot(w, w)));
        double wb[3], Nb[3] = {0, 0, 0}, Na[3] = {0, 0, 0}, Fa[3] = {0, 0, 0};
        m3tmulv(wb, R, w);
#pragma unroll 1
        for (int it = 0; it < max_sub; it++) {
            const int s = bc.sub0 + it;
            if (s < bc.sub1) {
                const double* sd = s_sub + s * G8_SUBW;
                const double scom[3] = {sd[0], sd[1], sd[2]};
                double srot[9];
#pragma unroll
                for (int k = 0; k < 9; k++) srot[k] = sd[3 + k];
                double t[3], x[3], v[3], wl[3], nl[3], nb[3], xf[3];
                m3mulv(t, R, scom);
                x[0] = p[0] + t[0]; x[1] = p[1] + t[1]; x[2] = p[2] + t[2];
                v3cross(v, w, x);
                v[0] += vO[0]; v[1] += vO[1]; v[2] += vO[2];
                const double kl = ph.lin_damping * (1.0 + (double)sqrtf((float)v3dot(v, v)));
                m3tmulv(wl, srot, wb);
#pragma unroll
                for (int k = 0; k < 3; k++) nl[k] = -sd[12 + k] * wl[k] * ka;
                m3mulv(nb, srot, nl);
                const double ms = sd[15];
                const double f[3] = {-ms * v[0] * kl, -ms * v[1] * kl, -ms * v[2] * kl};
                v3cross(xf, x, f);
#pragma unroll
                for (int k = 0; k < 3; k++) { Nb[k] += nb[k]; Na[k] += xf[k]; Fa[k] += f[k]; }
            }
        }
        double nwv[3];
        m3mulv(nwv, R, Nb);
#pragma unroll
        for (int k = 0; k < 3; k++) { acc[k] = nwv[k] + Na[k]; acc[3 + k] = Fa[k]; }
    }
    // ---- suffix sums over the tree: subtree damping wrench, composite inertia
    Tp::template subtree_sum<16>(acc, j);
    double tau = v3dot(a, acc) + v3dot(lin, acc + 3) - ph.joint_damping * st.qd;

    // ---- CRBA: (n, f) = composite inertia of subtree j applied to joint j's motion; M[k][j] = S_k . (n, f) for ancestors k
    double Mcol[Tp::MAXDEPTH + 1];
    {
        double n[3], f[3];
        SpI sp;
        sp.m = acc[6];
#pragma unroll
        for (int k = 0; k < 3; k++) sp.h[k] = acc[7 + k];
#pragma unroll
        for (int k = 0; k < 6; k++) sp.I[k] = acc[10 + k];
        spi_apply(sp, a, lin, n, f);
        Mcol[0] = v3dot(a, n) + v3dot(lin, f);
        int anc = bc.par;
#pragma unroll
        for (int d = 1; d <= Tp::MAXDEPTH; d++) {
            const int src = anc < 0 ? j : anc;
            double ak[3], lk[3];
#pragma unroll
            for (int k = 0; k < 3; k++) { ak[k] = g8_get(a[k], src); lk[k] = g8_get(lin[k], src); }
            const int anc2 = g8_geti(bc.par, src);
            Mcol[d] = anc >= 0 ? v3dot(ak, n) + v3dot(lk, f) : 0.0;
            anc = anc >= 0 ? anc2 : -1;
        }
    }
    // ---- every lane gathers the whole matrix, factorises it, and solves for its own column of A = M^-1
    double A[NB]; // A[r] = A[r][j] = A[j][r]
    {
        double Mf[NB][NB], L[NB][NB], dinv[NB];
#pragma unroll
        for (int jj = 0; jj < NB; jj++)
#pragma unroll
            for (int k = 0; k <= jj; k++)
                Mf[k][jj] = Tp::is_anc(k, jj) ? g8_get(Mcol[Tp::depth(jj) - Tp::depth(k)], jj) : 0.0;
        g8_cholesky<NB>(Mf, L, dinv);
        double y[NB];
#pragma unroll
        for (int k = 0; k < NB; k++) {      // L y = e_j
            double s = j == k ? 1.0 : 0.0;
#pragma unroll
            for (int m = 0; m < k; m++) s -= L[k][m] * y[m];
            y[k] = s * dinv[k];
        }
#pragma unroll
        for (int k = NB - 1; k >= 0; k--) { // L^T x = y
            double s = y[k];
#pragma unroll
            for (int m = k + 1; m < NB; m++) s -= L[m][k] * A[m];
            A[k] = s * dinv[k];
        }
    }
    // ---- unconstrained velocity update: qdd_j = sum_k A[j][k] tau_k
    double Ajj = 0.0;
    {
        double qdd = 0.0;
#pragma unroll
        for (int k = 0; k < NB; k++) { qdd += A[k] * g8_get(tau, k); Ajj = j == k ? A[k] : Ajj; }
        st.qd += ph.dt * qdd;
    }

    // ---- motor rows (J = e_i, response column A[:, i], |impulse| <= force * dt) and projected Gauss-Seidel
    const double lim = max_force * ph.dt;
    const double dinv_m = Ajj > 2.2204460492503131e-16 ? 1.0 / Ajj : 0.0;
    double rhs;
    {
        const double v = st.qd;
        const double pos_stab = mode == 1 ? kp * ((tpos - st.q) / ph.dt) : 0.0;
        const double rhs_v = pos_stab + v + kd * (tvel - v);
        rhs = (rhs_v - v) * dinv_m;
    }
    double applied = 0.0, dv = 0.0;
    if (lim != 0.0) {
        bool active = live;
        const unsigned gshift = (threadIdx.x & 24);  // bit offset of this group's lanes in a warp ballot
#pragma unroll 1
        for (int it = 0; it < ph.solver_iters; it++) {
            if (__ballot_sync(G8_FULL, active) == 0u) break;
            double resid = 0.0;
            auto row = [&](int r) {
                double delta = rhs - dv * dinv_m;
                const double sum = applied + delta;
                const bool lo = sum < -lim, hi = sum > lim;
                delta = lo ? (-lim - applied) : (hi ? (lim - applied) : delta);
                const double napp = lo ? -lim : (hi ? lim : sum);
                const bool mine = active && j == r;
                applied = mine ? napp : applied;
                const double dvel = delta * Ajj;
                resid = mine ? dvel * dvel : resid;
                const double dl = g8_get(active ? delta : 0.0, r);
                dv += A[r] * dl;
            };
            if (it & 1) {
#pragma unroll
                for (int r = 0; r < NB; r++) row(r);
            } else {
#pragma unroll
                for (int r = NB - 1; r >= 0; r--) row(r);
            }
            const unsigned big = __ballot_sync(G8_FULL, active && resid > ph.solver_residual_threshold);
            if (((big >> gshift) & 0xffu) == 0u) active = false;
        }
    }
    st.qd += dv;
    const double d = ph.dt * st.qd;
    st.q += d;
    double sc[2] = {st.s, st.c};
    sc_advance(sc, st.q, d);
    st.s = sc[0]; st.c = sc[1];
}

// g8_substep with the constrained pole of object_balance in the world - substep_obj (tg_dyn.cuh) on the group's lanes (a copy of
// g8_substep with the additions, so that the motor-row kernels' code stays exactly what it was).  The pole's
// state and everything derived from it alone is REPLICATED on the eight lanes (every lane computes the same numbers: no
// divergence, no broadcast); what couples it to the arm - the three point-to-point rows' arm Jacobian / response entries, one
// per lane, and their dot products with the lanes' velocity changes - is a shuffle sum over the group per row update.
// on_path: body j is the TCP body or one of its ancestors (computed once per env step).
template <class T>
TGD void g8_substep_obj(const TgPhysics& ph, const G8Body& bc, const double* __restrict__ s_sub, int max_sub, int j, bool live, G8Lane& st,
                    int mode, double kp, double kd, double max_force, double tpos, double tvel,
                    const TgArm* armp, const TgTask* taskp, ObjState* op, bool on_path)
{
    constexpr bool OBJ = true;
    constexpr int NB = T::NB;
    using Tp = G8Topo<T>;
    double R[9], p[3];
    g8_fk(bc, j, st.s, st.c, R, p);
    double a[3], lin[3];
    m3mulv(a, R, bc.axis);
    v3cross(lin, p, a); // velocity of the origin-coincident point for unit joint rate
    double pa[3] = {0, 0, 0};   // OBJ: the constraint's pivot on the arm = the TCP point (tcp_world)
    if (OBJ) {
        double t[3], pos[3];
        m3mulv(t, R, armp->tcp_pos);
        pos[0] = p[0] + t[0]; pos[1] = p[1] + t[1]; pos[2] = p[2] + t[2];
#pragma unroll
        for (int k = 0; k < 3; k++) pa[k] = g8_get(pos[k], armp->tcp_body);
    }

    // ---- body velocities about the world origin: prefix sums of a qd, lin qd over ancestors-or-self
    double w[3], vO[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { w[k] = a[k] * st.qd; vO[k] = lin[k] * st.qd; }
    {
        int anc = bc.par;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const int src = anc < 0 ? j : anc;
            double wa[3], va[3];
#pragma unroll
            for (int k = 0; k < 3; k++) { wa[k] = g8_get(w[k], src); va[k] = g8_get(vO[k], src); }
            const int anc2 = g8_geti(anc, src);
            if (anc >= 0) {
#pragma unroll
                for (int k = 0; k < 3; k++) { w[k] += wa[k]; vO[k] += va[k]; }
                anc = anc2;
            }
        }
    }

    // ---- lane-local: spatial inertia about the world origin (body_inertias) and bullet's per-link damping wrench (damping_forces)
    double acc[16]; // [0..5] damping wrench (N, F) about the origin, [6..15] spatial inertia m, h(3), I(6): summed over subtrees below
    {
        double cw[3], t[3];
        m3mulv(t, R, bc.com);
        cw[0] = p[0] + t[0]; cw[1] = p[1] + t[1]; cw[2] = p[2] + t[2];
        const double m = bc.mass;
        const double* I = bc.inertia;
        double Ic[9] = {I[0], I[1], I[2], I[1], I[3], I[4], I[2], I[4], I[5]}, RI[9];
        m3mul(RI, R, Ic);
        const double xx = RI[0] * R[0] + RI[1] * R[1] + RI[2] * R[2];
        const double xy = RI[0] * R[3] + RI[1] * R[4] + RI[2] * R[5];
        const double xz = RI[0] * R[6] + RI[1] * R[7] + RI[2] * R[8];
        const double yy = RI[3] * R[3] + RI[4] * R[4] + RI[5] * R[5];
        const double yz = RI[3] * R[6] + RI[4] * R[7] + RI[5] * R[8];
        const double zz = RI[6] * R[6] + RI[7] * R[7] + RI[8] * R[8];
        const double c2 = v3dot(cw, cw);
        acc[6] = m;
        acc[7] = m * cw[0]; acc[8] = m * cw[1]; acc[9] = m * cw[2];
        acc[10] = xx + m * (c2 - cw[0] * cw[0]);
        acc[11] = xy - m * cw[0] * cw[1];
        acc[12] = xz - m * cw[0] * cw[2];
        acc[13] = yy + m * (c2 - cw[1] * cw[1]);
        acc[14] = yz - m * cw[1] * cw[2];
        acc[15] = zz + m * (c2 - cw[2] * cw[2]);
    }
    {
        const double ka = ph.ang_damping * (1.0 + (double)sqrtf((float)v3dot(w, w)));
        double wb[3], Nb[3] = {0, 0, 0}, Na[3] = {0, 0, 0}, Fa[3] = {0, 0, 0};
        m3tmulv(wb, R, w);
#pragma unroll 1
        for (int it = 0; it < max_sub; it++) {
            const int s = bc.sub0 + it;
            if (s < bc.sub1) {
                const double* sd = s_sub + s * G8_SUBW;
                const double scom[3] = {sd[0], sd[1], sd[2]};
                double srot[9];
#pragma unroll
                for (int k = 0; k < 9; k++) srot[k] = sd[3 + k];
                double t[3], x[3], v[3], wl[3], nl[3], nb[3], xf[3];
                m3mulv(t, R, scom);
                x[0] = p[0] + t[0]; x[1] = p[1] + t[1]; x[2] = p[2] + t[2];
                v3cross(v, w, x);
                v[0] += vO[0]; v[1] += vO[1]; v[2] += vO[2];
                const double kl = ph.lin_damping * (1.0 + (double)sqrtf((float)v3dot(v, v)));
                m3tmulv(wl, srot, wb);
#pragma unroll
                for (int k = 0; k < 3; k++) nl[k] = -sd[12 + k] * wl[k] * ka;
                m3mulv(nb, srot, nl);
                const double ms = sd[15];
                const double f[3] = {-ms * v[0] * kl, -ms * v[1] * kl, -ms * v[2] * kl};
                v3cross(xf, x, f);
#pragma unroll
                for (int k = 0; k < 3; k++) { Nb[k] += nb[k]; Na[k] += xf[k]; Fa[k] += f[k]; }
            }
        }
        double nwv[3];
        m3mulv(nwv, R, Nb);
#pragma unroll
        for (int k = 0; k < 3; k++) { acc[k] = nwv[k] + Na[k]; acc[3 + k] = Fa[k]; }
    }
    // ---- suffix sums over the tree: subtree damping wrench, composite inertia
    Tp::template subtree_sum<16>(acc, j);
    double tau = v3dot(a, acc) + v3dot(lin, acc + 3) - ph.joint_damping * st.qd;

    // ---- CRBA: (n, f) = composite inertia of subtree j applied to joint j's motion; M[k][j] = S_k . (n, f) for ancestors k
    double Mcol[Tp::MAXDEPTH + 1];
    {
        double n[3], f[3];
        SpI sp;
        sp.m = acc[6];
#pragma unroll
        for (int k = 0; k < 3; k++) sp.h[k] = acc[7 + k];
#pragma unroll
        for (int k = 0; k < 6; k++) sp.I[k] = acc[10 + k];
        spi_apply(sp, a, lin, n, f);
        Mcol[0] = v3dot(a, n) + v3dot(lin, f);
        int anc = bc.par;
#pragma unroll
        for (int d = 1; d <= Tp::MAXDEPTH; d++) {
            const int src = anc < 0 ? j : anc;
            double ak[3], lk[3];
#pragma unroll
            for (int k = 0; k < 3; k++) { ak[k] = g8_get(a[k], src); lk[k] = g8_get(lin[k], src); }
            const int anc2 = g8_geti(bc.par, src);
            Mcol[d] = anc >= 0 ? v3dot(ak, n) + v3dot(lk, f) : 0.0;
            anc = anc >= 0 ? anc2 : -1;
        }
    }
    // ---- every lane gathers the whole matrix, factorises it, and solves for its own column of A = M^-1
    double A[NB]; // A[r] = A[r][j] = A[j][r]
    {
        double Mf[NB][NB], L[NB][NB], dinv[NB];
#pragma unroll
        for (int jj = 0; jj < NB; jj++)
#pragma unroll
            for (int k = 0; k <= jj; k++)
                Mf[k][jj] = Tp::is_anc(k, jj) ? g8_get(Mcol[Tp::depth(jj) - Tp::depth(k)], jj) : 0.0;
        g8_cholesky<NB>(Mf, L, dinv);
        double y[NB];
#pragma unroll
        for (int k = 0; k < NB; k++) {      // L y = e_j
            double s = j == k ? 1.0 : 0.0;
#pragma unroll
            for (int m = 0; m < k; m++) s -= L[k][m] * y[m];
            y[k] = s * dinv[k];
        }
#pragma unroll
        for (int k = NB - 1; k >= 0; k--) { // L^T x = y
            double s = y[k];
#pragma unroll
            for (int m = k + 1; m < NB; m++) s -= L[m][k] * A[m];
            A[k] = s * dinv[k];
        }
    }
    // ---- unconstrained velocity update: qdd_j = sum_k A[j][k] tau_k
    double Ajj = 0.0;
    {
        double qdd = 0.0;
#pragma unroll
        for (int k = 0; k < NB; k++) { qdd += A[k] * g8_get(tau, k); Ajj = j == k ? A[k] : Ajj; }
        st.qd += ph.dt * qdd;
    }

    // ---- OBJ: the pole's unconstrained update and the three point-to-point rows (substep_obj, tg_dyn.cuh)
    double Rb[9], dw[3], cw[3], vc[3] = {0, 0, 0}, Iinv[3];
    double jr[3] = {0, 0, 0}, ur[3] = {0, 0, 0};   // this lane's entry of the rows' arm Jacobian / response
    double jba[3][3], uba[3][3], rhs_p[3], dinv_p[3], diag_p[3], app_p[3] = {0, 0, 0};
    double dvl[3] = {0, 0, 0}, dva[3] = {0, 0, 0};
    auto g8_sum = [](double v) {
        v += __shfl_xor_sync(G8_FULL, v, 1, G8); v += __shfl_xor_sync(G8_FULL, v, 2, G8); v += __shfl_xor_sync(G8_FULL, v, 4, G8);
        return v;
    };
    if (OBJ) {
        const TgTask& task = *taskp;
        ObjState& o = *op;
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
        // arm side: row i's Jacobian entry of joint j = -(a_j x (pa - p_j))[i] = -(lin_j + a_j x pa)[i] on the TCP's ancestors
        {
            double t[3];
            v3cross(t, a, pa);
#pragma unroll
            for (int i = 0; i < 3; i++) jr[i] = (on_path && j < NB) ? -(lin[i] + t[i]) : 0.0;
        }
        double pb[3], rb[3];
        {
            double t[3];
            const double pl[3] = {0.0, 0.0, o.pivot_z};
            m3mulv(t, Rb, pl);
#pragma unroll
            for (int c = 0; c < 3; c++) { pb[c] = o.pos[c] + t[c]; rb[c] = pb[c] - cw[c]; }
        }
#pragma unroll
        for (int i = 0; i < 3; i++) {
            // response of joint j to row i: sum_e A[j][e] jr_i[e] (A[e] on this lane = A[e][j])
            double u = 0.0;
#pragma unroll
            for (int e2 = 0; e2 < NB; e2++) u += A[e2] * g8_get(jr[i], e2);
            ur[i] = j < NB ? u : 0.0;
            double denom = g8_sum(jr[i] * ur[i]);
            double rel = g8_sum(jr[i] * st.qd);
            double nB[3] = {0, 0, 0};
            nB[i] = 1.0;
            double jl[3], ul[3];
            v3cross(jba[i], rb, nB);
            m3tmulv(jl, Rb, jba[i]);
#pragma unroll
            for (int c = 0; c < 3; c++) ul[c] = jl[c] * Iinv[c];
            m3mulv(uba[i], Rb, ul);
            denom += 1.0 / task.obj_mass + v3dot(jba[i], uba[i]);
            rel += vc[i] + v3dot(jba[i], o.omg);
            diag_p[i] = denom;
            dinv_p[i] = denom > 2.2204460492503131e-16 ? 1.0 / denom : 0.0;
            const double pos_error = -(pa[i] - pb[i]);               // (pivotA - pivotB) . (-e_i)
            const double positional = -pos_error * task.p2p_erp / ph.dt;
            rhs_p[i] = (positional - rel) * dinv_p[i];
        }
    }

    // ---- motor rows (J = e_i, response column A[:, i], |impulse| <= force * dt) and projected Gauss-Seidel
    const double lim = max_force * ph.dt;
    const double dinv_m = Ajj > 2.2204460492503131e-16 ? 1.0 / Ajj : 0.0;
    double rhs;
    {
        const double v = st.qd;
        const double pos_stab = mode == 1 ? kp * ((tpos - st.q) / ph.dt) : 0.0;
        const double rhs_v = pos_stab + v + kd * (tvel - v);
        rhs = (rhs_v - v) * dinv_m;
    }
    double applied = 0.0, dv = 0.0;
    if (lim != 0.0 || OBJ) {
        bool active = live;
        const unsigned gshift = (threadIdx.x & 24);  // bit offset of this group's lanes in a warp ballot
#pragma unroll 1
        for (int it = 0; it < ph.solver_iters; it++) {
            if (__ballot_sync(G8_FULL, active) == 0u) break;
            double resid = 0.0;
            auto row = [&](int r) {
                double delta = rhs - dv * dinv_m;
                const double sum = applied + delta;
                const bool lo = sum < -lim, hi = sum > lim;
                delta = lo ? (-lim - applied) : (hi ? (lim - applied) : delta);
                const double napp = lo ? -lim : (hi ? lim : sum);
                const bool mine = active && j == r;
                applied = mine ? napp : applied;
                const double dvel = delta * Ajj;
                resid = mine ? dvel * dvel : resid;
                const double dl = g8_get(active ? delta : 0.0, r);
                dv += A[r] * dl;
            };
            // point-to-point row i: every lane of the group computes the same impulse (the arm's part of the row's velocity is a
            // shuffle sum of the lanes' entries), then applies its own entry of the response
            double resid_p = 0.0;
            auto row_p = [&](int i) {
                const double plim = taskp->p2p_max_impulse;
                const double dot = (dvl[i] + v3dot(jba[i], dva)) + g8_sum(jr[i] * dv);
                double delta = rhs_p[i] - dot * dinv_p[i];
                const double sum = app_p[i] + delta;
                const bool lo = sum < -plim, hi = sum > plim;
                delta = lo ? (-plim - app_p[i]) : (hi ? (plim - app_p[i]) : delta);
                const double napp = lo ? -plim : (hi ? plim : sum);
                delta = active ? delta : 0.0;
                app_p[i] = active ? napp : app_p[i];
                dv += ur[i] * delta;
                dvl[i] += delta / taskp->obj_mass;
#pragma unroll
                for (int c = 0; c < 3; c++) dva[c] += uba[i][c] * delta;
                const double dvel = delta * diag_p[i];
                resid_p = fmax(resid_p, dvel * dvel);
            };
            if (it & 1) {
                if (lim != 0.0) {
#pragma unroll
                    for (int r = 0; r < NB; r++) row(r);
                }
                if (OBJ) {
#pragma unroll
                    for (int i = 0; i < 3; i++) row_p(i);
                }
            } else {
                if (OBJ) {
#pragma unroll
                    for (int i = 2; i >= 0; i--) row_p(i);
                }
                if (lim != 0.0) {
#pragma unroll
                    for (int r = NB - 1; r >= 0; r--) row(r);
                }
            }
            if (OBJ) resid = fmax(resid, resid_p);
            const unsigned big = __ballot_sync(G8_FULL, active && resid > ph.solver_residual_threshold);
            if (((big >> gshift) & 0xffu) == 0u) active = false;
        }
    }
    st.qd += dv;
    const double d = ph.dt * st.qd;
    st.q += d;
    double sc[2] = {st.s, st.c};
    sc_advance(sc, st.q, d);
    st.s = sc[0]; st.c = sc[1];
    if (OBJ) {
        // the pole: apply the impulses, integrate (exponential map), back to the base-link COM (replicated on the lanes)
        const TgTask& task = *taskp;
        ObjState& o = *op;
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
    }
}

// step kernel, 8 lanes per env: edge_follow, surface_follow (motor rows only) and object_balance (+ the pole on its three
// point-to-point rows), TCP_velocity_control, gravity compensation on.  Blocks >= b.step_blocks keep the standby role of step_kernel.
template <class T, int TASK>
__global__ void __launch_bounds__(128)
step_kernel_g8(const __grid_constant__ TgArm arm, const __grid_constant__ TgPhysics ph, const __grid_constant__ TgTask task,
               EnvBuffers b, const float* __restrict__ actions, float* __restrict__ reward, unsigned char* __restrict__ done, int autoreset)
{
    constexpr int NB = T::NB;
    if ((int)blockIdx.x >= b.step_blocks) {
        standby_role<T>(arm, ph, task, b, b.step_blocks, false);
        return;
    }
    __shared__ double s_sub[TG_MAXSUB * G8_SUBW];
    g8_fill_sub_table(arm, s_sub);
    __syncthreads();
    const int j = threadIdx.x & (G8 - 1);
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const bool live = e < b.n;
    G8Body bc;
    g8_load_body<T>(arm, j, bc);
    const int max_sub = g8_max_sub(bc);

    G8Lane st;
    st.q = 0.0; st.qd = 0.0;
    if (live && j < NB) { st.q = b.q[(size_t)j * b.n + e]; st.qd = b.qd[(size_t)j * b.n + e]; }
    sincos(st.q, &st.s, &st.c);

    // ---- head of the step.  Lane-parallel: FK, TCP pose, the TCP Jacobian (column j on lane j).  On the group's first lane:
    // action encoding, TCP limits, the 6 x NB solve (env_prologue_core, the one-thread formulation, once per env step).
    constexpr bool balance = TASK == TG_TASK_OBJECT_BALANCE;
    double q[NB], qd[NB], tvel = 0.0;
    bool on_path_b = false; // (object_balance keeps it for the substeps)
    {
        bool on_path = false; // body j is the TCP body or one of its ancestors
        double R[9], p[3], tp[3], tq[4];
        g8_fk(bc, j, st.s, st.c, R, p);
        g8_tcp_and_camera(arm, R, p, tp, tq, nullptr);
        // geometric Jacobian column of joint j (tcp_jacobian): a_j x (tcp - p_j), a_j for the ancestors-or-self of the TCP body
        double ax[3], col[6];
        m3mulv(ax, R, bc.axis);
        const double r[3] = {tp[0] - p[0], tp[1] - p[1], tp[2] - p[2]};
        v3cross(col, ax, r);
        col[3] = ax[0]; col[4] = ax[1]; col[5] = ax[2];
        {
            int a = arm.tcp_body;
#pragma unroll
            for (int k = 0; k < G8; k++) { on_path = on_path || a == j; const int pa = g8_geti(bc.par, a < 0 ? 0 : a); a = a >= 0 ? pa : -1; }
        }
        double J[6][NB];
#pragma unroll
        for (int c = 0; c < 6; c++) {
            const double v = on_path && j < NB ? col[c] : 0.0;
#pragma unroll
            for (int i = 0; i < NB; i++) J[c][i] = g8_get(v, i);
        }
        double tv_all[NB];
#pragma unroll
        for (int i = 0; i < NB; i++) tv_all[i] = 0.0;
        if (live && j == 0) {
            double v[6];
            Motors<NB> mot;
            env_prologue_core<T, TASK>(arm, ph, task, b, e, actions, tp, tq, J, v, mot);
#pragma unroll
            for (int i = 0; i < NB; i++) tv_all[i] = mot.target_vel[i];
        }
#pragma unroll
        for (int i = 0; i < NB; i++) {
            const double t = g8_get(tv_all[i], 0);
            tvel = j == i ? t : tvel;
        }
        if (balance) on_path_b = on_path;
    }
    ObjState ob_b;   // object_balance only (the other tasks never touch it)
    if (balance) {
        ObjState& ob = ob_b;
        // the pole's state, replicated on the group's lanes (a group without an env carries a benign dummy)
#pragma unroll
        for (int c = 0; c < 3; c++) { ob.pos[c] = 0.0; ob.vel[c] = 0.0; ob.omg[c] = 0.0; ob.ext_pos[c] = 0.0; ob.quat[c] = 0.0; }
        ob.quat[3] = 1.0; ob.ext_pending = 0; ob.grav_z = 0.0; ob.pivot_z = 0.0; ob.mass = 0.0;
        if (live) obj_load(b.obj + (size_t)e * 13, b.obj_ext + (size_t)e * 4, b.grav[e], b.embed[e], task, ob);
#pragma unroll 1
        for (int s = 0; s < ph.substeps; s++)
            g8_substep_obj<T>(ph, bc, s_sub, max_sub, j, live, st, 0, 0.0, ph.vel_gain, ph.max_force, 0.0, tvel, &arm, &task, &ob, on_path_b);
        if (live && j == 0) obj_store(b.obj + (size_t)e * 13, b.obj_ext + (size_t)e * 4, ob);
    } else {
#pragma unroll 1
        for (int s = 0; s < ph.substeps; s++) g8_substep<T>(ph, bc, s_sub, max_sub, j, live, st, 0, 0.0, ph.vel_gain, ph.max_force, 0.0, tvel);
    }

    // ---- tail of the step: FK of the final pose on the lanes (exact sin / cos, as the one-thread kernels take them), TCP pose
    // and camera frame handed to the first lane, which writes the step data
    {
        double R[9], p[3], tp[3], tq[4], cam[12], sn, cs;
        sincos(st.q, &sn, &cs);
        g8_fk(bc, j, sn, cs, R, p);
        g8_tcp_and_camera(arm, R, p, tp, tq, cam);
#pragma unroll
        for (int i = 0; i < NB; i++) { q[i] = g8_get(st.q, i); qd[i] = g8_get(st.qd, i); }
        if (live && j == 0) {
            if (balance) env_epilogue_core<T, TASK>(arm, ph, task, b, e, q, qd, tp, tq, cam, ob_b, reward, done, autoreset);
            else {
                ObjState ob;
                env_epilogue_core<T, TASK>(arm, ph, task, b, e, q, qd, tp, tq, cam, ob, reward, done, autoreset);
            }
        }
    }
}

// test hook: nsteps x g8_substep with velocity motors from given joint states (tg_test_substep_g8)
template <class T>
__global__ void __launch_bounds__(128)
test_substep_g8_kernel(const __grid_constant__ TgArm arm, const __grid_constant__ TgPhysics ph, int n, int nsteps, double* q_io, double* qd_io,
                       const double* target_vel)
{
    constexpr int NB = T::NB;
    __shared__ double s_sub[TG_MAXSUB * G8_SUBW];
    g8_fill_sub_table(arm, s_sub);
    __syncthreads();
    const int j = threadIdx.x & (G8 - 1);
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const bool live = e < n, mine = live && j < NB;
    G8Body bc;
    g8_load_body<T>(arm, j, bc);
    const int max_sub = g8_max_sub(bc);
    G8Lane st;
    st.q = mine ? q_io[e * NB + j] : 0.0;
    st.qd = mine ? qd_io[e * NB + j] : 0.0;
    const double tv = mine ? target_vel[e * NB + j] : 0.0;
    sincos(st.q, &st.s, &st.c);
#pragma unroll 1
    for (int s = 0; s < nsteps; s++) g8_substep<T>(ph, bc, s_sub, max_sub, j, live, st, 0, 0.0, ph.vel_gain, ph.max_force, 0.0, tv);
    if (mine) { q_io[e * NB + j] = st.q; qd_io[e * NB + j] = st.qd; }
}
