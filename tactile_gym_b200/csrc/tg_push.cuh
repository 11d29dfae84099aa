// tg_push.cuh - object_push: one Robot.step_sim() with the cube in the world (motor rows + contact rows), the
// trajectory of goals, reward / termination and the extended feature.
//
//   substep_push      <- pb.stepSimulation() (robots/arms/robot.py:141) for the scene of ObjectPushEnv
//                        (rl_envs/nonprehensile_manipulation/object_push/object_push_env.py:204-229 cube dynamics,
//                        sensors/tactile_sensor.py:314-332 tip contact stiffness / damping / friction)
//   push_trajectory   <- update_trajectory_simplex / _straight + np.gradient (object_push_env.py:255-320)
//   push_step_data    <- get_step_data / dense_reward / sparse_reward / termination (:456-569)
//   push_features     <- get_extended_feature_array (:611-629)
//
// The contact model is the one oracle/tg_oracle.c:or_step_sim_push states (see oracle/tg_oracle.h for what is restated
// from bullet and what is simplified): memoryless manifolds rebuilt every substep - cube vertices against the table
// plane, tip-hull vertices inside the cube reduced to <= 4 points - one normal row (impulse >= 0, contact
// stiffness / damping as per-point erp / cfm) and two friction rows per point inside the cone mu * normal impulse,
// solved after the motor rows by the same projected Gauss-Seidel sweep, same residual exit.
//
// One thread per env.  The rows of a substep (<= 24 x (arm part NB + cube part 6)) live in per-thread local memory
// (~4 KB, L1-resident at the env counts of the configs); the hull scan reads the same vertex in every lane (one
// broadcast transaction per vertex per warp).
#pragma once
#include "tg_dyn.cuh"
#include "tg_surface.cuh"

#define PUSH_MAXC 8
#define PUSH_NTRAJ TG_PUSH_NTRAJ
#define PUSH_BLOCK 56 // envs per block: 56 x 492 slots x 8 B = 220 KB of the 227 KB a block may have, one block per SM
// The PGS sweep count differs from env to env (median ~55, 2 % of the substeps run into the cap of 150) and a warp sweeps
// until its slowest lane is done, so the 56 envs of a block are spread over 7 warps of PUSH_LANES = 8 active lanes:
// the expected slowest-of-8 is much shorter than the slowest-of-28, and 7 warps keep all four schedulers of the SM busy
// where 2 left half of them idle (ncu, 2 warps: issue slots 24 % busy, 48 % of the stall samples fixed-latency waits).
#define PUSH_LANES 8
#define PUSH_THREADS (32 * (PUSH_BLOCK / PUSH_LANES))

TGD long long push_qkey(double x) { return __double2ll_rn(x * 1e9); } // comparisons on a 1 nm grid: ties break by index

// signed distance of a cube-local point to the cube surface (negative inside) and the face it belongs to
TGD double cube_sd(const double* half, const double* l, int& axis, int& sign)
{
    double best = -1e300;
    int a = 0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double d = fabs(l[c]) - half[c];
        if (d > best) { best = d; a = c; }
    }
    axis = a;
    sign = (a == 0 ? l[0] : (a == 1 ? l[1] : l[2])) >= 0 ? 1 : -1;
    return best;
}

TGD void plane_space1(const double* n, double* p, double* q) // [EXT] btPlaneSpace1
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

struct PushContact {
    double n[3], pa[3], pb[3], dist;
    int on_arm;
};

// is joint j an ancestor-or-self of the body that carries the tip?
template <class T>
TGD bool tip_ancestor(int tcp_body, int j)
{
    bool anc = false;
    int a = tcp_body;
#pragma unroll
    for (int s = 0; s < T::NB; s++) {
        if (a == j) anc = true;
        if (a >= 0) {
            int pa = -1;
#pragma unroll
            for (int b = 0; b < T::NB; b++) if (a == b) pa = T::parent(b);
            a = pa;
        }
    }
    return anc;
}

// ---- narrow phase on the poses at the start of the substep ---------------------------------------------------------
// An env is stepped by ONE lane (its owner, lanes 0..PUSH_LANES-1 of a warp), but the scan of the tip hull's ~600
// vertices against the cube is shared with the warp's otherwise idle lanes: env slot s is served by the PUSH_PARTS lanes
// s, s + 8, s + 16, s + 24, each testing every fourth vertex; the partial results are merged with shuffles.
#define PUSH_PARTS (32 / PUSH_LANES)

struct HullScan {
    long long deep_key, ext_key[3][2]; // 1 nm keys of the deepest signed distance / of the extreme cube-local coordinates
    int deep, ext[3][2], ncand;        // and the hull vertices that attain them (lowest index on ties)
};

// all 32 lanes of the warp call this together.  M, t: cube-local coordinates of hull vertex v are M v + t (valid in the
// owner lane only; broadcast here).
__device__ __noinline__ void push_scan(const double* __restrict__ hull, int n_hull, const double* Mo, const double* to, const double* half,
                                       double slop, HullScan& h)
{
    const int lane = threadIdx.x & 31, slot = lane & (PUSH_LANES - 1), part = lane / PUSH_LANES;
    double M[9], t[3];
#pragma unroll
    for (int i = 0; i < 9; i++) M[i] = __shfl_sync(0xffffffffu, Mo[i], slot);
#pragma unroll
    for (int i = 0; i < 3; i++) t[i] = __shfl_sync(0xffffffffu, to[i], slot);
    h.deep_key = 0x7fffffffffffffffLL; h.deep = 0x7fffffff; h.ncand = 0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        h.ext_key[c][0] = 0x7fffffffffffffffLL; h.ext_key[c][1] = -0x7fffffffffffffffLL - 1;
        h.ext[c][0] = 0x7fffffff; h.ext[c][1] = 0x7fffffff;
    }
#pragma unroll 4
    for (int i = part; i < n_hull; i += PUSH_PARTS) {
        const double v[3] = {__ldg(hull + 3 * i), __ldg(hull + 3 * i + 1), __ldg(hull + 3 * i + 2)};
        double l[3];
        m3mulv(l, M, v);
        l[0] += t[0]; l[1] += t[1]; l[2] += t[2];
        int ax, sg;
        const double sd = cube_sd(half, l, ax, sg);
        if (sd > slop) continue;
        const long long key = push_qkey(sd);
        if (key < h.deep_key) { h.deep_key = key; h.deep = i; }
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const long long lk = push_qkey(l[c]);
            if (lk < h.ext_key[c][0]) { h.ext_key[c][0] = lk; h.ext[c][0] = i; }
            if (lk > h.ext_key[c][1]) { h.ext_key[c][1] = lk; h.ext[c][1] = i; }
        }
        h.ncand++;
    }
    // merge the parts (xor over the lane bits above the slot bits); ties go to the lower vertex index, as in a
    // sequential scan in index order
#pragma unroll
    for (int m = PUSH_LANES; m < 32; m <<= 1) {
        {
            const long long ok = __shfl_xor_sync(0xffffffffu, h.deep_key, m);
            const int oi = __shfl_xor_sync(0xffffffffu, h.deep, m);
            if (ok < h.deep_key || (ok == h.deep_key && oi < h.deep)) { h.deep_key = ok; h.deep = oi; }
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {
            long long ok = __shfl_xor_sync(0xffffffffu, h.ext_key[c][0], m);
            int oi = __shfl_xor_sync(0xffffffffu, h.ext[c][0], m);
            if (ok < h.ext_key[c][0] || (ok == h.ext_key[c][0] && oi < h.ext[c][0])) { h.ext_key[c][0] = ok; h.ext[c][0] = oi; }
            ok = __shfl_xor_sync(0xffffffffu, h.ext_key[c][1], m);
            oi = __shfl_xor_sync(0xffffffffu, h.ext[c][1], m);
            if (ok > h.ext_key[c][1] || (ok == h.ext_key[c][1] && oi < h.ext[c][1])) { h.ext_key[c][1] = ok; h.ext[c][1] = oi; }
        }
        h.ncand += __shfl_xor_sync(0xffffffffu, h.ncand, m);
    }
}

// cube <-> table: cube vertices at or below the table top (owner lane)
TGD int push_table_contacts(const TgTask& task, const ObjState& o, const double* Rb, PushContact* C)
{
    int nc = 0;
#pragma unroll 1
    for (int v = 0; v < 8 && nc < 4; v++) {
        const double l[3] = {(v & 1) ? task.push_half[0] : -task.push_half[0], (v & 2) ? task.push_half[1] : -task.push_half[1],
                             (v & 4) ? task.push_half[2] : -task.push_half[2]};
        double w[3];
        m3mulv(w, Rb, l);
        w[0] += o.pos[0]; w[1] += o.pos[1]; w[2] += o.pos[2];
        const double dist = w[2] - task.push_table_z;
        if (dist > task.push_slop) continue;
        PushContact& c = C[nc++];
        c.on_arm = 0;
        c.n[0] = 0; c.n[1] = 0; c.n[2] = 1;
#pragma unroll
        for (int q = 0; q < 3; q++) { c.pb[q] = w[q]; c.pa[q] = w[q]; }
        c.dist = dist;
    }
    return nc;
}

// tip core hull <-> cube: reduce the penetrating vertices to <= 4 contacts (owner lane): the deepest, the two extremes
// along the first tangent axis of its face, the extreme along the second that is farther from the deepest
__device__ __noinline__ int push_tip_contacts(const TgTask& task, const double* __restrict__ hull, const HullScan& h, const double* M, const double* t,
                                              const double* Rb, const double* Rt, const double* pt, PushContact* C, int nc)
{
    if (h.ncand <= 0) return nc;
    int A, sg;
    double l[3];
    {
        const double v[3] = {__ldg(hull + 3 * h.deep), __ldg(hull + 3 * h.deep + 1), __ldg(hull + 3 * h.deep + 2)};
        m3mulv(l, M, v);
        l[0] += t[0]; l[1] += t[1]; l[2] += t[2];
        cube_sd(task.push_half, l, A, sg);
    }
    const int U = (A + 1) % 3, V = (A + 2) % 3;
    const long long dv = push_qkey(V == 0 ? l[0] : (V == 1 ? l[1] : l[2]));
    long long kV0 = 0, kV1 = 0;
    int eU0 = 0, eU1 = 0, eV0 = 0, eV1 = 0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        if (c == U) { eU0 = h.ext[c][0]; eU1 = h.ext[c][1]; }
        if (c == V) { kV0 = h.ext_key[c][0]; kV1 = h.ext_key[c][1]; eV0 = h.ext[c][0]; eV1 = h.ext[c][1]; }
    }
    const long long a0 = llabs(kV0 - dv), a1 = llabs(kV1 - dv);
    const int sel[4] = {h.deep, eU0, eU1, a0 >= a1 ? eV0 : eV1};
#pragma unroll 1
    for (int j = 0; j < 4; j++) {
        bool dup = false;
        for (int i = 0; i < j; i++) if (sel[i] == sel[j]) dup = true;
        if (dup) continue;
        const double v[3] = {__ldg(hull + 3 * sel[j]), __ldg(hull + 3 * sel[j] + 1), __ldg(hull + 3 * sel[j] + 2)};
        m3mulv(l, M, v);
        l[0] += t[0]; l[1] += t[1]; l[2] += t[2];
        int ax;
        const double sd = cube_sd(task.push_half, l, ax, sg);
        PushContact& c = C[nc++];
        c.on_arm = 1;
        // world normal = sign * column `ax` of Rb
        c.n[0] = sg * (ax == 0 ? Rb[0] : (ax == 1 ? Rb[1] : Rb[2]));
        c.n[1] = sg * (ax == 0 ? Rb[3] : (ax == 1 ? Rb[4] : Rb[5]));
        c.n[2] = sg * (ax == 0 ? Rb[6] : (ax == 1 ? Rb[7] : Rb[8]));
        double w[3];
        m3mulv(w, Rt, v);
#pragma unroll
        for (int q = 0; q < 3; q++) { c.pa[q] = pt[q] + w[q]; c.pb[q] = c.pa[q] - sd * c.n[q]; }
        c.dist = sd;
    }
    return nc;
}

// object_roll narrow phase (owner lane): sphere <-> table = one point below the centre; sphere <-> the cap of the tip
// core's cylinder that faces it, while the centre projects inside the cap (rim / side contacts not generated)
// (oracle/tg_oracle.c:push_contacts, shape 1)
TGD int roll_contacts(const TgTask& task, const ObjState& o, double radius, const double* Rt, const double* pt, PushContact* C)
{
    int nc = 0;
    const double dist = o.pos[2] - radius - task.push_table_z;
    if (dist <= task.push_slop) {
        PushContact& c = C[nc++];
        c.on_arm = 0;
        c.n[0] = 0; c.n[1] = 0; c.n[2] = 1;
        c.pb[0] = o.pos[0]; c.pb[1] = o.pos[1]; c.pb[2] = o.pos[2] - radius;
#pragma unroll
        for (int q = 0; q < 3; q++) c.pa[q] = c.pb[q];
        c.dist = dist;
    }
    double cc[3], ax[3], t[3];
    m3mulv(t, Rt, task.roll_cyl_pos);
    m3mulv(ax, Rt, task.roll_cyl_axis);
#pragma unroll
    for (int q = 0; q < 3; q++) cc[q] = pt[q] + t[q];
    const double d[3] = {o.pos[0] - cc[0], o.pos[1] - cc[1], o.pos[2] - cc[2]};
    double h = v3dot(d, ax);
    if (h < 0) { ax[0] = -ax[0]; ax[1] = -ax[1]; ax[2] = -ax[2]; h = -h; } // the cap that faces the sphere
    const double lat[3] = {d[0] - h * ax[0], d[1] - h * ax[1], d[2] - h * ax[2]};
    const double sd = h - task.roll_cyl_half_len - radius;
    if (sd <= task.push_slop && v3dot(lat, lat) <= task.roll_cyl_radius * task.roll_cyl_radius) {
        PushContact& c = C[nc++];
        c.on_arm = 1;
#pragma unroll
        for (int q = 0; q < 3; q++) { c.n[q] = -ax[q]; c.pb[q] = o.pos[q] - radius * ax[q]; c.pa[q] = c.pb[q] - sd * ax[q]; }
        c.dist = sd;
    }
    return nc;
}

// ---- shared-memory staging of one env's constraint rows -------------------------------------------------------------
// A substep's rows are read ~55 times (PGS sweeps) each: they live in shared memory, one COLUMN per env
// (slot s of thread t at sm[s * stride + t]: consecutive threads touch consecutive doubles, no bank conflicts).
// ncu of the first version, which kept them in per-thread local memory: 74 % of the stall samples on local loads, L1 hit
// rate 18 %, 15.6 GB of DRAM reads per launch (profiles/r01_push_step_v0.md).
//   motors      : A = M^-1 upper triangle NB (NB + 1) / 2, then rhs / dinv / applied per joint
//   table rows  : 3 per contact x 4 contacts: ja(3) ua(3) rhs dinv diagc applied  (normal +z, fixed friction basis)
//   tip rows    : 3 per contact x 4 contacts: jl(3) ja(3) ua(3) jr(NA) ur(NB) rhs dinv diagc applied
// with NA = joints between the base and the tip (the MG400's three slaved joints are not among them).
template <class T>
struct PushLayout {
    static constexpr int NB = T::NB;
    static constexpr int NA = NB == 8 ? 5 : NB;   // TopoMG400: chain 0-1-2-3-4 carries the tip; TopoChain6: all six
    static constexpr int TRI = NB * (NB + 1) / 2;
    // every section starts on an even slot and rows have even lengths: slots are stored in PAIRS (slot 2k and 2k+1 of
    // an env are adjacent), so that the loads of a row pair up into 16-byte LDS (half as many load instructions)
    static constexpr int MOT = (TRI + 1) & ~1;               // + 3 * NB
    static constexpr int TAB = MOT + ((3 * NB + 1) & ~1);    // 12 rows x 10
    static constexpr int TAB_ROW = 10;
    static constexpr int TIP = TAB + 12 * TAB_ROW;
    static constexpr int TIP_ROW = (13 + NA + NB + 1) & ~1;
    static constexpr int SLOTS = TIP + 12 * TIP_ROW;
    __host__ __device__ static constexpr int tri(int i, int j) { return i <= j ? i * NB - i * (i - 1) / 2 + (j - i) : j * NB - j * (j - 1) / 2 + (i - j); }
};


// Robot.step_sim() with the cube in the world.  `col` = this env's column of the block's row store.  All 32 lanes of the
// warp call it together; only the owner lanes (`owner`) carry an env, the others just lend a hand in the hull scan.
// Returns the number of PGS sweeps.
template <class T>
__device__ __noinline__ int substep_push(const TgArm& arm, const TgPhysics& ph, const TgTask& task, const double* __restrict__ hull, int n_hull,
                                         double* q, double* qd, double (&sc)[T::NB][2], const Motors<T::NB>& mot, ObjState& o, const int col, const bool owner)
{
    extern __shared__ __align__(16) double push_rows[]; // [PushLayout<T>::SLOTS / 2][PUSH_BLOCK][2]; indexed directly so that the accesses are LDS / STS
    using LY = PushLayout<T>;
    constexpr int NB = T::NB, NA = LY::NA;
    constexpr double EPS = 2.2204460492503131e-16;
#define SM(slot) push_rows[(((slot) >> 1) * PUSH_BLOCK + col) * 2 + ((slot) & 1)]
// same for an EVEN run-time base b and a compile-time offset k: the pair index and the parity are then constants
#define SMB(b, k) push_rows[((((b) >> 1) + ((k) >> 1)) * PUSH_BLOCK + col) * 2 + ((k) & 1)]
    PushContact C[PUSH_MAXC];
    int nc = 0, ntab = 0;
    const bool sphere = task.push_shape == 1; // object_roll: the episode's marble radius rides in o.ext_pos[0]
    double Rb[9], Iinv[3], Rt[9], pt[3], Mc[9], tc[3], ipm[3];
#pragma unroll
    for (int c = 0; c < 3; c++) ipm[c] = sphere ? 0.4 * o.ext_pos[0] * o.ext_pos[0] : task.push_inertia_per_mass[c]; // [EXT] btSphereShape: 2/5 m r^2
    Kin<NB> k;
    const double mass = o.mass, minv = 1.0 / mass;
    const double lim_m = mot.max_force * ph.dt;
#pragma unroll
    for (int i = 0; i < 9; i++) { Mc[i] = 0.0; Rb[i] = 0.0; }
    tc[0] = tc[1] = tc[2] = 0.0;
    if (owner) {
        double A[NB][NB];
        robot_pre<T>(arm, ph, q, qd, sc, A);
        // motor rows, as in substep(): J = e_i, response column A[:, i]
#pragma unroll
        for (int i = 0; i < NB; i++) {
#pragma unroll
            for (int j = i; j < NB; j++) SM(LY::tri(i, j)) = A[i][j];
            const double denom = A[i][i];
            const double mdinv = denom > EPS ? 1.0 / denom : 0.0;
            const double v = qd[i];
            const double pos_stab = mot.mode == 1 ? mot.kp * ((mot.target_pos[i] - q[i]) / ph.dt) : 0.0;
            const double rhs_v = pos_stab + v + mot.kd * (mot.target_vel[i] - v);
            SM(LY::MOT + 3 * i) = (rhs_v - v) * mdinv;
            SM(LY::MOT + 3 * i + 1) = mdinv;
            SM(LY::MOT + 3 * i + 2) = 0.0;
        }
        fk_sc<T>(arm, sc, k);
        mat_from_quat(o.quat, Rb);
        if (!sphere) nc = push_table_contacts(task, o, Rb, C);
        // cube-local coordinates of a hull vertex v (tip body frame) are Mc v + tc
#pragma unroll
        for (int b2 = 0; b2 < NB; b2++)
            if (arm.tcp_body == b2) {
#pragma unroll
                for (int i = 0; i < 9; i++) Rt[i] = k.R[b2][i];
#pragma unroll
                for (int i = 0; i < 3; i++) pt[i] = k.p[b2][i];
            }
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) Mc[3 * i + j] = Rb[i] * Rt[j] + Rb[3 + i] * Rt[3 + j] + Rb[6 + i] * Rt[6 + j]; // Rb^T Rt
        const double d[3] = {pt[0] - o.pos[0], pt[1] - o.pos[1], pt[2] - o.pos[2]};
        m3tmulv(tc, Rb, d);
    }
    HullScan hs;
    if (!sphere) push_scan(hull, n_hull, Mc, tc, task.push_half, task.push_slop, hs); // the whole warp
    if (!owner) return 0;
    if (sphere) nc = roll_contacts(task, o, o.ext_pos[0], Rt, pt, C);
    else nc = push_tip_contacts(task, hull, hs, Mc, tc, Rb, Rt, pt, C, nc);
    {
        // cube: unconstrained velocity update about its COM (gravity, [EXT] multibody base damping, gyroscopic term)
#pragma unroll
        for (int c = 0; c < 3; c++) Iinv[c] = 1.0 / (ipm[c] * mass);
        {
            double wl[3], Iw[3], gy[3], al[3], aw[3];
            m3tmulv(wl, Rb, o.omg);
#pragma unroll
            for (int c = 0; c < 3; c++) Iw[c] = ipm[c] * mass * wl[c];
            v3cross(gy, wl, Iw);
            const double ka = task.push_ang_damping * (1.0 + sqrt(v3dot(o.omg, o.omg))), kl = task.push_lin_damping * (1.0 + sqrt(v3dot(o.vel, o.vel)));
#pragma unroll
            for (int c = 0; c < 3; c++) al[c] = (-Iw[c] * ka - gy[c]) * Iinv[c];
            m3mulv(aw, Rb, al);
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const double f = ph.gravity[c] * mass - mass * o.vel[c] * kl;
                o.vel[c] += ph.dt * f / mass;
                o.omg[c] += ph.dt * aw[c];
            }
        }

        // contact rows: row 3c normal, 3c+1 / 3c+2 friction; table contacts come first
        const double dtk = fmax(ph.dt * task.push_tip_k + task.push_tip_d, EPS);
#pragma unroll 1
        for (int c = 0; c < nc; c++) {
            const PushContact& ct = C[c];
            const bool on_arm = ct.on_arm != 0;
            if (!on_arm) ntab = c + 1;
            const double erp = on_arm ? (ph.dt * task.push_tip_k) / dtk : task.push_erp;
            const double cfm = on_arm ? (1.0 / dtk) / ph.dt : 0.0;
            double lin[NA][3]; // velocity of the arm's contact point per unit rate of the joints that carry the tip
            double va[3] = {0, 0, 0}, vb[3], rb[3], vrel[3];
            if (on_arm) {
#pragma unroll
                for (int j = 0; j < NA; j++) {
                    const double r[3] = {ct.pa[0] - k.p[j][0], ct.pa[1] - k.p[j][1], ct.pa[2] - k.p[j][2]};
                    v3cross(lin[j], k.a[j], r);
#pragma unroll
                    for (int x = 0; x < 3; x++) va[x] += qd[j] * lin[j][x];
                }
            }
            {
                double t[3];
#pragma unroll
                for (int x = 0; x < 3; x++) rb[x] = ct.pb[x] - o.pos[x];
                v3cross(t, o.omg, rb);
#pragma unroll
                for (int x = 0; x < 3; x++) { vb[x] = o.vel[x] + t[x]; vrel[x] = on_arm ? va[x] - vb[x] : vb[x]; }
            }
            // friction basis: fixed per normal (see oracle/tg_oracle.c: the velocity-aligned first direction is not restated)
            double dir[3][3];
#pragma unroll
            for (int x = 0; x < 3; x++) dir[0][x] = ct.n[x];
            plane_space1(ct.n, dir[1], dir[2]);
            const double scb = on_arm ? -1.0 : 1.0; // the cube is the second body of a tip contact
#pragma unroll
            for (int qq = 0; qq < 3; qq++) {
                const int base = on_arm ? LY::TIP + (3 * (c - ntab) + qq) * LY::TIP_ROW : LY::TAB + (3 * c + qq) * LY::TAB_ROW;
                const int o_ja = on_arm ? 3 : 0, o_sc = on_arm ? 9 + NA + NB : 6;
                double den = 0, ja[3], ua[3], a[3], bq[3];
                if (on_arm) {
                    double jr[NA];
#pragma unroll
                    for (int j = 0; j < NA; j++) { jr[j] = v3dot(dir[qq], lin[j]); SM(base + 9 + j) = jr[j]; }
#pragma unroll
                    for (int j = 0; j < NB; j++) {
                        double u = 0;
#pragma unroll
                        for (int e = 0; e < NA; e++) u += SM(LY::tri(j, e)) * jr[e];
                        SM(base + 9 + NA + j) = u;
                        if (j < NA) den += jr[j < NA ? j : 0] * u;
                    }
#pragma unroll
                    for (int x = 0; x < 3; x++) SM(base + x) = scb * dir[qq][x];
                }
                v3cross(ja, rb, dir[qq]);
#pragma unroll
                for (int x = 0; x < 3; x++) ja[x] *= scb;
                m3tmulv(a, Rb, ja);
#pragma unroll
                for (int x = 0; x < 3; x++) bq[x] = a[x] * Iinv[x];
                m3mulv(ua, Rb, bq);
#pragma unroll
                for (int x = 0; x < 3; x++) { SM(base + o_ja + x) = ja[x]; SM(base + o_ja + 3 + x) = ua[x]; }
                den += v3dot(dir[qq], dir[qq]) * minv + v3dot(ja, ua);
                const double rel = v3dot(dir[qq], vrel);
                double dinv, rhs, diagc;
                if (qq == 0) {
                    dinv = 1.0 / (den + cfm);
                    const double positional = ct.dist > 0 ? 0.0 : -ct.dist * erp / ph.dt;
                    const double velerr = -rel - (ct.dist > 0 ? ct.dist / ph.dt : 0.0);
                    rhs = (positional + velerr) * dinv;
                    diagc = den + cfm;
                } else {
                    dinv = den > EPS ? 1.0 / den : 0.0;
                    rhs = -rel * dinv;
                    diagc = den;
                }
                SM(base + o_sc) = rhs; SM(base + o_sc + 1) = dinv; SM(base + o_sc + 2) = diagc; SM(base + o_sc + 3) = 0.0;
            }
        }
    }
    const int ntip = nc - ntab;
    const double cfm_tip = (1.0 / fmax(ph.dt * task.push_tip_k + task.push_tip_d, EPS)) / ph.dt;

    // ---- projected Gauss-Seidel ------------------------------------------------------------------------------------
    // The sweep is one long dependent chain (each row reads the velocities the previous row wrote), so a row update is
    // written for latency: all of its operands are loaded first, dot products are summed as trees (depth 5 instead of
    // 11-14 dependent FMAs), the cone clamp uses rsqrt instead of sqrt + divide.
    double dv[NB], dvl[3] = {0, 0, 0}, dva[3] = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < NB; i++) dv[i] = 0;
    auto row_m = [&](int r, double& resid) {
        double Ar[NB];
#pragma unroll
        for (int i = 0; i < NB; i++) Ar[i] = SM(LY::tri(r, i));
        const double mrhs = SM(LY::MOT + 3 * r), mdinv = SM(LY::MOT + 3 * r + 1), mapp = SM(LY::MOT + 3 * r + 2);
        double delta = mrhs - dv[r] * mdinv;
        const double sum = mapp + delta;
        const bool lo = sum < -lim_m, hi = sum > lim_m;
        delta = lo ? (-lim_m - mapp) : (hi ? (lim_m - mapp) : delta);
        SM(LY::MOT + 3 * r + 2) = lo ? -lim_m : (hi ? lim_m : sum);
#pragma unroll
        for (int i = 0; i < NB; i++) dv[i] += Ar[i] * delta;
        const double dvel = delta * Ar[r];
        resid = fmax(resid, dvel * dvel);
    };
    // one row's operands, in registers
    struct TabRow { double ja[3], ua[3], rhs, dinv, diagc, app; };
    struct TipRow { double jl[3], ja[3], ua[3], jr[NA], ur[NB], rhs, dinv, diagc, app; };
    auto tab_load = [&](int base, TabRow& R) {
#pragma unroll
        for (int x = 0; x < 3; x++) { R.ja[x] = SMB(base, x); R.ua[x] = SMB(base, 3 + x); }
        R.rhs = SMB(base, 6); R.dinv = SMB(base, 7); R.diagc = SMB(base, 8); R.app = SMB(base, 9);
    };
    auto tip_load = [&](int base, TipRow& R) {
#pragma unroll
        for (int x = 0; x < 3; x++) { R.jl[x] = SMB(base, x); R.ja[x] = SMB(base, 3 + x); R.ua[x] = SMB(base, 6 + x); }
#pragma unroll
        for (int j = 0; j < NA; j++) R.jr[j] = SMB(base, 9 + j);
#pragma unroll
        for (int j = 0; j < NB; j++) R.ur[j] = SMB(base, 9 + NA + j);
        R.rhs = SMB(base, 9 + NA + NB); R.dinv = SMB(base, 10 + NA + NB); R.diagc = SMB(base, 11 + NA + NB); R.app = SMB(base, 12 + NA + NB);
    };
    // table rows: directions are (0,0,1), (0,-1,0), (1,0,0) = btPlaneSpace1 of +z; the cube is the first body
    auto tab_dot = [&](const TabRow& R, int qq) {
        const double lin = qq == 0 ? dvl[2] : (qq == 1 ? -dvl[1] : dvl[0]);
        return (lin + R.ja[0] * dva[0]) + (R.ja[1] * dva[1] + R.ja[2] * dva[2]);
    };
    auto tab_apply = [&](const TabRow& R, int qq, double delta) {
        const double d = delta * minv;
        if (qq == 0) dvl[2] += d; else if (qq == 1) dvl[1] -= d; else dvl[0] += d;
#pragma unroll
        for (int x = 0; x < 3; x++) dva[x] += R.ua[x] * delta;
    };
    auto tip_dot = [&](const TipRow& R) {
        const double p0 = R.jl[0] * dvl[0] + R.jl[1] * dvl[1], p1 = R.jl[2] * dvl[2] + R.ja[0] * dva[0], p2 = R.ja[1] * dva[1] + R.ja[2] * dva[2];
        double pa[(NA + 1) / 2];
#pragma unroll
        for (int j = 0; j + 1 < NA; j += 2) pa[j / 2] = R.jr[j] * dv[j] + R.jr[j + 1] * dv[j + 1];
        if (NA & 1) pa[NA / 2] = R.jr[NA - 1] * dv[NA - 1];
        double sa = (pa[0] + pa[1]) + pa[2];                   // NA = 5 or 6: three partial sums
        return ((p0 + p1) + p2) + sa;
    };
    auto tip_apply = [&](const TipRow& R, double delta) {
        const double d = delta * minv;
#pragma unroll
        for (int x = 0; x < 3; x++) { dvl[x] += R.jl[x] * d; dva[x] += R.ua[x] * delta; }
#pragma unroll
        for (int j = 0; j < NB; j++) dv[j] += R.ur[j] * delta;
    };
    int it = 0;
#pragma unroll 1
    for (; it < ph.solver_iters; it++) {
        double resid = 0;
        if (lim_m != 0.0) {
            if (it & 1) {
#pragma unroll
                for (int r = 0; r < NB; r++) row_m(r, resid);
            } else {
#pragma unroll
                for (int r = NB - 1; r >= 0; r--) row_m(r, resid);
            }
        }
        // normal rows: impulse >= 0
#pragma unroll 1
        for (int c = 0; c < ntab; c++) {
            const int base = LY::TAB + 3 * c * LY::TAB_ROW;
            TabRow R;
            tab_load(base, R);
            double delta = R.rhs - tab_dot(R, 0) * R.dinv; // cfm = 0 on the table
            const double sum = R.app + delta;
            const bool lo = sum < 0.0;
            delta = lo ? -R.app : delta;
            SMB(base, 9) = lo ? 0.0 : sum;
            tab_apply(R, 0, delta);
            const double dvel = delta * R.diagc;
            resid = fmax(resid, dvel * dvel);
        }
#pragma unroll 1
        for (int c = 0; c < ntip; c++) {
            const int base = LY::TIP + 3 * c * LY::TIP_ROW;
            TipRow R;
            tip_load(base, R);
            const double pre = R.rhs - R.app * (cfm_tip * R.dinv);
            double delta = pre - tip_dot(R) * R.dinv;
            const double sum = R.app + delta;
            const bool lo = sum < 0.0;
            delta = lo ? -R.app : delta;
            SMB(base, 12 + NA + NB) = lo ? 0.0 : sum;
            tip_apply(R, delta);
            const double dvel = delta * R.diagc;
            resid = fmax(resid, dvel * dvel);
        }
        // friction pairs inside the cone mu * normal impulse
#pragma unroll 1
        for (int c = 0; c < ntab; c++) {
            const int b0 = LY::TAB + 3 * c * LY::TAB_ROW, b1 = b0 + LY::TAB_ROW, b2 = b1 + LY::TAB_ROW;
            const double napp = SMB(b0, 9);
            if (!(napp > 0.0)) continue;
            TabRow R1, R2;
            tab_load(b1, R1); tab_load(b2, R2);
            const double lim = task.push_mu_table * napp;
            double s1 = R1.app + (R1.rhs - tab_dot(R1, 1) * R1.dinv);
            double s2 = R2.app + (R2.rhs - tab_dot(R2, 2) * R2.dinv);
            const double nrm2 = s1 * s1 + s2 * s2;
            if (nrm2 > lim * lim) { const double scl = lim * rsqrt(nrm2); s1 *= scl; s2 *= scl; }
            const double d1 = s1 - R1.app, d2 = s2 - R2.app;
            SMB(b1, 9) = s1; SMB(b2, 9) = s2;
            tab_apply(R1, 1, d1);
            tab_apply(R2, 2, d2);
            const double e1 = d1 * R1.diagc, e2 = d2 * R2.diagc;
            resid = fmax(resid, fmax(e1 * e1, e2 * e2));
        }
#pragma unroll 1
        for (int c = 0; c < ntip; c++) {
            const int b0 = LY::TIP + 3 * c * LY::TIP_ROW, b1 = b0 + LY::TIP_ROW, b2 = b1 + LY::TIP_ROW, so = 9 + NA + NB;
            const double napp = SMB(b0, so + 3);
            if (!(napp > 0.0)) continue;
            TipRow R1, R2;
            tip_load(b1, R1); tip_load(b2, R2);
            const double lim = task.push_mu_tip * napp;
            double s1 = R1.app + (R1.rhs - tip_dot(R1) * R1.dinv);
            double s2 = R2.app + (R2.rhs - tip_dot(R2) * R2.dinv);
            const double nrm2 = s1 * s1 + s2 * s2;
            if (nrm2 > lim * lim) { const double scl = lim * rsqrt(nrm2); s1 *= scl; s2 *= scl; }
            const double d1 = s1 - R1.app, d2 = s2 - R2.app;
            SMB(b1, so + 3) = s1; SMB(b2, so + 3) = s2;
            tip_apply(R1, d1);
            tip_apply(R2, d2);
            const double e1 = d1 * R1.diagc, e2 = d2 * R2.diagc;
            resid = fmax(resid, fmax(e1 * e1, e2 * e2));
        }
        if (resid <= ph.solver_residual_threshold) { it++; break; }
    }
#undef SMB
#undef SM
#pragma unroll
    for (int i = 0; i < NB; i++) {
        qd[i] += dv[i];
        const double d = ph.dt * qd[i];
        q[i] += d;
        sc_advance(sc[i], q[i], d);
    }
#pragma unroll
    for (int c = 0; c < 3; c++) { o.vel[c] += dvl[c]; o.omg[c] += dva[c]; o.pos[c] += ph.dt * o.vel[c]; }
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
    return it;
}

// ---------------------------------------------------------------- trajectory of goals
// update_trajectory (object_push_env.py:255-320): traj[i] = y_i (work frame), traj[PUSH_NTRAJ + i] = Rz_i = np.gradient(y, spacing)
__device__ __noinline__ void push_trajectory(const TgTask& task, double third, double* traj)
{
    if (task.push_traj_straight) {
        double sn, cs;
        sincos(third, &sn, &cs);
#pragma unroll 1
        for (int i = 0; i < PUSH_NTRAJ; i++) traj[i] = ((double)i * task.push_traj_spacing) * sn;
    } else {
        unsigned char perm[256];
        os_perm((long long)third, perm);
        double first = 0;
#pragma unroll 1
        for (int i = 0; i < PUSH_NTRAJ; i++) {
            const double noise = os_noise2(perm, (double)i * 0.1, 1.0) * task.push_traj_perturb;
            if (i == 0) first = -noise;
            traj[i] = first + noise;
        }
    }
    const double h = task.push_traj_spacing;
#pragma unroll 1
    for (int i = 0; i < PUSH_NTRAJ; i++) {
        double g;
        if (i == 0) g = (traj[1] - traj[0]) / h;
        else if (i == PUSH_NTRAJ - 1) g = (traj[i] - traj[i - 1]) / h;
        else g = (traj[i + 1] - traj[i - 1]) / (2.0 * h);
        traj[PUSH_NTRAJ + i] = g;
    }
}

// x of goal i in the work frame: straight trajectories advance along the trajectory direction (:312-320)
TGD double push_goal_x(const TgTask& task, double third, int i)
{
    if (task.push_traj_straight) return task.push_traj_offset + ((double)i * task.push_traj_spacing) * cos(third);
    return task.push_traj_offset + (double)i * task.push_traj_spacing;
}

// goal i in the world frame: workframe_to_worldframe (base_robot_arm.py:47-60), then getQuaternionFromEuler of the rpy
TGD void push_goal_world(const TgTask& task, double gx, double gy, double grz, double* pos, double* quat)
{
    double wq[4], gq[4], R[9], t[3], oq[4], rpy[3];
    const double lp[3] = {gx, gy, 0.0}, lr[3] = {0.0, 0.0, grz};
    quat_from_euler(task.workframe_rpy, wq);
    quat_from_euler(lr, gq);
    mat_from_quat(wq, R);
    m3mulv(t, R, lp);
    pos[0] = task.workframe_pos[0] + t[0]; pos[1] = task.workframe_pos[1] + t[1]; pos[2] = task.workframe_pos[2] + t[2];
    quat_mul(oq, wq, gq);
    euler_from_quat(oq, rpy);
    quat_from_euler(rpy, quat);
}

// get_step_data (:456-569): reward against the CURRENT goal, then termination() may advance the goal.
// goal index is updated in place; returns done.
TGD void push_step_data(const TgTask& task, const ObjState& o, const double* tcp_quat, const double* traj, double third, int& goal,
                        int steps, float* reward, unsigned char* done)
{
    const int gi = goal < PUSH_NTRAJ ? goal : PUSH_NTRAJ - 1;
    double gp[3], gq[4];
    push_goal_world(task, push_goal_x(task, third, gi), traj[gi], traj[PUSH_NTRAJ + gi], gp, gq);
    const double dx = o.pos[0] - gp[0], dy = o.pos[1] - gp[1], dz = o.pos[2] - gp[2];
    const double pos_dist = sqrt(dx * dx + dy * dy + dz * dz);
    if (task.push_sparse_reward) *reward = pos_dist < task.push_term_dist ? 1.0f : 0.0f;
    else {
        const double ip = gq[0] * o.quat[0] + gq[1] * o.quat[1] + gq[2] * o.quat[2] + gq[3] * o.quat[3];
        const double orn_dist = acos(fmin(fmax(2.0 * (ip * ip) - 1.0, -1.0), 1.0));
        double Ro[9], Rt[9];
        mat_from_quat(o.quat, Ro);
        mat_from_quat(tcp_quat, Rt);
        const double ov[3] = {Ro[0], Ro[3], Ro[6]}, tv[3] = {Rt[0], Rt[3], Rt[6]};
        const double cos_dist = 1.0 - v3dot(ov, tv) / (sqrt(v3dot(ov, ov)) * sqrt(v3dot(tv, tv)));
        *reward = (float)(-((1.0 * pos_dist) + (1.0 * orn_dist) + (1.0 * cos_dist)));
    }
    bool d = false;
    if (pos_dist < task.push_term_dist) {
        goal++;
        if (goal >= PUSH_NTRAJ) d = true;
    }
    if (steps >= task.max_steps) d = true;
    *done = d ? 1 : 0;
}

// get_extended_feature_array (:611-629): TCP pose and current goal pose, both in the work frame
TGD void push_features(const TgTask& task, const double* tcp_pos, const double* tcp_quat, const double* traj, double third, int goal, float* out)
{
    double wp[3], wr[3];
    world_to_work(task, tcp_pos, tcp_quat, wp, wr);
    const int gi = goal < PUSH_NTRAJ ? (goal < 0 ? 0 : goal) : PUSH_NTRAJ - 1;
    out[0] = (float)wp[0]; out[1] = (float)wp[1]; out[2] = (float)wp[2];
    out[3] = (float)wr[0]; out[4] = (float)wr[1]; out[5] = (float)wr[2];
    out[6] = (float)push_goal_x(task, third, gi); out[7] = (float)traj[gi]; out[8] = 0.0f;
    out[9] = 0.0f; out[10] = 0.0f; out[11] = (float)traj[PUSH_NTRAJ + gi];
}

// ---------------------------------------------------------------- object_roll task level
// update_goal (object_roll_env.py:258-286): the goal is fixed in the TCP frame -> world; get_step_data / termination /
// rewards (:299-360): xy distance marble <-> goal, done below 1 mm
TGD void roll_step_data(const TgTask& task, const ObjState& o, const double* tcp_pos, const double* tcp_quat, double gx, double gy,
                        int steps, float* reward, unsigned char* done)
{
    double R[9], t[3];
    const double g[3] = {gx, gy, 0.0};
    mat_from_quat(tcp_quat, R);
    m3mulv(t, R, g);
    const double dx = o.pos[0] - (tcp_pos[0] + t[0]), dy = o.pos[1] - (tcp_pos[1] + t[1]);
    const double d = sqrt(dx * dx + dy * dy);
    const bool near_ = d < task.push_term_dist;
    *reward = task.push_sparse_reward ? (near_ ? 1.0f : 0.0f) : (float)(-(1.0 * d));
    *done = (near_ || steps >= task.max_steps) ? 1 : 0;
}
