// tg_env.cuh - env-level kernels (one thread per env): step (+ standby resets), reset.
//
//   step_kernel   <- BaseTactileEnv.step (rl_envs/base_tactile_env.py:166-185) up to, not including, the
//                    tactile render: action encode/scale, tcp_velocity_control, 24 x step_sim, step data.
//   reset_env     <- EdgeFollowEnv.reset (edge_follow_env.py:311-336) / Robot.reset (robot.py:114-125):
//                    rest pose, IK, blocking move.
// Both leave behind, per env, the camera frame and stimulus pose the raster kernel consumes.
//
// Reset pipeline.  A reset depends only on the env's next random draws, never on the episode that just ended
// (arm.reset() rewinds to the rest pose, robot.py:114-125).  So every env keeps a STANDBY start-of-episode
// state computed ahead of time.  When an env finishes, step_kernel swaps the standby in (a copy) instead of
// running IK + blocking move on the critical path.  The consumed slot is then rebuilt by the extra blocks of the
// following step launches, RESUMABLY: one launch does the draws + IK, every later one RESET_CHUNK iterations of
// the blocking move (a fraction of an env step's work), the partial state living in the sb_* buffers, so a launch
// that carries resets lasts no longer than one that does not.  Slots are handed between the step thread and the
// standby threads through sb_ready (EMPTY / PARTIAL / BUSY / READY, atomicCAS + fences); an env that finishes
// before its slot is READY claims it and completes it inline (counted in stall_count), so any episode length is
// exact.  Draws are consumed in episode order and the chunk schedule is fixed (the incremental sin/cos are
// re-synchronised every RESET_CHUNK iterations on every path), so results do not depend on who ran which chunk.
#pragma once
#include "tg_dyn.cuh"
#include "tg_rng.cuh"
#include "tg_surface.cuh"
#include "tg_push.cuh"

#define PUSH_TRAJ_SZ (2 * PUSH_NTRAJ + 1) // y_i, Rz_i (work frame) and the episode's third draw (seed / direction)

struct EnvBuffers {
    int n;
    int lanes;            // active lanes per warp in the step role
    int step_blocks;      // blocks [0, step_blocks) step envs, the rest recompute standbys
    int pipeline;         // 1: standby reset pipeline on
    int ik_chunk, reset_chunk, surf_chunk; // quantum sizes of the resumable reset (IK / blocking-move iterations, heightfield points per launch)
    double* q;            // [NB][N]
    double* qd;           // [NB][N]
    double* embed;        // [N]
    double* edge_ang;     // [N]
    int* steps;           // [N]
    int* reset_substeps;  // [N]
    int* reset_count;     // [N] draws consumed since tg_set_draws
    const double* draws;  // [N][rounds][n_draws] or null: a ring, the k-th reset of env e reads slot k % rounds
    int draw_rounds;
    const int* draw_avail; // [N] draws uploaded so far per env (reset_count[e] < draw_avail[e] or the draw is missing)
    uint32_t* mt;         // [N][624] device RNG: every env's MT19937 state (tg_set_rng_state), or null
    int* mt_pos;          // [N]
    int mt_active;        // 1: resets draw from the device RNG (TgTask.draw_kind), 0: from the host's ring
    int epoch;            // launch counter: tags the slots consumed in THIS launch (SB_CONSUMED + epoch)
    double* cam;          // [N][12] eye fwd up right
    double* stim;         // [N][12] R(9) t(3) of the stimulus frame
    double* tcp;          // [N][7] tcp world pos + quat (state export)
    const double* rest_q; // [NB]
    // free object (object_balance pole / object_push cube): [N][13] pos quat vel omg, [N][4] ext force point + pending
    // flag, [N] the episode's scalar (gravity_z for object_balance, the cube's mass for object_push)
    double *obj, *obj_ext, *grav;
    double *sb_obj, *sb_obj_ext, *sb_grav;
    // object_push: trajectory of goals [N][PUSH_TRAJ_SZ], current goal index [N], tip hull (tcp body frame), and the
    // caller's feature buffers [N][TG_PUSH_NFEAT] f32 (tg_bind_features; may be null)
    double *traj, *sb_traj;
    int *goal, *sb_goal;
    const double* hull;
    int n_hull;
    float *feat, *term_feat;
    // observation_mode "oracle": caller's buffers [N][TG_ORACLE_NOBS] f32 (tg_bind_oracle_obs; may be null)
    float *oracle, *term_oracle;
    // standby start-of-episode state
    double *sb_q, *sb_qd, *sb_embed, *sb_ang, *sb_cam, *sb_stim, *sb_tcp;
    int* sb_substeps;
    int* sb_ready;           // [N] slot state: SB_EMPTY / SB_READY / SB_PARTIAL / SB_BUSY
    int* sb_ik;              // [N] partial resets: IK iteration (ResetState::ik_it)
    double *sb_targ, *sb_cv, *sb_draw; // partial resets: IK target joints [NB][N], blocking-move step [N], draws [N][TG_MAXDRAW]
    // surface_follow: heightfields [N][2][64*64] (one live, one being / been built for the next episode), hf_cur [N] = the
    // live one, hf_meta [N][2][SURF_META]; partial resets: noise permutation [N][256], next grid point [N], float min/max [N][2]
    double *height, *hf_meta;
    double* accum;           // [N] surface_follow, reward_mode "sparse": the episode's accumulated dense reward (accum_rew)
    int* hf_cur;
    unsigned char* sb_perm;
    int* sb_surf_it;
    float* sb_hmm;
    // camera / stimulus of the state an env terminated in (for the terminal observation)
    double *term_cam, *term_stim;
    int* error_flag;         // sticky error flag (unused slots of the pipeline; kept for the C-ABI)
    int* stall_count;        // episode ends that had to complete their standby inline
    int* nan_count;          // env steps that ended in a non-finite state (the episode is ended, error flag bit 2)
};

// SB_CONSUMED + epoch: swapped in during launch `epoch`.  It is EMPTY for every later launch, but the launch that consumed it
// must not start the rebuild: for surface_follow the rebuild overwrites the heightfield of the episode that just ended, which
// the terminal-observation raster still reads after this launch.
enum { SB_EMPTY = 0, SB_READY = 1, SB_PARTIAL = 2, SB_BUSY = 3, SB_CONSUMED = 16 };
#define SB_EPOCH_MASK 0x0fffffff
// Quantum sizes of the resumable reset: blocking-move substeps and IK iterations per quantum are chosen per world from the
// episode length (tg_create: 2 and 3 at max_steps >= 200); heightfield points (OpenSimplex evaluations) per quantum:
#define SURF_CHUNK 96

// a reset in flight (between draws + IK and the end of the blocking move)
template <int NB>
struct ResetState {
    double q[NB], qd[NB], targ_j[NB], cv, embed, edge_ang, draw[TG_MAXDRAW];
    int nsteps;
    int ik_it; // >= 0: the IK solve is at this iteration; -1: solved, the blocking move is running
    int surf_it;       // surface_follow: next heightfield point to generate (SURF_PTS = done)
    float hmin, hmax;  // running min / max of the float32 heights (the drawn mesh is centred on their middle)
};

// one env's start-of-episode state
template <int NB>
struct EpisodeStart {
    double q[NB], qd[NB], embed, edge_ang, cam[12], stim[12], tcp[7];
    int substeps;
    ObjState obj; // object_balance, object_push
    double traj[PUSH_TRAJ_SZ]; // object_push only
    int goal;
};

TGD void obj_load(const double* o13, const double* e4, double g, double embed, const TgTask& task, ObjState& o)
{
#pragma unroll
    for (int c = 0; c < 3; c++) { o.pos[c] = o13[c]; o.vel[c] = o13[7 + c]; o.omg[c] = o13[10 + c]; o.ext_pos[c] = e4[c]; }
#pragma unroll
    for (int c = 0; c < 4; c++) o.quat[c] = o13[3 + c];
    o.ext_pending = e4[3] != 0.0;
    o.grav_z = g;
    o.mass = g; // object_push keeps the cube's mass in the episode scalar
    o.pivot_z = -task.obj_base_h * 0.5 + embed; // update_constraints (object_balance_env.py:285-293)
}
TGD void obj_store(double* o13, double* e4, const ObjState& o)
{
#pragma unroll
    for (int c = 0; c < 3; c++) { o13[c] = o.pos[c]; o13[7 + c] = o.vel[c]; o13[10 + c] = o.omg[c]; e4[c] = o.ext_pos[c]; }
#pragma unroll
    for (int c = 0; c < 4; c++) o13[3 + c] = o.quat[c];
    e4[3] = o.ext_pending ? 1.0 : 0.0;
}
// stimulus frame of the object = its base LINK frame: R(quat), pos - R base_com
TGD void obj_stim(const TgTask& task, const ObjState& o, double* stim)
{
    double R[9], t[3];
    mat_from_quat(o.quat, R);
    m3mulv(t, R, task.obj_base_com);
#pragma unroll
    for (int i = 0; i < 9; i++) stim[i] = R[i];
#pragma unroll
    for (int i = 0; i < 3; i++) stim[9 + i] = o.pos[i] - t[i];
}
// object_balance reward / termination (object_balance_env.py:470-526): roll / pitch more than 35 deg from the initial
// orientation, or the base more than 0.1 m from its initial position
TGD void balance_step_data(const TgTask& task, const ObjState& o, double embed, int steps, float* reward, unsigned char* done)
{
    double rpy[3];
    euler_from_quat(o.quat, rpy);
    bool fall = false;
#pragma unroll
    for (int c = 0; c < 2; c++) {
        const double cur = rpy[c] * 180.0 / M_PI, ini = task.obj_init_rpy[c] * 180.0 / M_PI;
        double m = fmod((cur - ini) + 180.0, 360.0);
        if (m < 0) m += 360.0; // numpy's % has the sign of the divisor
        if (fabs(m - 180.0) > task.obj_term_deg) fall = true;
    }
    const double ip[3] = {task.workframe_pos[0], task.workframe_pos[1], task.workframe_pos[2] + task.obj_base_h * 0.5 - embed};
    const double dx = o.pos[0] - ip[0], dy = o.pos[1] - ip[1], dz = o.pos[2] - ip[2];
    if (sqrt(dx * dx + dy * dy + dz * dz) > task.obj_term_pos) fall = true;
    *reward = task.sparse_reward ? (fall ? -1.0f : 0.0f) : 1.0f; // sparse_reward (:508-518) / dense_reward (:520-526)
    *done = (fall || steps >= task.max_steps) ? 1 : 0;
}

TGD int env_index(const EnvBuffers& b)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (lane >= b.lanes) return -1;
    const int e = warp * b.lanes + lane;
    return e < b.n ? e : -1;
}

// camera frame (sensors/tactile_sensor.py:150-229): eye, forward = R ex, up = R ez, then computeViewMatrix's
// orthonormalisation: f = norm(target - eye), s = norm(f x up), u = s x f
TGD void camera_from_frame(const double* pos, const double* R, double* cam)
{
    double f[3] = {R[0], R[3], R[6]}, u[3] = {R[2], R[5], R[8]}, s[3];
    double fn = 1.0 / sqrt(v3dot(f, f)); f[0] *= fn; f[1] *= fn; f[2] *= fn;
    double un = 1.0 / sqrt(v3dot(u, u)); u[0] *= un; u[1] *= un; u[2] *= un;
    v3cross(s, f, u);
    double sn = 1.0 / sqrt(v3dot(s, s)); s[0] *= sn; s[1] *= sn; s[2] *= sn;
    v3cross(u, s, f);
#pragma unroll
    for (int c = 0; c < 3; c++) { cam[c] = pos[c]; cam[3 + c] = f[c]; cam[6 + c] = u[c]; cam[9 + c] = s[c]; }
}
template <class T>
TGD void write_camera(const TgArm& arm, const Kin<T::NB>& k, double* cam)
{
    double pos[3], R[9];
#pragma unroll
    for (int b = 0; b < T::NB; b++)
        if (arm.cam_body == b) frame_pose<T::NB>(k, b, arm.cam_pos, arm.cam_rot, pos, R);
    camera_from_frame(pos, R, cam);
}

// edge_follow reward / termination (edge_follow_env.py:371-452)
TGD void edge_step_data(const TgTask& task, const double* tcp_pos, double edge_ang, int steps, float* reward, unsigned char* done)
{
    double s, c;
    sincos(edge_ang, &s, &c);
    const double gx = task.edge_pos[0] + task.edge_len * c, gy = task.edge_pos[1] + task.edge_len * s;
    const double p1x = task.edge_pos[0] - task.edge_len * c, p1y = task.edge_pos[1] - task.edge_len * s;
    const double dx = tcp_pos[0] - gx, dy = tcp_pos[1] - gy;
    const double goal_dist = sqrt(dx * dx + dy * dy);
    // |cross(p2 - p1, p1 - p3)| / |p2 - p1|
    const double ex = gx - p1x, ey = gy - p1y;
    const double fx = p1x - tcp_pos[0], fy = p1y - tcp_pos[1];
    const double edge_dist = fabs(ex * fy - ey * fx) / sqrt(ex * ex + ey * ey);
    *reward = (float)(-((1.0 * goal_dist) + (10.0 * edge_dist) + (1.0 * 0.0)));
    if (task.sparse_reward) *reward = goal_dist < task.termination_dist ? 1.0f : 0.0f; // sparse_reward (:430-438)
    *done = (goal_dist < task.termination_dist || steps >= task.max_steps) ? 1 : 0;
}

// pb.calculateInverseKinematics as restated in oracle/tg_oracle.c:or_inverse_kinematics (base_robot_arm.py:201-209)
// Resumable: runs iterations [it, it + count) of the <= 100 and returns the next iteration, or -1 once the solve has
// ended (converged or out of iterations).
template <class T>
TGD int ik_chunk(const TgArm& arm, double* q, const double* tpos, const double* tquat, int it, int count)
{
    constexpr int NB = T::NB;
    static_assert(NB == 6 || NB == 8, "topology");
#pragma unroll 1
    for (int c = 0; c < count; c++, it++) {
        if (it >= 100) return -1;
        Kin<NB> k;
        fk<T>(arm, q, k);
        double tp[3], tq[4], e[6];
        tcp_world<T>(arm, k, tp, tq);
        e[0] = tpos[0] - tp[0]; e[1] = tpos[1] - tp[1]; e[2] = tpos[2] - tp[2];
        const double res = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
        if (it > 0 && res < 1e-8) return -1;
        double qi[4] = {-tq[0], -tq[1], -tq[2], tq[3]}, dq[4];
        quat_mul(dq, tquat, qi);
        const double wv = fmin(fmax(dq[3], -1.0), 1.0);
        double ang = 2 * acos(wv);
        const double sn = sqrt(dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2]);
        if (ang > M_PI) ang -= 2 * M_PI;
#pragma unroll
        for (int c = 0; c < 3; c++) e[3 + c] = sn > 1e-300 ? ang * dq[c] / sn : 0.0;
        double J[6][NB];
        tcp_jacobian<T>(arm, k, tp, J);
        {
            double Mx[NB][NB + 1], d[NB];
#pragma unroll
            for (int i = 0; i < NB; i++) {
                double bi = 0;
#pragma unroll
                for (int r = 0; r < 6; r++) bi += J[r][i] * e[r];
                Mx[i][NB] = bi;
#pragma unroll
                for (int j = 0; j < NB; j++) {
                    double s = i == j ? 0.5 : 0.0;
#pragma unroll
                    for (int r = 0; r < 6; r++) s += J[r][i] * J[r][j];
                    Mx[i][j] = s;
                }
            }
            solveN<NB>(Mx, d);
            double mx = 0;
#pragma unroll
            for (int i = 0; i < NB; i++) mx = fmax(mx, fabs(d[i]));
            const double sc = mx > M_PI / 4 ? (M_PI / 4) / mx : 1.0;
#pragma unroll
            for (int i = 0; i < NB; i++) q[i] += sc * d[i];
        }
    }
    return it >= 100 ? -1 : it;
}

// workframe_to_worldframe (base_robot_arm.py:47-60) of the init pose [0,0,embed], init_rpy
// (surface_follow: centre_h = the new surface's height at the grid centre, base_surface_env.py:549-573)
TGD void reset_target(const TgTask& task, double embed, double centre_h, double* tpos, double* targ_orn)
{
    const bool balance = task.task == TG_TASK_OBJECT_BALANCE || task.task == TG_TASK_OBJECT_PUSH || task.task == TG_TASK_OBJECT_ROLL;
    double wq[4], tq[4], R[9], t[3], oq[4], rpy[3];
    const double lp[3] = {0.0, 0.0, balance ? 0.0 : embed}; // update_init_pose: edge_follow_env.py:301-309 / base_object_env.py:104-111
    quat_from_euler(task.workframe_rpy, wq);
    quat_from_euler(task.init_rpy, tq);
    mat_from_quat(wq, R);
    m3mulv(t, R, lp);
    tpos[0] = task.workframe_pos[0] + t[0]; tpos[1] = task.workframe_pos[1] + t[1]; tpos[2] = task.workframe_pos[2] + t[2];
    if (task.task == TG_TASK_SURFACE_FOLLOW) { tpos[0] = task.surf_pos[0]; tpos[1] = task.surf_pos[1]; tpos[2] = task.surf_pos[2] + centre_h - embed; }
    if (task.task == TG_TASK_SURFACE_FOLLOW && task.surf_vertical) { tpos[0] = task.surf_pos[0] - (centre_h - embed); tpos[2] = task.surf_pos[2]; } // :556-563
    if (task.task == TG_TASK_OBJECT_ROLL) tpos[2] = centre_h; // update_workframe (object_roll_env.py:197-202): the caller passes 2 r - embed_dist
    quat_mul(oq, wq, tq);
    euler_from_quat(oq, rpy);
    quat_from_euler(rpy, targ_orn);
}

// workframe_to_worldframe (base_robot_arm.py:47-60) + getQuaternionFromEuler of the resulting rpy
TGD void work_to_world(const TgTask& task, const double* pos, const double* rpy, double* wpos, double* wquat)
{
    double wq[4], tq[4], R[9], t[3], oq[4], r2[3];
    quat_from_euler(task.workframe_rpy, wq);
    quat_from_euler(rpy, tq);
    mat_from_quat(wq, R);
    m3mulv(t, R, pos);
    wpos[0] = task.workframe_pos[0] + t[0]; wpos[1] = task.workframe_pos[1] + t[1]; wpos[2] = task.workframe_pos[2] + t[2];
    quat_mul(oq, wq, tq);
    euler_from_quat(oq, r2);
    quat_from_euler(r2, wquat);
}

// Reset, part 1: consume the env's next draws, rest pose; the IK of the start pose starts in reset_advance.
template <class T>
__device__ __noinline__ void reset_begin(const TgArm& arm, const TgTask& task, const EnvBuffers& b, int e, ResetState<T::NB>& r)
{
    constexpr int NB = T::NB;
    // reset_task draws (edge_follow_env.py:285-299): embed_dist then edge_ang
#pragma unroll
    for (int d = 0; d < TG_MAXDRAW; d++) r.draw[d] = task.draw_default[d];
    {
        const int cnt = b.reset_count[e];
        if (b.mt_active) {
            // the env's own generator, numpy call semantics, the reference's call order (tg_rng.cuh)
            MtState ms{b.mt + (size_t)e * MT_N, b.mt_pos[e]};
#pragma unroll 1
            for (int d = 0; d < task.n_draws; d++) {
                if (task.surf_mode == 4 && task.task == TG_TASK_SURFACE_FOLLOW && d == 1) break; // drawn after the heights (reset_advance)
                r.draw[d] = mt_draw(ms, task.draw_kind[d], task.draw_lo[d], task.draw_hi[d], task.draw_default[d]);
            }
            b.mt_pos[e] = ms.pos;
        } else if (b.draws) {
            if (cnt < b.draw_avail[e]) {
                const int slot = cnt % b.draw_rounds;
#pragma unroll
                for (int d = 0; d < TG_MAXDRAW; d++)
                    if (d < task.n_draws) r.draw[d] = b.draws[((size_t)e * b.draw_rounds + slot) * task.n_draws + d];
            } else atomicOr(b.error_flag, 2); // draws exhausted: the defaults are used and the host is told (tg_draws_poll)
        }
        b.reset_count[e] = cnt + 1;
    }
    const bool balance = task.task == TG_TASK_OBJECT_BALANCE;
    // edge_follow draws: embed_dist, edge_ang; object_balance draws: gravity_z, embed_dist, fx, fy;
    // surface_follow draws: OpenSimplex seed, goal direction angle (kept in edge_ang)
    // object_push draws: init_obj_ang, obj_mass, OpenSimplex seed | trajectory direction (consumed in reset_finish)
    r.embed = balance ? r.draw[1] : r.draw[0];
    r.edge_ang = balance ? 0.0 : r.draw[1];
    if (task.task == TG_TASK_OBJECT_PUSH) { r.embed = 0.0; r.edge_ang = 0.0; }
    // object_roll draws: scaling_factor, embed_dist, dx, dy, goal_ang, goal_dist
    if (task.task == TG_TASK_OBJECT_ROLL) { r.embed = r.draw[1]; r.edge_ang = 0.0; }
    r.surf_it = SURF_PTS; r.hmin = 0.f; r.hmax = 0.f;
    if (task.task == TG_TASK_SURFACE_FOLLOW) {
        r.embed = task.surf_embed;
        os_perm((long long)r.draw[0], b.sb_perm + (size_t)e * 256);
        r.surf_it = 0; r.hmin = 3.0e38f; r.hmax = -3.0e38f;
    }
#pragma unroll
    for (int i = 0; i < NB; i++) { r.q[i] = b.rest_q[i]; r.qd[i] = 0.0; r.targ_j[i] = r.q[i]; }
    r.cv = 0.001;
    r.nsteps = 0;
    r.ik_it = 0;
}

// Reset, part 2, one quantum: IK_CHUNK iterations of the IK of the start pose (base_robot_arm.py:201-209), or, once
// that is solved, up to RESET_CHUNK iterations of Robot.blocking_move(max_steps=1000, constant_vel=0.001)
// (robot.py:188-260).  Returns true when the move has ended.
template <class T>
__device__ __noinline__ bool reset_advance(const TgArm& arm, const TgPhysics& ph, const TgTask& task, const EnvBuffers& b, int e, ResetState<T::NB>& r)
{
    constexpr int NB = T::NB;
    double centre_h = 0.0;
    if (task.task == TG_TASK_SURFACE_FOLLOW) {
        // the next episode's heightfield goes into the buffer the live episode does not use
        double* H = b.height + ((size_t)e * 2 + (size_t)(1 - b.hf_cur[e])) * SURF_PTS;
        if (r.surf_it < SURF_PTS) {
            // gen_heigtfield_simplex_2d (base_surface_env.py:311-327): h[x][y] = noise2(x * 0.05, y * 0.05) * 0.025
            const unsigned char* perm = b.sb_perm + (size_t)e * 256;
            const int end = min(r.surf_it + b.surf_chunk, SURF_PTS);
            if (task.surf_mode == 4 && b.mt_active) {
                // gen_heigtfield_noisey (base_surface_env.py:302-318): one uniform(0, 0.2 range) per 2 x 2 block, columns outer,
                // rows inner; then make_goal's draw (:508 / :513), which the reference makes after the surface
                MtState ms{b.mt + (size_t)e * MT_N, b.mt_pos[e]};
                const int half = SURF_N / 2;
#pragma unroll 1
                for (int t = r.surf_it / 4; t < end / 4; t++) {
                    const int jj = t / half, ii = t % half;
                    const double h = mt_uniform(ms, 0.0, task.surf_range * 0.2);
                    H[(2 * ii) * SURF_N + 2 * jj] = h; H[(2 * ii + 1) * SURF_N + 2 * jj] = h;
                    H[(2 * ii) * SURF_N + 2 * jj + 1] = h; H[(2 * ii + 1) * SURF_N + 2 * jj + 1] = h;
                    r.hmin = fminf(r.hmin, (float)h); r.hmax = fmaxf(r.hmax, (float)h);
                }
                if (end == SURF_PTS) {
                    r.draw[1] = mt_draw(ms, task.draw_kind[1], task.draw_lo[1], task.draw_hi[1], task.draw_default[1]);
                    r.edge_ang = r.draw[1];
                }
                b.mt_pos[e] = ms.pos;
                r.surf_it = end;
                return false;
            }
#pragma unroll 1
            for (int k = r.surf_it; k < end; k++) {
                // surf_mode 1: gen_heigtfield_simplex_1d (:339-357), noise along y only; 2: noise_mode "none" (:436-437)
                // 3: gen_heigtfield_simplex_1d_vertical (:359-379), noise along the rows only
                const double nx = task.surf_mode == 1 ? 1.0 * task.surf_interp : (double)(k / SURF_N) * task.surf_interp;
                const double ny = task.surf_mode == 3 ? 1.0 * task.surf_interp : (double)(k % SURF_N) * task.surf_interp;
                const double h = (task.surf_mode == 2 || task.surf_mode == 4) ? 0.0 : os_noise2(perm, nx, ny) * task.surf_range;
                H[k] = h;
                r.hmin = fminf(r.hmin, (float)h); r.hmax = fmaxf(r.hmax, (float)h);
            }
            r.surf_it = end;
            return false;
        }
        centre_h = H[(SURF_N / 2) * SURF_N + SURF_N / 2];
    }
    if (task.task == TG_TASK_OBJECT_ROLL) centre_h = 2.0 * (task.roll_radius * r.draw[0]) - r.embed; // this episode's workframe height
    double tpos[3], targ_orn[4];
    reset_target(task, r.embed, centre_h, tpos, targ_orn);
    if (r.ik_it >= 0) {
        r.ik_it = ik_chunk<T>(arm, r.targ_j, tpos, targ_orn, r.ik_it, b.ik_chunk);
        return false;
    }
    Motors<NB> mot;
    mot.mode = 1; mot.kp = ph.pos_gain; mot.kd = ph.vel_gain; mot.max_force = ph.blocking_force;
    double sc[NB][2];
#pragma unroll
    for (int i = 0; i < NB; i++) sincos(r.q[i], &sc[i][0], &sc[i][1]);
#pragma unroll 1
    for (int it = 0; it < b.reset_chunk; it++) {
        if (r.nsteps >= 1000) return true;
        double tp[3], tq[4];
        {
            Kin<NB> k;
            fk_sc<T>(arm, sc, k);
            tcp_world<T>(arm, k, tp, tq);
        }
        double nrm = 0, tot = 0;
        bool all_small = true;
        double diff[NB];
#pragma unroll
        for (int i = 0; i < NB; i++) { diff[i] = r.targ_j[i] - r.q[i]; nrm += diff[i] * diff[i]; tot += fabs(r.qd[i]); }
        nrm = sqrt(nrm);
#pragma unroll
        for (int i = 0; i < NB; i++) {
            const double vdir = nrm > 0 ? diff[i] / nrm : 0.0;
            mot.target_pos[i] = r.q[i] + vdir * r.cv; mot.target_vel[i] = 0.0;
            if (!(fabs(diff[i]) < r.cv)) all_small = false;
        }
        if (all_small) r.cv *= 0.5;
        substep<T>(arm, ph, r.q, r.qd, sc, mot);
        r.nsteps++;
        const double pe = fabs(tpos[0] - tp[0]) + fabs(tpos[1] - tp[1]) + fabs(tpos[2] - tp[2]);
        const double ip = targ_orn[0] * tq[0] + targ_orn[1] * tq[1] + targ_orn[2] * tq[2] + targ_orn[3] * tq[3];
        const double ca = fmin(fmax(2 * ip * ip - 1, -1.0), 1.0);
        const double oe = acos(ca);
        if (pe < 2e-4 && oe < 1e-3 && tot < 0.1) return true;
    }
    return r.nsteps >= 1000;
}

// Reset, part 3: the start-of-episode state the step and raster kernels consume.
template <class T>
__device__ __noinline__ void reset_finish(const TgArm& arm, const TgTask& task, const EnvBuffers& b, int e, const ResetState<T::NB>& r, EpisodeStart<T::NB>& out)
{
    constexpr int NB = T::NB;
    const bool balance = task.task == TG_TASK_OBJECT_BALANCE;
    out.embed = r.embed; out.edge_ang = r.edge_ang;
    out.substeps = r.nsteps;
#pragma unroll
    for (int i = 0; i < NB; i++) { out.q[i] = r.q[i]; out.qd[i] = r.qd[i]; }
    {
        Kin<NB> k;
        fk<T>(arm, r.q, k);
        double tp[3], tq[4];
        tcp_world<T>(arm, k, tp, tq);
        write_camera<T>(arm, k, out.cam);
#pragma unroll
        for (int c = 0; c < 3; c++) out.tcp[c] = tp[c];
#pragma unroll
        for (int c = 0; c < 4; c++) out.tcp[3 + c] = tq[c];
    }
    if (task.task == TG_TASK_SURFACE_FOLLOW) {
        // make_goal (base_surface_env.py:501-537) on the new surface; the drawn mesh's height offset
        const int sbuf = 1 - b.hf_cur[e];
        const double* H = b.height + ((size_t)e * 2 + (size_t)sbuf) * SURF_PTS;
        double* meta = b.hf_meta + ((size_t)e * 2 + (size_t)sbuf) * SURF_META;
        double sn, cs, wq[4], R[9];
        sincos(r.edge_ang, &sn, &cs);
        // yz / yzRx: workframe_directions = (0, choice([-1, 1]), 0) (:512-514), the draw itself
        const double dirs[3] = {task.surf_dir_mode ? 0.0 : cs, task.surf_dir_mode ? r.edge_ang : sn, 0.0};
        double wd[3];
        quat_from_euler(task.workframe_rpy, wq);
        mat_from_quat(wq, R);
        m3mulv(wd, R, dirs);
        const double gx = task.surf_pos[0] + task.surf_extent * wd[0], gy = task.surf_pos[1] + task.surf_extent * wd[1];
        int gi, gj;
        surf_index(task, gx, gy, gi, gj);
        meta[0] = 0.5 * ((double)r.hmin + (double)r.hmax);
        meta[1] = dirs[0]; meta[2] = dirs[1];
        meta[3] = gx; meta[4] = gy; meta[5] = H[gi * SURF_N + gj] + task.surf_pos[2];
        if (task.surf_vertical) {
            // :523-527 the goal is the (flipped) surface point itself: local (X - sx, Y - sy, h) -> surf_pos + (-h, Y - sy, X - sx)
            meta[3] = task.surf_pos[0] - H[gi * SURF_N + gj];
            meta[4] = surf_bin(task.surf_pos[1], task.surf_grid, gi);
            meta[5] = task.surf_pos[2] + (surf_bin(task.surf_pos[0], task.surf_grid, gj) - task.surf_pos[0]);
        }
        meta[6] = (double)r.hmin; meta[7] = (double)r.hmax; // float32 height range (the raster's first slab bound)
#pragma unroll
        for (int i = 0; i < 12; i++) out.stim[i] = 0.0;
    } else if (task.task == TG_TASK_OBJECT_ROLL) {
        // reset_object (object_roll_env.py:204-236): the marble back on the table under the tip (+ random offset), at rest;
        // make_goal (:238-256): goal at (dist, ang) in the TCP frame.  The episode's radius and workframe height ride in
        // the object's spare slots (ext_pos[0], ext_pos[1]); the goal in traj[0..1].
        ObjState& o = out.obj;
        const double rad = task.roll_radius * r.draw[0];
        o.pos[0] = task.push_init_pos[0] + r.draw[2]; o.pos[1] = task.push_init_pos[1] + r.draw[3]; o.pos[2] = rad;
        o.quat[0] = 0.0; o.quat[1] = 0.0; o.quat[2] = 0.0; o.quat[3] = 1.0;
#pragma unroll
        for (int c = 0; c < 3; c++) { o.vel[c] = 0.0; o.omg[c] = 0.0; }
        o.ext_pos[0] = rad; o.ext_pos[1] = 2.0 * rad - r.embed; o.ext_pos[2] = 0.0;
        o.ext_pending = 0; o.pivot_z = 0.0;
        o.mass = task.obj_mass; o.grav_z = task.obj_mass;
#pragma unroll 1
        for (int c = 0; c < PUSH_TRAJ_SZ; c++) out.traj[c] = 0.0;
        double sn, cs;
        sincos(r.draw[4], &sn, &cs);
        out.traj[0] = r.draw[5] * cs; out.traj[1] = r.draw[5] * sn;
        out.goal = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) out.stim[i] = 0.0;
        out.stim[0] = rad;
#pragma unroll
        for (int i = 0; i < 3; i++) out.stim[9 + i] = o.pos[i];
    } else if (task.task == TG_TASK_OBJECT_PUSH) {
        // reset_object (object_push_env.py:204-229): cube back at init_obj_pos, yaw pi/2 + init_obj_ang, at rest;
        // make_goal (:322-333): new trajectory, first goal; BaseObjectEnv.reset then calls get_step_data() once, whose
        // termination() already advances the goal when the cube starts within termination_pos_dist of it
        ObjState& o = out.obj;
        const double rpy[3] = {task.obj_init_rpy[0], task.obj_init_rpy[1], task.obj_init_rpy[2] + r.draw[0]};
        quat_from_euler(rpy, o.quat);
#pragma unroll
        for (int c = 0; c < 3; c++) { o.pos[c] = task.push_init_pos[c]; o.vel[c] = 0.0; o.omg[c] = 0.0; o.ext_pos[c] = 0.0; }
        o.ext_pending = 0; o.pivot_z = 0.0;
        o.mass = r.draw[1]; o.grav_z = r.draw[1];
        push_trajectory(task, r.draw[2], out.traj);
        out.traj[2 * PUSH_NTRAJ] = r.draw[2];
        out.goal = 0;
        {
            float rw; unsigned char dn;
            push_step_data(task, o, out.tcp + 3, out.traj, r.draw[2], out.goal, 0, &rw, &dn);
        }
        obj_stim(task, o, out.stim);
    } else if (!balance) {
        double sn, cs;
        sincos(r.edge_ang * 0.5, &sn, &cs);
        double qz[4] = {0, 0, sn, cs}, R[9]; // getQuaternionFromEuler([0,0,ang]) (edge_follow_env.py:241)
        mat_from_quat(qz, R);
#pragma unroll
        for (int i = 0; i < 9; i++) out.stim[i] = R[i];
#pragma unroll
        for (int i = 0; i < 3; i++) out.stim[9 + i] = task.edge_pos[i];
    } else {
        // reset_object (object_balance_env.py:322-358): pole back at init_obj_pos / init_obj_orn, at rest, 0.1 N pushing
        // down at a random point of the base plate during the next stepSimulation (apply_random_force_base :360-381)
        ObjState& o = out.obj;
        o.pos[0] = task.workframe_pos[0]; o.pos[1] = task.workframe_pos[1];
        o.pos[2] = task.workframe_pos[2] + task.obj_base_h * 0.5 - r.embed;
        quat_from_euler(task.obj_init_rpy, o.quat);
#pragma unroll
        for (int c = 0; c < 3; c++) { o.vel[c] = 0.0; o.omg[c] = 0.0; }
        o.ext_pos[0] = o.pos[0] + r.draw[2] * task.obj_base_w * 0.5;
        o.ext_pos[1] = o.pos[1] + r.draw[3] * task.obj_base_w * 0.5;
        o.ext_pos[2] = o.pos[2];
        o.ext_pending = 1;
        o.grav_z = r.draw[0];
        o.pivot_z = -task.obj_base_h * 0.5 + r.embed;
        obj_stim(task, o, out.stim);
    }
}

// One whole reset of env e (same chunk schedule as the resumable path)
template <class T>
TGD void reset_env(const TgArm& arm, const TgPhysics& ph, const TgTask& task, const EnvBuffers& b, int e, EpisodeStart<T::NB>& out)
{
    ResetState<T::NB> r;
    reset_begin<T>(arm, task, b, e, r);
#pragma unroll 1
    while (!reset_advance<T>(arm, ph, task, b, e, r)) {}
    reset_finish<T>(arm, task, b, e, r, out);
}

template <int NB>
TGD void store_live(const EnvBuffers& b, int e, const EpisodeStart<NB>& s)
{
#pragma unroll
    for (int i = 0; i < NB; i++) { b.q[(size_t)i * b.n + e] = s.q[i]; b.qd[(size_t)i * b.n + e] = s.qd[i]; }
    b.embed[e] = s.embed; b.edge_ang[e] = s.edge_ang; b.steps[e] = 0; b.reset_substeps[e] = s.substeps;
#pragma unroll
    for (int c = 0; c < 12; c++) { b.cam[(size_t)e * 12 + c] = s.cam[c]; b.stim[(size_t)e * 12 + c] = s.stim[c]; }
#pragma unroll
    for (int c = 0; c < 7; c++) b.tcp[(size_t)e * 7 + c] = s.tcp[c];
    if (b.obj) { obj_store(b.obj + (size_t)e * 13, b.obj_ext + (size_t)e * 4, s.obj); b.grav[e] = s.obj.grav_z; }
    if (b.traj) {
#pragma unroll 1
        for (int c = 0; c < PUSH_TRAJ_SZ; c++) b.traj[(size_t)e * PUSH_TRAJ_SZ + c] = s.traj[c];
        b.goal[e] = s.goal;
    }
}

template <int NB>
TGD void store_standby(const EnvBuffers& b, int e, const EpisodeStart<NB>& s)
{
#pragma unroll
    for (int i = 0; i < NB; i++) { b.sb_q[(size_t)i * b.n + e] = s.q[i]; b.sb_qd[(size_t)i * b.n + e] = s.qd[i]; }
    b.sb_embed[e] = s.embed; b.sb_ang[e] = s.edge_ang; b.sb_substeps[e] = s.substeps;
#pragma unroll
    for (int c = 0; c < 12; c++) { b.sb_cam[(size_t)e * 12 + c] = s.cam[c]; b.sb_stim[(size_t)e * 12 + c] = s.stim[c]; }
#pragma unroll
    for (int c = 0; c < 7; c++) b.sb_tcp[(size_t)e * 7 + c] = s.tcp[c];
    if (b.obj) { obj_store(b.sb_obj + (size_t)e * 13, b.sb_obj_ext + (size_t)e * 4, s.obj); b.sb_grav[e] = s.obj.grav_z; }
    if (b.traj) {
#pragma unroll 1
        for (int c = 0; c < PUSH_TRAJ_SZ; c++) b.sb_traj[(size_t)e * PUSH_TRAJ_SZ + c] = s.traj[c];
        b.sb_goal[e] = s.goal;
    }
    __threadfence();
    atomicExch(&b.sb_ready[e], SB_READY);
}

// partial reset <-> sb_* buffers (the slot is owned: SB_BUSY)
template <int NB>
TGD void store_partial(const EnvBuffers& b, int e, const ResetState<NB>& r)
{
#pragma unroll
    for (int i = 0; i < NB; i++) {
        b.sb_q[(size_t)i * b.n + e] = r.q[i]; b.sb_qd[(size_t)i * b.n + e] = r.qd[i]; b.sb_targ[(size_t)i * b.n + e] = r.targ_j[i];
    }
    b.sb_embed[e] = r.embed; b.sb_ang[e] = r.edge_ang; b.sb_substeps[e] = r.nsteps; b.sb_cv[e] = r.cv; b.sb_ik[e] = r.ik_it;
    if (b.height) { b.sb_surf_it[e] = r.surf_it; b.sb_hmm[2 * e] = r.hmin; b.sb_hmm[2 * e + 1] = r.hmax; }
#pragma unroll
    for (int d = 0; d < TG_MAXDRAW; d++) b.sb_draw[(size_t)e * TG_MAXDRAW + d] = r.draw[d];
    __threadfence();
    atomicExch(&b.sb_ready[e], SB_PARTIAL);
}
template <int NB>
TGD void load_partial(const EnvBuffers& b, int e, ResetState<NB>& r)
{
#pragma unroll
    for (int i = 0; i < NB; i++) {
        r.q[i] = __ldcg(b.sb_q + (size_t)i * b.n + e); r.qd[i] = __ldcg(b.sb_qd + (size_t)i * b.n + e); r.targ_j[i] = __ldcg(b.sb_targ + (size_t)i * b.n + e);
    }
    r.embed = __ldcg(b.sb_embed + e); r.edge_ang = __ldcg(b.sb_ang + e); r.nsteps = __ldcg(b.sb_substeps + e); r.cv = __ldcg(b.sb_cv + e); r.ik_it = __ldcg(b.sb_ik + e);
    r.surf_it = SURF_PTS; r.hmin = 0.f; r.hmax = 0.f;
    if (b.height) { r.surf_it = __ldcg(b.sb_surf_it + e); r.hmin = __ldcg(b.sb_hmm + 2 * e); r.hmax = __ldcg(b.sb_hmm + 2 * e + 1); }
#pragma unroll
    for (int d = 0; d < TG_MAXDRAW; d++) r.draw[d] = __ldcg(b.sb_draw + (size_t)e * TG_MAXDRAW + d);
}

// swap the standby in as the live state of env e (a copy); the slot is rebuilt by the following launches
template <int NB>
TGD void consume_standby(const EnvBuffers& b, int e)
{
#pragma unroll
    for (int i = 0; i < NB; i++) { b.q[(size_t)i * b.n + e] = __ldcg(b.sb_q + (size_t)i * b.n + e); b.qd[(size_t)i * b.n + e] = __ldcg(b.sb_qd + (size_t)i * b.n + e); }
    b.embed[e] = __ldcg(b.sb_embed + e); b.edge_ang[e] = __ldcg(b.sb_ang + e); b.steps[e] = 0; b.reset_substeps[e] = __ldcg(b.sb_substeps + e);
#pragma unroll
    for (int c = 0; c < 12; c++) { b.cam[(size_t)e * 12 + c] = __ldcg(b.sb_cam + (size_t)e * 12 + c); b.stim[(size_t)e * 12 + c] = __ldcg(b.sb_stim + (size_t)e * 12 + c); }
#pragma unroll
    for (int c = 0; c < 7; c++) b.tcp[(size_t)e * 7 + c] = __ldcg(b.sb_tcp + (size_t)e * 7 + c);
    if (b.obj) {
#pragma unroll
        for (int c = 0; c < 13; c++) b.obj[(size_t)e * 13 + c] = __ldcg(b.sb_obj + (size_t)e * 13 + c);
#pragma unroll
        for (int c = 0; c < 4; c++) b.obj_ext[(size_t)e * 4 + c] = __ldcg(b.sb_obj_ext + (size_t)e * 4 + c);
        b.grav[e] = __ldcg(b.sb_grav + e);
    }
    if (b.traj) {
#pragma unroll 1
        for (int c = 0; c < PUSH_TRAJ_SZ; c++) b.traj[(size_t)e * PUSH_TRAJ_SZ + c] = __ldcg(b.sb_traj + (size_t)e * PUSH_TRAJ_SZ + c);
        b.goal[e] = __ldcg(b.sb_goal + e);
    }
    if (b.height) b.hf_cur[e] = 1 - b.hf_cur[e]; // the heightfield built for this episode becomes the live one
    __threadfence();
    atomicExch(&b.sb_ready[e], SB_CONSUMED + (b.epoch & SB_EPOCH_MASK));
}

// get_extended_feature_array (object_roll_env.py:402-408): the goal position in the TCP frame
TGD void roll_features(const double* traj, float* out)
{
    out[0] = (float)traj[0]; out[1] = (float)traj[1]; out[2] = 0.0f;
#pragma unroll
    for (int i = 3; i < TG_PUSH_NFEAT; i++) out[i] = 0.0f;
}

// surface_follow, reward_mode "sparse": reset() runs get_step_data once (base_surface_env.py:638), so the episode's accumulator
// starts at the dense reward of the start pose (reset_task zeroes it just before, :590-591)
TGD void surface_accum_start(const TgTask& task, const EnvBuffers& b, int e)
{
    const size_t hb = (size_t)e * 2 + (size_t)b.hf_cur[e];
    float r; unsigned char d; double dense;
    surface_step_data(task, b.height + hb * SURF_PTS, b.hf_meta + hb * SURF_META, b.tcp + (size_t)e * 7, b.tcp + (size_t)e * 7 + 3, 0, &r, &d, &dense);
    b.accum[e] = dense;
}

// object_push / object_roll: the extended feature of env e's live state (after a reset / standby swap)
TGD void write_live_features(const TgTask& task, const EnvBuffers& b, int e)
{
    if (!b.feat) return;
    if (task.task == TG_TASK_SURFACE_FOLLOW) {
        const double* meta = b.hf_meta + ((size_t)e * 2 + (size_t)b.hf_cur[e]) * SURF_META;
        surface_features(task, meta, b.tcp + (size_t)e * 7, b.tcp + (size_t)e * 7 + 3, b.feat + (size_t)e * TG_PUSH_NFEAT);
        return;
    }
    if (!b.traj) return;
    const double* tr = b.traj + (size_t)e * PUSH_TRAJ_SZ;
    if (task.task == TG_TASK_OBJECT_ROLL) { roll_features(tr, b.feat + (size_t)e * TG_PUSH_NFEAT); return; }
    push_features(task, b.tcp + (size_t)e * 7, b.tcp + (size_t)e * 7 + 3, tr, tr[2 * PUSH_NTRAJ], b.goal[e], b.feat + (size_t)e * TG_PUSH_NFEAT);
}

// get_oracle_obs of env e's CURRENT state in the buffers (observation_mode "oracle": the state vector that replaces the
// tactile image).  edge_follow_env.py:454-476 (10 values), base_surface_env.py:789-819 (20), object_balance_env.py:528-563 (26),
// object_push_env.py:571-609 (30), object_roll_env.py:371-409 (34).  TCP pose / velocity in the work frame as
// get_current_TCP_pos_vel_workframe builds them (base_robot_arm.py:136-172): getLinkState's inertial-frame pose and the
// velocity of that point, J(q) qd.  Kept out of line: it runs once per env step and only when the buffer is bound.
template <class T>
__device__ __noinline__ void oracle_obs_env(const TgArm& arm, const TgTask& task, const EnvBuffers& b, int e, float* out)
{
    constexpr int NB = T::NB;
    double q[NB], qd[NB];
#pragma unroll
    for (int i = 0; i < NB; i++) { q[i] = b.q[(size_t)i * b.n + e]; qd[i] = b.qd[(size_t)i * b.n + e]; }
    Kin<NB> k;
    fk<T>(arm, q, k);
    double tp[3], tq[4], J[6][NB], vw[6];
    tcp_world<T>(arm, k, tp, tq);
    tcp_jacobian<T>(arm, k, tp, J);
#pragma unroll
    for (int r = 0; r < 6; r++) {
        double a = 0.0;
#pragma unroll
        for (int i = 0; i < NB; i++) a += J[r][i] * qd[i];
        vw[r] = a;
    }
    const bool roll = task.task == TG_TASK_OBJECT_ROLL;
    // object_roll moves the workframe every episode (object_roll_env.py:197-202)
    const double wf[3] = {task.workframe_pos[0], task.workframe_pos[1], roll ? b.obj_ext[(size_t)e * 4 + 1] : task.workframe_pos[2]};
    double wp[3], wr[3], wo[4], wq[4], wqi[4], Ri[9], vl[3], va[3];
    world_to_work_at(task, wf, tp, tq, wp, wr);
    quat_from_euler(wr, wo);
    // worldvel_to_workvel (base_robot_arm.py:107-118): the matrix of the inverted workframe quaternion
    quat_from_euler(task.workframe_rpy, wq);
    wqi[0] = -wq[0]; wqi[1] = -wq[1]; wqi[2] = -wq[2]; wqi[3] = wq[3];
    mat_from_quat(wqi, Ri);
    m3mulv(vl, Ri, vw); m3mulv(va, Ri, vw + 3);
    const double ident[4] = {0.0, 0.0, 0.0, 1.0};
    int n = 0;
    auto put3 = [&](const double* v) { out[n++] = (float)v[0]; out[n++] = (float)v[1]; out[n++] = (float)v[2]; };
    auto put4 = [&](const double* v) { out[n++] = (float)v[0]; out[n++] = (float)v[1]; out[n++] = (float)v[2]; out[n++] = (float)v[3]; };
    if (task.task == TG_TASK_EDGE_FOLLOW) {
        double s, c, gp[3], gr[3];
        sincos(b.edge_ang[e], &s, &c);
        const double g[3] = {task.edge_pos[0] + task.edge_len * c, task.edge_pos[1] + task.edge_len * s, task.edge_pos[2] + task.edge_height};
        world_to_work_at(task, wf, g, ident, gp, gr);
        put3(wp); put3(vl); put3(gp);
        out[n++] = (float)b.edge_ang[e];
    } else if (task.task == TG_TASK_SURFACE_FOLLOW) {
        const size_t hb = (size_t)e * 2 + (size_t)b.hf_cur[e];
        const double* H = b.height + hb * SURF_PTS;
        const double* meta = b.hf_meta + hb * SURF_META;
        int ti, tj;
        surf_index(task, tp[0], tp[1], ti, tj);
        double g0, g1, gp[3], gr[3], nw[3];
        surf_gradient(H, task.surf_grid, ti, tj, g0, g1);
        double nrm[3] = {-g1, -g0, 1.0};
        const double nn = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
        nrm[0] /= nn; nrm[1] /= nn; nrm[2] /= nn;
        double surf_h = H[ti * SURF_N + tj] + task.surf_pos[2];
        if (task.surf_vertical) {   // the flipped surface_array / surface_normals (:472-506)
            const double fl[3] = {-nrm[2], nrm[1], nrm[0]};
            nrm[0] = fl[0]; nrm[1] = fl[1]; nrm[2] = fl[2];
            surf_h = task.surf_pos[2] + (surf_bin(task.surf_pos[0], task.surf_grid, tj) - task.surf_pos[0]);
        }
        m3mulv(nw, Ri, nrm);
        const double g[3] = {meta[3], meta[4], meta[5]};
        world_to_work_at(task, wf, g, ident, gp, gr);
        put3(wp); put4(wo); put3(vl); put3(va); put3(gp);
        out[n++] = (float)surf_h;
        put3(nw);
    } else {
        // the free object: getBasePositionAndOrientation / getBaseVelocity brought to the work frame (base_object_env.py:118-139)
        const double* o13 = b.obj + (size_t)e * 13;
        double op[3], orr[3], oq[4], ol[3], oa[3];
        world_to_work_at(task, wf, o13, o13 + 3, op, orr);
        quat_from_euler(orr, oq);
        m3mulv(ol, Ri, o13 + 7); m3mulv(oa, Ri, o13 + 10);
        if (task.task == TG_TASK_OBJECT_PUSH) {
            const double* tr = b.traj + (size_t)e * PUSH_TRAJ_SZ;
            const int goal = b.goal[e], gi = goal < PUSH_NTRAJ ? (goal < 0 ? 0 : goal) : PUSH_NTRAJ - 1;
            const double gp[3] = {push_goal_x(task, tr[2 * PUSH_NTRAJ], gi), tr[gi], 0.0}, gr[3] = {0.0, 0.0, tr[PUSH_NTRAJ + gi]};
            put3(wp); put3(wr); put3(vl); put3(va); put3(op); put3(orr); put3(ol); put3(oa); put3(gp); put3(gr);
        } else {
            put3(wp); put4(wo); put3(vl); put3(va); put3(op); put4(oq); put3(ol); put3(oa);
            if (roll) {
                const double* tr = b.traj + (size_t)e * PUSH_TRAJ_SZ;
                const double gp[3] = {tr[0], tr[1], 0.0};
                put3(gp); put4(ident);
                out[n++] = (float)b.obj_ext[(size_t)e * 4];   // scaled_obj_radius
            }
        }
    }
    for (; n < TG_ORACLE_NOBS; n++) out[n] = 0.0f;
}

// Work on the standby slot of env e if it is free to take: EMPTY -> draws + IK, PARTIAL -> one chunk of the
// blocking move (complete: everything, to READY).
template <class T>
TGD void standby_work(const TgArm& arm, const TgPhysics& ph, const TgTask& task, const EnvBuffers& b, int e, bool complete)
{
    const int s = atomicAdd(&b.sb_ready[e], 0);
    if (s == SB_READY || s == SB_BUSY) return;
    if (!complete && s == SB_CONSUMED + (b.epoch & SB_EPOCH_MASK)) return; // consumed in this launch: the next one starts the rebuild
    if (atomicCAS(&b.sb_ready[e], s, SB_BUSY) != s) return;
    __threadfence();
    ResetState<T::NB> r;
    bool fin = false;
    if (s == SB_EMPTY || s >= SB_CONSUMED) reset_begin<T>(arm, task, b, e, r);
    else load_partial<T::NB>(b, e, r);
    fin = reset_advance<T>(arm, ph, task, b, e, r);
    if (complete) {
#pragma unroll 1
        while (!fin) fin = reset_advance<T>(arm, ph, task, b, e, r);
    }
    if (fin) {
        EpisodeStart<T::NB> es;
        reset_finish<T>(arm, task, b, e, r, es);
        store_standby<T::NB>(b, e, es);
    } else store_partial<T::NB>(b, e, r);
}

// the step thread of a finished env needs its standby NOW: wait for / complete the slot, then swap it in
template <class T>
TGD void acquire_standby(const TgArm& arm, const TgPhysics& ph, const TgTask& task, const EnvBuffers& b, int e)
{
    bool stalled = false;
#pragma unroll 1
    for (;;) {
        const int s = atomicAdd(&b.sb_ready[e], 0);
        if (s == SB_READY) break;
        stalled = true;
        if (s == SB_BUSY) { __nanosleep(500); continue; } // a standby thread holds the slot for one chunk
        standby_work<T>(arm, ph, task, b, e, true);
    }
    if (stalled) atomicAdd(b.stall_count, 1);
    __threadfence();
    consume_standby<T::NB>(b, e);
}

// standby role: threads scan the envs and advance every slot that is not READY
template <class T>
TGD void standby_role(const TgArm& arm, const TgPhysics& ph, const TgTask& task, const EnvBuffers& b, int first_block, bool complete)
{
    const int t = (blockIdx.x - first_block) * blockDim.x + threadIdx.x;
    const int nt = (gridDim.x - first_block) * blockDim.x;
    for (int e = t; e < b.n; e += nt) standby_work<T>(arm, ph, task, b, e, complete);
}

// autoreset: 1 = finished envs start their next episode inside this launch (VecEnv semantics)
// TASK is a template parameter so that each task's step is its own lean kernel (with all four tasks behind run-time
// branches the edge_follow step was 45 % slower than alone: registers, instruction cache).
// object_push runs PUSH_BLOCK envs per block, all lanes active, its constraint rows staged in dynamic shared memory, and
// has no standby blocks: its episodes are long and its reset short, so each step thread advances its own env's standby
// slot by one quantum after the step.
// control_mode TCP_position_control, one env step (robot.py:156-186): tcp_position_control (base_robot_arm.py:228-279) -
// the scaled action is a pose delta in the work frame, clipped to tcp_lims (check_TCP_pos_lims :349-355), brought to the world
// frame, solved by IK from the current joints, held by position motors - then blocking_move(max_steps, constant_vel=None)
// (robot.py:188-260): step until pose error and joint speed, both read BEFORE the step, pass.  Out of line: the velocity
// mode's registers stay what they were.
// the IK target of tcp_position_control and the joint targets it leads to; `wf` = the work frame's origin (object_roll moves
// it with the marble's size every episode, the other tasks pass task.workframe_pos)
template <class T>
__device__ __noinline__ void position_control_target(const TgArm& arm, const TgTask& task, const double* wf, const double* q, const double* delta,
                                                     Motors<T::NB>& mot, double* tpos, double* torn)
{
    constexpr int NB = T::NB;
    {
        Kin<NB> k;
        fk<T>(arm, q, k);
        double tp[3], tq[4], wp[3], wr[3], np_[3], nr[3];
        tcp_world<T>(arm, k, tp, tq);
        world_to_work_at(task, wf, tp, tq, wp, wr);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            np_[c] = fmin(fmax(wp[c] + delta[c], task.tcp_lims[c][0]), task.tcp_lims[c][1]);
            nr[c] = fmin(fmax(wr[c] + delta[3 + c], task.tcp_lims[3 + c][0]), task.tcp_lims[3 + c][1]);
        }
        work_to_world(task, np_, nr, tpos, torn);
#pragma unroll
        for (int c = 0; c < 3; c++) tpos[c] += wf[c] - task.workframe_pos[c];
    }
#pragma unroll
    for (int i = 0; i < NB; i++) { mot.target_pos[i] = q[i]; mot.target_vel[i] = 0.0; }
    for (int it = 0; it >= 0;) it = ik_chunk<T>(arm, mot.target_pos, tpos, torn, it, 100);
    if (arm.topo == TG_TOPO_MG400 && NB == 8) {
        // MG400.tcp_position_control (mg400.py:167-172): the parallelogram's slaved joints follow j2_1 / j3_1 in the IK result too
        mot.target_pos[NB - 3] = mot.target_pos[1];
        mot.target_pos[NB - 2] = -mot.target_pos[1];
        mot.target_pos[NB - 1] = mot.target_pos[1] + mot.target_pos[2];
    }
}

// blocking_move's exit test (robot.py:222-245) on the TCP pose and the joint speeds read BEFORE the step
template <class T>
TGD bool position_control_reached(const TgArm& arm, const double (&sc)[T::NB][2], const double* qd, const double* tpos, const double* torn)
{
    constexpr int NB = T::NB;
    double tp[3], tq[4], tot = 0.0;
    {
        Kin<NB> k;
        fk_sc<T>(arm, sc, k);
        tcp_world<T>(arm, k, tp, tq);
    }
#pragma unroll
    for (int i = 0; i < NB; i++) tot += fabs(qd[i]);
    const double pe = fabs(tpos[0] - tp[0]) + fabs(tpos[1] - tp[1]) + fabs(tpos[2] - tp[2]);
    const double ip = torn[0] * tq[0] + torn[1] * tq[1] + torn[2] * tq[2] + torn[3] * tq[3];
    const double oe = acos(fmin(fmax(2 * ip * ip - 1, -1.0), 1.0));
    return pe < 2e-4 && oe < 1e-3 && tot < 0.1;
}

// Robot.apply_action(control_mode="TCP_position_control") (robot.py:156-186): tcp_position_control (base_robot_arm.py:228-279:
// work-frame pose + delta, check_TCP_pos_lims, IK from the current joints, position motors at the env's max force), then
// blocking_move(max_steps=_max_blocking_pos_move_steps, constant_vel=None) (robot.py:188-260): step until pose error and joint
// speed, both read BEFORE the step, pass.  Out of line: the velocity mode's registers stay what they were.  `ob`: the
// constrained object of object_balance (enabled) steps with the arm.
template <class T>
__device__ __noinline__ void position_control_move(const TgArm& arm, const TgPhysics& ph, const TgTask& task, double* q, double* qd, const double* delta,
                                                   ObjState* ob)
{
    constexpr int NB = T::NB;
    double tpos[3], torn[4];
    Motors<NB> mot;
    mot.mode = 1; mot.kp = ph.pos_gain; mot.kd = ph.vel_gain; mot.max_force = ph.max_force;
    position_control_target<T>(arm, task, task.workframe_pos, q, delta, mot, tpos, torn);
    double sc[NB][2];
#pragma unroll
    for (int i = 0; i < NB; i++) sincos(q[i], &sc[i][0], &sc[i][1]);
#pragma unroll 1
    for (int s = 0; s < task.pos_max_steps; s++) {
        const bool reached = position_control_reached<T>(arm, sc, qd, tpos, torn);
        if (ob) substep_obj<T>(arm, ph, task, q, qd, sc, mot, *ob);
        else substep<T>(arm, ph, q, qd, sc, mot);
        if (reached) break;
    }
}

// Head of one env step for env e (BaseTactileEnv.step up to the physics): encode_actions + scale_actions, then - for
// TCP_velocity_control - check_TCP_vel_lims, the twist in the world frame and the joint velocity targets (mot.target_vel).
// v[6] = the scaled 6-vector (TCP_position_control consumes it as a pose delta).
template <class T, int TASK>
TGD void env_prologue_core(const TgArm& arm, const TgPhysics& ph, const TgTask& task, const EnvBuffers& b, int e, const float* __restrict__ actions,
                           const double* tp, const double* tq, const double (&J)[6][T::NB], double* v, Motors<T::NB>& mot)
{
    constexpr int NB = T::NB;
    constexpr bool roll = TASK == TG_TASK_OBJECT_ROLL, surface = TASK == TG_TASK_SURFACE_FOLLOW;
    constexpr bool push = TASK == TG_TASK_OBJECT_PUSH || roll;
    // encode_actions + scale_actions (edge_follow_env.py:345-369, base_tactile_env.py:141-164)
    mot.mode = 0; mot.kp = 0; mot.kd = ph.vel_gain; mot.max_force = ph.max_force;
    {
        double wq[4], Rw[9];
        quat_from_euler(task.workframe_rpy, wq);
        mat_from_quat(wq, Rw);
        double enc[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int kk = 0; kk < 6; kk++)
            if (kk < task.act_dim) {
                const double a = (double)actions[(size_t)e * task.act_dim + kk];
#pragma unroll
                for (int s = 0; s < 6; s++) if (task.act_index[kk] == s) enc[s] = a;
            }
        if (surface && task.surf_drive != 0.0) {
            // SurfaceFollowAutoEnv.encode_actions (surface_follow_auto_env.py:27-57): constant drive towards the goal
            // (surface_follow-v1 has none: the policy's own x / y stay)
            const double* meta = b.hf_meta + ((size_t)e * 2 + (size_t)b.hf_cur[e]) * SURF_META;
            if (!task.surf_drive_y_only) enc[0] = meta[1] * task.surf_drive;   // -v2 (SurfaceFollowVertEnv): x is the policy's
            enc[1] = meta[2] * task.surf_drive;
        }
        if (push && !roll) {
            // ObjectPushEnv.encode_actions (object_push_env.py:369-454)
            if (task.push_mode == TG_PUSH_WORK_DRIVE) enc[0] = task.act_max;
            if (task.push_mode == TG_PUSH_TCP_TYRZ || task.push_mode == TG_PUSH_TCP_TXTYRZ) {
                // directions along / across the tip in the TCP frame, brought to the work frame (worldvec_to_workvec)
                double Rt[9], parw[3], perpw[3], par[3], perp[3];
                mat_from_quat(tq, Rt);
                const double ex[3] = {1.0, 0.0, 0.0}, ey[3] = {0.0, -1.0, 0.0};
                m3mulv(parw, Rt, ex); m3mulv(perpw, Rt, ey);
                m3tmulv(par, Rw, parw); m3tmulv(perp, Rw, perpw);
                const bool ty = task.push_mode == TG_PUSH_TCP_TYRZ;
                const double a0 = (double)actions[(size_t)e * task.act_dim], a1 = (double)actions[(size_t)e * task.act_dim + 1];
                const double par_scale = ty ? 1.0 * task.act_max : a0, perp_scale = ty ? a0 : a1;
                const double rz = ty ? a1 : (double)actions[(size_t)e * task.act_dim + (task.act_dim > 2 ? 2 : 1)];
                enc[0] = 0.0; enc[1] = 0.0; enc[5] = 0.0;
                enc[0] += perp[0] * perp_scale + par[0] * par_scale;
                enc[1] += perp[1] * perp_scale + par[1] * par_scale;
                enc[5] += rz;
            }
        }
        const double in_range = task.act_max - task.act_min;
#pragma unroll
        for (int s = 0; s < 6; s++) {
            const double a = fmin(fmax(enc[s], task.act_min), task.act_max);
            v[s] = (((a - task.act_min) * (task.act_hi[s] - task.act_lo[s])) / in_range) + task.act_lo[s];
        }
        // tcp_velocity_control (base_robot_arm.py:281-332); TCP_position_control is handled by position_control_move below
        if (task.control_mode != 1) {
        double wp[3], wr[3];
        if (roll) {
            const double wf[3] = {task.workframe_pos[0], task.workframe_pos[1], b.obj_ext[(size_t)e * 4 + 1]}; // this episode's workframe
            world_to_work_at(task, wf, tp, tq, wp, wr);
        } else world_to_work(task, tp, tq, wp, wr);
#pragma unroll
        for (int s = 0; s < 6; s++) {
            const double cur = s < 3 ? wp[s] : wr[s - 3];
            const bool ex = (cur < task.tcp_lims[s][0] && v[s] < 0) || (cur > task.tcp_lims[s][1] && v[s] > 0);
            if (ex) v[s] = 0.0;
        }
        double vw[6];
        m3mulv(vw, Rw, v); m3mulv(vw + 3, Rw, v + 3);
        bool use_pinv = NB != 6 || arm.topo == TG_TOPO_MG400;   // mg400.py:109 always uses the pseudo-inverse
        if (NB == 6) {
            double Mx[6][7];
#pragma unroll
            for (int r = 0; r < 6; r++) {
#pragma unroll
                for (int c = 0; c < 6; c++) Mx[r][c] = J[r][c < NB ? c : 0];
                Mx[r][6] = vw[r];
            }
            if (solve6(Mx, mot.target_vel) < 1e-13) use_pinv = true; // np.linalg.matrix_rank(J) < 6 (base_robot_arm.py:316-319)
        }
        if (use_pinv) pinv_apply<NB>(J, vw, mot.target_vel);
        if (arm.topo == TG_TOPO_MG400 && NB == 8) {
            // parallelogram emulation (mg400.py:111-120): the three slaved joints follow j2_1 / j3_1
            mot.target_vel[NB - 3] = mot.target_vel[1];
            mot.target_vel[NB - 2] = -mot.target_vel[1];
            mot.target_vel[NB - 1] = mot.target_vel[1] + mot.target_vel[2];
        }
#pragma unroll
        for (int i = 0; i < NB; i++) mot.target_pos[i] = 0.0;
        }
    }
}

// the same from the joint angles (one-thread-per-env kernels): FK, TCP pose, Jacobian, then the core
template <class T, int TASK>
TGD void env_prologue(const TgArm& arm, const TgPhysics& ph, const TgTask& task, const EnvBuffers& b, int e, const float* __restrict__ actions,
                      const double* q, double* v, Motors<T::NB>& mot)
{
    Kin<T::NB> k;
    fk<T>(arm, q, k);
    double tp[3], tq[4], J[6][T::NB];
    tcp_world<T>(arm, k, tp, tq);
    tcp_jacobian<T>(arm, k, tp, J);
    env_prologue_core<T, TASK>(arm, ph, task, b, e, actions, tp, tq, J, v, mot);
}

// Tail of one env step for env e: store the joint state, step data (reward / done), features / oracle vector, then either the
// camera for the raster or - for a finished env under auto-reset - the terminal camera and the swap to the next episode.
template <class T, int TASK>
TGD void env_epilogue_core(const TgArm& arm, const TgPhysics& ph, const TgTask& task, const EnvBuffers& b, int e, const double* q, const double* qd,
                           const double* tp, const double* tq, const double* cam, const ObjState& ob, float* __restrict__ reward,
                           unsigned char* __restrict__ done, int autoreset)
{
    constexpr int NB = T::NB;
    constexpr bool balance = TASK == TG_TASK_OBJECT_BALANCE, roll = TASK == TG_TASK_OBJECT_ROLL, surface = TASK == TG_TASK_SURFACE_FOLLOW;
    constexpr bool push = TASK == TG_TASK_OBJECT_PUSH || roll;
    const int steps = b.steps[e] + 1;
    b.steps[e] = steps;
#pragma unroll
    for (int i = 0; i < NB; i++) { b.q[(size_t)i * b.n + e] = q[i]; b.qd[(size_t)i * b.n + e] = qd[i]; }
    float r; unsigned char d;
    if (roll) {
        const double* tr = b.traj + (size_t)e * PUSH_TRAJ_SZ;
        roll_step_data(task, ob, tp, tq, tr[0], tr[1], steps, &r, &d);
        double* st = b.stim + (size_t)e * 12;
        st[0] = ob.ext_pos[0]; st[9] = ob.pos[0]; st[10] = ob.pos[1]; st[11] = ob.pos[2];
        float* f = (d && autoreset && b.term_feat) ? b.term_feat : b.feat;
        if (f) roll_features(tr, f + (size_t)e * TG_PUSH_NFEAT);
    } else if (push) {
        const double* tr = b.traj + (size_t)e * PUSH_TRAJ_SZ;
        int goal = b.goal[e];
        push_step_data(task, ob, tq, tr, tr[2 * PUSH_NTRAJ], goal, steps, &r, &d);
        b.goal[e] = goal;
        obj_stim(task, ob, b.stim + (size_t)e * 12);
        // the observation is taken after get_step_data: the feature carries the (possibly advanced) goal
        float* f = (d && autoreset && b.term_feat) ? b.term_feat : b.feat;
        if (f) push_features(task, tp, tq, tr, tr[2 * PUSH_NTRAJ], goal, f + (size_t)e * TG_PUSH_NFEAT);
    } else if (balance) {
        balance_step_data(task, ob, b.embed[e], steps, &r, &d);
        obj_stim(task, ob, b.stim + (size_t)e * 12); // the pole moves: the raster needs its pose every step
    } else if (surface) {
        const size_t hb = (size_t)e * 2 + (size_t)b.hf_cur[e];
        double dense;
        surface_step_data(task, b.height + hb * SURF_PTS, b.hf_meta + hb * SURF_META, tp, tq, steps, &r, &d, &dense);
        if (task.sparse_reward) {
            // sparse_reward (surface_follow_auto_env.py:59-73): accum_rew += dense; paid out inside termination_dist of the goal
            const double acc = b.accum[e] + dense;
            b.accum[e] = acc;
            const double* meta = b.hf_meta + hb * SURF_META;
            const double gx = tp[0] - meta[3], gy = tp[1] - meta[4], gz = tp[2] - meta[5];
            r = sqrt(gx * gx + gy * gy + gz * gz) < task.termination_dist ? (float)acc : 0.0f;
        }
        float* f = (d && autoreset && b.term_feat) ? b.term_feat : b.feat;
        if (f) surface_features(task, b.hf_meta + hb * SURF_META, tp, tq, f + (size_t)e * TG_PUSH_NFEAT);
    } else edge_step_data(task, tp, b.edge_ang[e], steps, &r, &d);
    {
        // SURVEY 5 (per-env NaN / inf guard): a diverged env ends its episode here - with auto-reset its next episode starts from
        // the pre-computed standby as usual - instead of carrying non-finite numbers through the batch; counted and flagged
        double chk = 0.0;
#pragma unroll
        for (int i = 0; i < NB; i++) chk += q[i] + qd[i];
        if (push || balance) chk += (ob.pos[0] + ob.pos[1] + ob.pos[2]) + (ob.vel[0] + ob.vel[1] + ob.vel[2]) + (ob.omg[0] + ob.omg[1] + ob.omg[2]) +
                                    (ob.quat[0] + ob.quat[1] + ob.quat[2] + ob.quat[3]);
        if (!isfinite(chk)) { d = 1; r = 0.0f; atomicAdd(b.nan_count, 1); atomicOr(b.error_flag, 4); }
    }
    reward[e] = r; done[e] = d;
    if (b.oracle) {
        float* oo = (d && autoreset) ? b.term_oracle : b.oracle;   // a finished env's last state goes to the terminal buffer
        if (oo) oracle_obs_env<T>(arm, task, b, e, oo + (size_t)e * TG_ORACLE_NOBS);
    }
    if (d && autoreset && b.pipeline) {
        // terminal camera for the terminal observation, then the standby becomes the live state
#pragma unroll
        for (int c = 0; c < 12; c++) { b.term_cam[(size_t)e * 12 + c] = cam[c]; b.term_stim[(size_t)e * 12 + c] = b.stim[(size_t)e * 12 + c]; }
        acquire_standby<T>(arm, ph, task, b, e);
        write_live_features(task, b, e);
        if (surface && task.sparse_reward) surface_accum_start(task, b, e);
        if (b.oracle) oracle_obs_env<T>(arm, task, b, e, b.oracle + (size_t)e * TG_ORACLE_NOBS);
    } else {
#pragma unroll
        for (int c = 0; c < 12; c++) b.cam[(size_t)e * 12 + c] = cam[c];
#pragma unroll
        for (int c = 0; c < 3; c++) b.tcp[(size_t)e * 7 + c] = tp[c];
#pragma unroll
        for (int c = 0; c < 4; c++) b.tcp[(size_t)e * 7 + 3 + c] = tq[c];
    }
    if (push && b.pipeline) standby_work<T>(arm, ph, task, b, e, false); // one quantum of this env's next-episode rebuild, if due
}

// the same from the joint angles (one-thread-per-env kernels)
template <class T, int TASK>
TGD void env_epilogue(const TgArm& arm, const TgPhysics& ph, const TgTask& task, const EnvBuffers& b, int e, const double* q, const double* qd,
                      const ObjState& ob, float* __restrict__ reward, unsigned char* __restrict__ done, int autoreset)
{
    Kin<T::NB> k;
    fk<T>(arm, q, k);
    double tp[3], tq[4], cam[12];
    tcp_world<T>(arm, k, tp, tq);
    write_camera<T>(arm, k, cam);
    env_epilogue_core<T, TASK>(arm, ph, task, b, e, q, qd, tp, tq, cam, ob, reward, done, autoreset);
}

// POSCTL: object_push / object_roll compile their TCP_position_control step as its own kernel (with the blocking-move loop
// behind a run-time branch the velocity-control step of config 4 was measured 6 % slower, same-box A/B); the other tasks
// branch at run time on task.control_mode (their position-control move is one out-of-line call).
template <class T, int TASK, bool POSCTL = false>
__global__ void __launch_bounds__((TASK == TG_TASK_OBJECT_PUSH || TASK == TG_TASK_OBJECT_ROLL) ? PUSH_THREADS : 128)
step_kernel(const __grid_constant__ TgArm arm, const __grid_constant__ TgPhysics ph, const __grid_constant__ TgTask task,
            EnvBuffers b, const float* __restrict__ actions, float* __restrict__ reward, unsigned char* __restrict__ done, int autoreset)
{
    constexpr int NB = T::NB;
    constexpr bool balance = TASK == TG_TASK_OBJECT_BALANCE, roll = TASK == TG_TASK_OBJECT_ROLL, surface = TASK == TG_TASK_SURFACE_FOLLOW;
    constexpr bool push = TASK == TG_TASK_OBJECT_PUSH || roll; // object_roll runs on object_push's contact machinery
    int e;
    int col = 0;       // object_push: this env's column in the block's shared-memory row store
    bool owner = true; // object_push: lanes 0..PUSH_LANES-1 of a warp each step an env; the others only help in the hull scan
    if (push) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        col = warp * PUSH_LANES + (lane & (PUSH_LANES - 1));
        e = blockIdx.x * PUSH_BLOCK + col;
        owner = lane < PUSH_LANES && e < b.n;
    } else {
        if ((int)blockIdx.x >= b.step_blocks) {
            standby_role<T>(arm, ph, task, b, b.step_blocks, false);
            return;
        }
        e = env_index(b);
        if (e < 0) return;
    }
    double q[NB], qd[NB];
#pragma unroll
    for (int i = 0; i < NB; i++) { q[i] = 0.0; qd[i] = 0.0; }
    if (owner) {
#pragma unroll
        for (int i = 0; i < NB; i++) { q[i] = b.q[(size_t)i * b.n + e]; qd[i] = b.qd[(size_t)i * b.n + e]; }
    }

    double v[6];
    Motors<NB> mot;
    mot.mode = 0; mot.kp = 0; mot.kd = ph.vel_gain; mot.max_force = ph.max_force;
    if (owner) env_prologue<T, TASK>(arm, ph, task, b, e, actions, q, v, mot);
    ObjState ob;
    {
        double sc[NB][2]; // (sin q, cos q): exact here, then advanced by the trig identity after every substep
#pragma unroll
        for (int i = 0; i < NB; i++) sincos(q[i], &sc[i][0], &sc[i][1]);
        if (push) {
            if (owner) obj_load(b.obj + (size_t)e * 13, b.obj_ext + (size_t)e * 4, b.grav[e], 0.0, task, ob);
            if (POSCTL) {
                // TCP_position_control with the cube / marble in the world: the blocking move's steps are whole-warp substeps, so
                // the warp loops until its last env has reached its target (or used its pos_max_steps); lanes that are through
                // only help in the hull scan from then on (substep_push returns before touching the state of a non-owner)
                double tpos[3] = {0, 0, 0}, torn[4] = {0, 0, 0, 1};
                if (owner) {
                    mot.mode = 1; mot.kp = ph.pos_gain; mot.kd = ph.vel_gain; mot.max_force = ph.max_force;
                    double wf[3] = {task.workframe_pos[0], task.workframe_pos[1], task.workframe_pos[2]};
                    if (roll) wf[2] = b.obj_ext[(size_t)e * 4 + 1];   // this episode's work frame
                    position_control_target<T>(arm, task, wf, q, v, mot, tpos, torn);
                }
                bool moving = owner;
#pragma unroll 1
                for (int s = 0; s < task.pos_max_steps; s++) {
                    if (!__any_sync(0xffffffffu, moving)) break;
                    const bool reached = moving && position_control_reached<T>(arm, sc, qd, tpos, torn);
                    substep_push<T>(arm, ph, task, b.hull, b.n_hull, q, qd, sc, mot, ob, col, moving);
                    if (reached) moving = false;
                }
            } else {
#pragma unroll 1
                for (int s = 0; s < ph.substeps; s++) substep_push<T>(arm, ph, task, b.hull, b.n_hull, q, qd, sc, mot, ob, col, owner); // whole warp
            }
            if (owner) obj_store(b.obj + (size_t)e * 13, b.obj_ext + (size_t)e * 4, ob);
        } else if (balance) {
            obj_load(b.obj + (size_t)e * 13, b.obj_ext + (size_t)e * 4, b.grav[e], b.embed[e], task, ob);
            if (task.control_mode == 1) position_control_move<T>(arm, ph, task, q, qd, v, &ob);
            else {
#pragma unroll 1
                for (int s = 0; s < ph.substeps; s++) substep_obj<T>(arm, ph, task, q, qd, sc, mot, ob);
            }
            obj_store(b.obj + (size_t)e * 13, b.obj_ext + (size_t)e * 4, ob);
        } else if (task.control_mode == 1) {
            position_control_move<T>(arm, ph, task, q, qd, v, nullptr);
        } else {
#pragma unroll 1
            for (int s = 0; s < ph.substeps; s++) substep<T>(arm, ph, q, qd, sc, mot);
        }
    }

    if (!owner) return;
    env_epilogue<T, TASK>(arm, ph, task, b, e, q, qd, ob, reward, done, autoreset);
}

// explicit reset of the masked envs.  With the pipeline on, an env's standby IS its next episode: take it and
// compute the following one, so draws stay in episode order.
template <class T>
__global__ void __launch_bounds__(128)
reset_kernel(const __grid_constant__ TgArm arm, const __grid_constant__ TgPhysics ph, const __grid_constant__ TgTask task,
             EnvBuffers b, const unsigned char* __restrict__ mask)
{
    const int e = env_index(b);
    if (e < 0) return;
    if (mask && !mask[e]) return;
    EpisodeStart<T::NB> s;
    if (b.pipeline) {
        // the standby IS the next episode (complete it if a rebuild is still in flight); the slot it leaves
        // empty is rebuilt by the standby blocks of the following step launches
        standby_work<T>(arm, ph, task, b, e, true);
        consume_standby<T::NB>(b, e);
    } else {
        reset_env<T>(arm, ph, task, b, e, s);
        store_live<T::NB>(b, e, s);
    }
    write_live_features(task, b, e);
    if (task.task == TG_TASK_SURFACE_FOLLOW && task.sparse_reward) surface_accum_start(task, b, e);
    if (b.oracle) oracle_obs_env<T>(arm, task, b, e, b.oracle + (size_t)e * TG_ORACLE_NOBS);
}

// recompute every missing standby (after tg_set_draws invalidated them)
template <class T>
__global__ void __launch_bounds__(128)
standby_kernel(const __grid_constant__ TgArm arm, const __grid_constant__ TgPhysics ph, const __grid_constant__ TgTask task, EnvBuffers b)
{
    standby_role<T>(arm, ph, task, b, 0, true);
}

// ---------------------------------------------------------------- test hooks
template <class T>
__global__ void test_id_kernel(const __grid_constant__ TgArm arm, const __grid_constant__ TgPhysics ph, int n,
                               const double* q_in, const double* qd_in, double* tau_out)
{
    constexpr int NB = T::NB;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    double q[NB], qd[NB], tau[NB];
    for (int i = 0; i < NB; i++) { q[i] = q_in[e * NB + i]; qd[i] = qd_in[e * NB + i]; }
    Kin<NB> k;
    fk<T>(arm, q, k);
    SpI sp[NB];
    body_inertias<T>(arm, k, sp);
    rnea<T>(ph, k, sp, qd, nullptr, tau);
    for (int i = 0; i < NB; i++) tau_out[e * NB + i] = tau[i];
}

template <class T>
__global__ void test_mass_kernel(const __grid_constant__ TgArm arm, int n, const double* q_in, double* M_out)
{
    constexpr int NB = T::NB;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    double q[NB];
    for (int i = 0; i < NB; i++) q[i] = q_in[e * NB + i];
    Kin<NB> k;
    fk<T>(arm, q, k);
    SpI sp[NB];
    body_inertias<T>(arm, k, sp);
    double M[NB][NB];
    crba<T>(k, sp, M);
#pragma unroll
    for (int i = 0; i < NB; i++)
#pragma unroll
        for (int j = 0; j < NB; j++) M_out[(e * NB + i) * NB + j] = M[i][j];
}

template <class T>
__global__ void test_substep_kernel(const __grid_constant__ TgArm arm, const __grid_constant__ TgPhysics ph, int n, int nsteps,
                                    double* q_io, double* qd_io, const double* target_vel)
{
    constexpr int NB = T::NB;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    double q[NB], qd[NB];
    Motors<NB> mot;
    mot.mode = 0; mot.kp = 0; mot.kd = ph.vel_gain; mot.max_force = ph.max_force;
    for (int i = 0; i < NB; i++) { q[i] = q_io[e * NB + i]; qd[i] = qd_io[e * NB + i]; mot.target_vel[i] = target_vel[e * NB + i]; mot.target_pos[i] = 0; }
    double sc[NB][2];
    for (int i = 0; i < NB; i++) sincos(q[i], &sc[i][0], &sc[i][1]);
#pragma unroll 1
    for (int s = 0; s < nsteps; s++) substep<T>(arm, ph, q, qd, sc, mot);
    for (int i = 0; i < NB; i++) { q_io[e * NB + i] = q[i]; qd_io[e * NB + i] = qd[i]; }
}
