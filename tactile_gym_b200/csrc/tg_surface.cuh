// tg_surface.cuh - surface_follow: OpenSimplex heightfield, grid lookups, reward terms.
//
//   os_perm / os_noise2   <- `OpenSimplex(seed=...)`, `.noise2(x, y)` (base_surface_env.py:311-327,443-448).  The
//                            `opensimplex` package is an unpinned, un-vendored dependency (requirements.txt:4): the
//                            published algorithm is restated (oracle/tg_oracle.c:or_opensimplex_*; pinned there by
//                            the package README's known answer noise2(10, 10) = 0.580279369186297 for seed 1234).
//   surf_index            <- xy_to_surface_idx (base_surface_env.py:273-288): np.digitize over np.linspace bins
//   surface_step_data     <- BaseSurfaceEnv.get_step_data / termination (:631-775) + SurfaceFollowAutoEnv.dense_reward
//                            (surface_follow_auto_env.py:76-94)
//   hf_vertex             <- [EXT] the mesh pybullet draws for createCollisionShape(GEOM_HEIGHTFIELD) (:402-424):
//                            float32 data, centred on the grid centre and on the middle of its height range, x = column
//                            index, y = row index (oracle/oracle.py:heightfield_local_vertices)
#pragma once
#include "tg_dyn.cuh"

#define SURF_N TG_SURF_N
#define SURF_PTS (SURF_N * SURF_N)
#define SURF_META 8 // per episode: zc, dir x, dir y, goal x y z, float32 height min, max

// permutation table of OpenSimplex(seed): LCG shuffle (int64 wrap-around; Python's floor modulo).
// One thread does this in ONE reset quantum, beside the step warps of the launch: the 64-bit signed `%` of the plain restatement
// (~100 emulated instructions each, 3 per entry) made it the longest quantum of surface_follow's rebuild - and with ~5 envs
// starting a rebuild every step, the critical path of every step launch.  floor_mod(s, n) for n <= 256 is taken from the two
// 32-bit halves of the unsigned value instead (u mod n, minus 2^64 mod n when s is negative): 32-bit arithmetic only.
TGD void os_perm(long long seed_in, unsigned char* perm /* [256], global */)
{
    unsigned char source[256];
    unsigned long long seed = (unsigned long long)seed_in;
#pragma unroll 1
    for (int i = 0; i < 256; i++) source[i] = (unsigned char)i;
#pragma unroll 1
    for (int k = 0; k < 3; k++) seed = seed * 6364136223846793005ULL + 1442695040888963407ULL;
#pragma unroll 1
    for (int i = 255; i >= 0; i--) {
        seed = seed * 6364136223846793005ULL + 1442695040888963407ULL;
        const unsigned n = (unsigned)(i + 1);
        const unsigned hi = (unsigned)(seed >> 32), lo = (unsigned)seed;
        const unsigned p32 = (0xffffffffu % n + 1u) % n;                     // 2^32 mod n
        unsigned r = ((hi % n) * p32 + lo % n) % n;                          // u mod n
        if ((long long)seed < 0) r = (r + n - (p32 * p32) % n) % n;          // (u - 2^64) mod n, Python's sign convention
        r = (r + 31u % n) % n;
        perm[i] = source[r];
        source[r] = source[i];
    }
}

TGD double os_extrapolate2(const unsigned char* perm, long long xsb, long long ysb, double dx, double dy)
{
    const int index = perm[(perm[xsb & 0xFF] + ysb) & 0xFF] & 0x0E;
    // gradients (5,2) (2,5) (-5,2) (-2,5) (5,-2) (2,-5) (-5,-2) (-2,-5), index = 2 * gradient number
    const int g = index >> 1;
    const double a = (g & 1) ? 2.0 : 5.0, bb = (g & 1) ? 5.0 : 2.0;
    const double g1 = (g & 2) ? -a : a, g2 = (g & 4) ? -bb : bb;
    return g1 * dx + g2 * dy;
}

TGD double os_noise2(const unsigned char* perm, double x, double y)
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
    double value = 0.0, dx_ext, dy_ext;
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

// np.linspace(c - 32 g, c + 32 g, 64)[k] (base_surface_env.py:264-271): k * step + start, the last one = stop
TGD double surf_bin(double centre, double grid, int k)
{
    const double start = centre - (SURF_N / 2) * grid, stop = centre + (SURF_N / 2) * grid;
    const double step = (stop - start) / (double)(SURF_N - 1);
    return k == SURF_N - 1 ? stop : (double)k * step + start;
}
// np.digitize(v, bins) (increasing bins, right=False) = number of bins <= v, then the reference's clamp of 64 -> 63
TGD int surf_digitize(double centre, double grid, double v)
{
    const double start = centre - (SURF_N / 2) * grid, stop = centre + (SURF_N / 2) * grid;
    const double step = (stop - start) / (double)(SURF_N - 1);
    int k = (int)fmin(fmax(floor((v - start) / step) + 1.0, 0.0), (double)SURF_N);
    while (k < SURF_N && surf_bin(centre, grid, k) <= v) k++;
    while (k > 0 && surf_bin(centre, grid, k - 1) > v) k--;
    return k == SURF_N ? SURF_N - 1 : k;
}
// xy_to_surface_idx: i from y, j from x
TGD void surf_index(const TgTask& task, double x, double y, int& i, int& j)
{
    i = surf_digitize(task.surf_pos[1], task.surf_grid, y);
    j = surf_digitize(task.surf_pos[0], task.surf_grid, x);
}

// np.gradient(h, grid) at (i, j): central differences inside, one-sided first order on the border
TGD void surf_gradient(const double* H, double grid, int i, int j, double& d_axis0, double& d_axis1)
{
    auto at = [&](int r, int c) { return H[r * SURF_N + c]; };
    d_axis0 = i == 0 ? (at(1, j) - at(0, j)) / grid : i == SURF_N - 1 ? (at(SURF_N - 1, j) - at(SURF_N - 2, j)) / grid : (at(i + 1, j) - at(i - 1, j)) / (2.0 * grid);
    d_axis1 = j == 0 ? (at(i, 1) - at(i, 0)) / grid : j == SURF_N - 1 ? (at(i, SURF_N - 1) - at(i, SURF_N - 2)) / grid : (at(i, j + 1) - at(i, j - 1)) / (2.0 * grid);
}

// reward / termination.  H: the live heightfield [64][64] (row = y index), meta: zc, dir, goal
TGD void surface_step_data(const TgTask& task, const double* H, const double* meta, const double* tp, const double* tq, int steps,
                           float* reward, unsigned char* done, double* dense = nullptr)
{
    int ti, tj;
    surf_index(task, tp[0], tp[1], ti, tj);
    double R[9];
    mat_from_quat(tq, R);
    const double gx = tp[0] - meta[3], gy = tp[1] - meta[4], gz = tp[2] - meta[5];
    const double goal_dist = sqrt(gx * gx + gy * gy + gz * gz);
    // z_dist_to_surface (:727-757): tip pushed embed_dist along its own -z
    const double emb_z = tp[2] + R[8] * (-task.surf_embed);
    const double surf_z = H[ti * SURF_N + tj] + task.surf_pos[2];
    double surf_dist = fabs(emb_z - surf_z);
    if (task.surf_vertical) surf_dist = fabs((tp[0] + R[0] * (-task.surf_embed)) - (task.surf_pos[0] - H[ti * SURF_N + tj]));
    // cos_dist_to_surface_normal (:701-725)
    double g0, g1;
    surf_gradient(H, task.surf_grid, ti, tj, g0, g1); // g0 = "surface_grad_y", g1 = "surface_grad_x"
    double n[3] = {-g1, -g0, 1.0};
    const double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    n[0] /= nn; n[1] /= nn; n[2] /= nn;
    double v[3] = {-R[2], -R[5], -R[8]};
    if (task.surf_vertical) {
        // the upright surface (:708-711, :733-735, :752-754): the forward sensor's axis is its -x, the surface distance is taken
        // along world x, the normal is the flipped one (local (x, y, z) -> (-z, y, x))
        const double fl[3] = {-n[2], n[1], n[0]};
        n[0] = fl[0]; n[1] = fl[1]; n[2] = fl[2];
        v[0] = -R[0]; v[1] = -R[3]; v[2] = -R[6];
    }
    const double cs = (n[0] * v[0] + n[1] * v[1] + n[2] * v[2]) / (sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]) * sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]));
    // SurfaceFollowAutoEnv.dense_reward (W_goal = 0, W_surf = 1) / SurfaceFollowGoalEnv.dense_reward (W_goal = 1, W_surf = 10)
    const double goal_xy = sqrt(gx * gx + gy * gy);
    const double rw = -((task.surf_w_goal * goal_xy) + (task.surf_w_surf * surf_dist) + (task.surf_w_norm * (1.0 - cs)));
    if (dense) *dense = rw;
    *reward = (float)rw;
    *done = (goal_dist < task.termination_dist || steps >= task.max_steps) ? 1 : 0;
}

// SurfaceFollowGoalEnv.get_extended_feature_array (surface_follow_goal_env.py:83-97): TCP and goal position, work frame
TGD void surface_features(const TgTask& task, const double* meta, const double* tp, const double* tq, float* out)
{
    double wp[3], wr[3], gp[3], gr[3];
    const double ident[4] = {0.0, 0.0, 0.0, 1.0}, g[3] = {meta[3], meta[4], meta[5]};
    world_to_work(task, tp, tq, wp, wr);
    world_to_work(task, g, ident, gp, gr);
#pragma unroll
    for (int i = 0; i < 3; i++) { out[i] = (float)wp[i]; out[3 + i] = (float)gp[i]; }
#pragma unroll
    for (int i = 6; i < TG_PUSH_NFEAT; i++) out[i] = 0.0f;
}

// world-space mesh vertex (column j = x index, row i = y index); zc = middle of the float32 height range
TGD void hf_vertex(const double* surf_pos, double grid, const double* H, double zc, int i, int j, double* v)
{
    v[0] = surf_pos[0] + (double)(float)(((double)j - (SURF_N - 1) / 2.0) * grid);
    v[1] = surf_pos[1] + (double)(float)(((double)i - (SURF_N - 1) / 2.0) * grid);
    v[2] = surf_pos[2] + (double)(float)((double)(float)H[i * SURF_N + j] - zc);
}
