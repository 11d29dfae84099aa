// tg_raster.cuh - tactile depth raster + post-process, one uint8 [S][S] image per env.
//
// Replaces TactileSensor.get_imgs (pb.getCameraImage over the whole ~270k-triangle scene,
// sensors/tactile_sensor.py:212-259) + TactileSensor.t_s_camera (:261-294).
//
// What makes it cheap ("decisive simplification"): the camera is rigidly attached to the
// sensor, so everything but the stimulus is static in the camera frame and already baked into the
// reference's nodef_dep / border_mask / nodef_gray images.  Per env only the few stimulus primitives are
// z-tested against nodef_dep.  Primitives are convex planar polygons with 3 or 4 vertices (coplanar triangle
// pairs of the stimulus mesh are merged into quads at scene-compile time: same coverage, same plane).
//
// Kernel shape (HBM-write bound; algorithmic bytes/env = S*S obs + 192 B camera/stimulus state):
//   * persistent CTAs, grid = #SMs; each CTA owns one row band, whose slice of nodef_dep (f32) and of the
//     pre-baked border image (u8) is fetched ONCE per CTA by TMA bulk copies (cp.async.bulk + mbarrier) into
//     shared memory and reused for every env; after that there is no block-level synchronisation: each WARP
//     renders whole env images (band slices) on its own;
//   * per env, one lane per primitive builds homogeneous edge equations in pixel coordinates (inside <=> all
//     E_i >= 0; 1/z_eye is affine in the pixel, hence so is the un-quantised output value
//     val = 5100 (nodef - d), d = F - F near / z): fp64 coefficients + float copies with an error margin;
//   * BIN: every 8 x 64 tile is classified per primitive from its corner pixels (out / in / partial, bbox
//     reject, exact occlusion cull); the baked row is copied to HBM; 16-pixel spans that a primitive really
//     touches and that contain skin pixels are compacted into a small per-warp list;
//   * SHADE: 32 listed spans at a time, one per lane - all lanes busy.  Covered pixels take
//     max over primitives of an affine float function, one FFMA with nodef, clamp, truncate;
//   * EXACT PATCH-UP: the float path carries a proven error bound (1e-3 of an output LSB); pixels whose value
//     is closer than that to a quantisation step, or closer than the float margin to a primitive edge, are
//     queued and recomputed with the oracle's arithmetic (fp64 depth -> float32 numpy post-process), so the
//     bytes equal the CPU oracle's.  About 0.3 % of the covered pixels take this path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "../../include/tactile_gym_b200.h"

#define RASTER_THREADS 512
#define RASTER_WARPS (RASTER_THREADS / 32)
#define TILE_ROWS 16 // 16 x 32 pixel tiles = 32 spans of 16 px, one per lane (squarish: an oblique edge crosses fewer of them)
#define TILE_COLS 32
#define TILE_SPR_SHIFT 1 // log2(spans per tile row)
#define RASTER_MAXPRIM 12
#define SPAN_LIST 64   // per-warp compacted span lists (entries each; there are two)
#define EXACT_QUEUE 640 // per-warp queue of pixels for the exact path (a shade batch adds <= 512; flushed above 128)
#define VAL_SCALE 5100.0f // 255 / 0.05
#define VAL_BOUND 1.0e-3f // proven bound on |float value - oracle value|

struct PrimCoef {
    double eA[5], eB[5], eC[5]; // fp64: E_i(c, r) = eA[i] c + eB[i] r + eC[i], i < 4 edges; index 4 = w = 1/z_eye
    float fA[5], fB[5], fC[5];  // float copies
    float vA, vB, vC;           // float: V(c, r) = 5100 (nd_ref - F + F near w(c, r))  -> val = 5100 (nodef - nd_ref) + V
    float margin;               // bound on the float evaluation error of the edge functions
    float wmargin;              // ... of w = 1/z_eye
    int steep;                  // 1: V varies too fast for the float value bound -> V is evaluated in fp64
    double dvA, dvB, dvC;       // fp64 V coefficients (used for steep primitives)
    float c_lo, c_hi, r_lo, r_hi; // conservative screen bbox in pixel units
    int valid, clipped;         // clipped: a vertex is outside [near, far] -> tiles check the 1/z range of the plane
    uint32_t behind;            // primitives whose plane this one lies entirely behind (or on: ties go to the lower index)
};

struct RasterArgs {
    int n, S, bands, nprim;
    double th;             // tan(fov/2)
    double F, near_, far_; // F = far/(far-near)
    const float* nodef;    // [S*S], border pixels = -1
    const uint8_t* base;   // [S*S], border pixels = (u8)nodef_gray, others 0
    const double* prims;   // [nprim][4][3] stimulus-frame polygons
    const int* prim_nv;    // [nprim] 3 or 4
    const double* cam;     // [N][12]
    const double* stim;    // [N][12]
    const uint8_t* mask;   // optional [N]
    uint8_t* obs;          // [N][S*S]
    float nd_ref;          // reference depth: the float path works on (nodef - nd_ref) to keep magnitudes (and errors) small
    // heightfield stimulus (surface_follow, raster_hf_kernel): per-env 64 x 64 heights instead of a shared primitive list
    const double* hf;      // [N][2][64*64] fp64 heights (row = y index)
    const int* hf_cur;     // [N] which of the two is the live episode's
    const double* hf_meta; // [N][2][SURF_META], [0] = the mesh's height offset (middle of the float32 range)
    int hf_flip;           // 1: use the OTHER buffer (terminal observation of an env that has just been reset)
    int hf_tile_rows, hf_tile_cols; // tile of the heightfield kernel: rows x cols pixels, rows * cols / 16 <= 32 spans
    double surf_pos[3], surf_grid;
    int scan_test_fallback; // test hook (TG_SCAN_TEST_FALLBACK): the scanline raster flags every odd env for the general kernel
};

struct SpanEntry {
    uint16_t off16;  // span index inside the band (pixel offset / 16)
    uint16_t in_m;   // primitives covering the whole span
    uint16_t part_m; // primitives crossing it
    uint16_t pad;
};

#define RASTER_PER_WARP_SMEM (sizeof(PrimCoef) * RASTER_MAXPRIM + sizeof(SpanEntry) * SPAN_LIST * 2 + sizeof(uint32_t) * EXACT_QUEUE + 16)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :
                 : "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// vp[k] = (column, row, 1/z) of vertex k in pixel units, meaningful when the return value (all vertices in front) is true
__device__ __forceinline__ bool prim_from_eye(const RasterArgs& a, const double (*ve)[3], int nv, PrimCoef& o, double (*vp)[3]);

__device__ __forceinline__ bool prim_setup(const RasterArgs& a, const double* cam, const double* stim, const double* pl, int nv, PrimCoef& o, double (*vp)[3])
{
    // stimulus frame -> world -> eye space (x right, y up, z forward)
    double ve[4][3];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const double* v = pl + 3 * (k < nv ? k : nv - 1);
        double w[3];
#pragma unroll
        for (int c = 0; c < 3; c++) w[c] = stim[3 * c] * v[0] + stim[3 * c + 1] * v[1] + stim[3 * c + 2] * v[2] + stim[9 + c] - cam[c];
        ve[k][0] = w[0] * cam[9] + w[1] * cam[10] + w[2] * cam[11];
        ve[k][1] = w[0] * cam[6] + w[1] * cam[7] + w[2] * cam[8];
        ve[k][2] = w[0] * cam[3] + w[1] * cam[4] + w[2] * cam[5];
    }
    return prim_from_eye(a, ve, nv, o, vp);
}

// world -> eye space
__device__ __forceinline__ void world_to_eye(const double* cam, const double* v, double* e)
{
    const double w[3] = {v[0] - cam[0], v[1] - cam[1], v[2] - cam[2]};
    e[0] = w[0] * cam[9] + w[1] * cam[10] + w[2] * cam[11];
    e[1] = w[0] * cam[6] + w[1] * cam[7] + w[2] * cam[8];
    e[2] = w[0] * cam[3] + w[1] * cam[4] + w[2] * cam[5];
}

// screen-space coefficient setup of one convex planar polygon given in eye space (ve[k], k >= nv repeat the last vertex)
__device__ __forceinline__ bool prim_from_eye(const RasterArgs& a, const double (*ve)[3], int nv, PrimCoef& o, double (*vp)[3])
{
    const double S = a.S;
    // plane: n . p = n . v0 ; along the pixel ray p = z d (d.z = 1):  1/z = (n . d) / (n . v0)
    double e1[3] = {ve[1][0] - ve[0][0], ve[1][1] - ve[0][1], ve[1][2] - ve[0][2]};
    double e2[3] = {ve[2][0] - ve[0][0], ve[2][1] - ve[0][1], ve[2][2] - ve[0][2]};
    double nrm[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    const double nd0 = nrm[0] * ve[0][0] + nrm[1] * ve[0][1] + nrm[2] * ve[0][2];
    const double nn = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
    o.valid = 0; o.behind = 0;
    if (!(fabs(nd0) > 1e-12 * nn * (fabs(ve[0][0]) + fabs(ve[0][1]) + fabs(ve[0][2]) + 1e-300))) return false; // edge-on or degenerate
    o.valid = 1;
    // d = (dx, dy, 1), dx = th ((2c+1)/S - 1), dy = th (1 - (2r+1)/S): a row vector g . d becomes A c + B r + C
    const double kx = a.th * 2.0 / S, x0 = a.th * (1.0 / S - 1.0), y0 = a.th * (1.0 - 1.0 / S);
    auto affine = [&](const double* g, double scale, int i) {
        o.eA[i] = g[0] * scale * kx; o.eB[i] = -g[1] * scale * kx; o.eC[i] = (g[0] * x0 + g[1] * y0 + g[2]) * scale;
    };
    affine(nrm, 1.0 / nd0, 4);
    double cen[3] = {0, 0, 0};
    for (int k = 0; k < nv; k++) { cen[0] += ve[k][0]; cen[1] += ve[k][1]; cen[2] += ve[k][2]; }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (i < nv) {
            const double* p = ve[i];
            const double* q = ve[i + 1 < nv ? i + 1 : 0];
            double g[3] = {p[1] * q[2] - p[2] * q[1], p[2] * q[0] - p[0] * q[2], p[0] * q[1] - p[1] * q[0]};
            const double sgn = (g[0] * cen[0] + g[1] * cen[1] + g[2] * cen[2]) >= 0 ? 1.0 : -1.0;
            const double gl = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
            affine(g, gl > 0 ? sgn / gl : 0.0, i);
        } else { o.eA[i] = 0; o.eB[i] = 0; o.eC[i] = 1.0; } // unused edge slot: always inside
    }
    float mg = 0.f;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        o.fA[i] = (float)o.eA[i]; o.fB[i] = (float)o.eB[i]; o.fC[i] = (float)o.eC[i];
        if (i < nv) mg = fmaxf(mg, (float)((fabs(o.eA[i]) + fabs(o.eB[i])) * S + fabs(o.eC[i])));
    }
    o.margin = mg * 2e-6f;
    o.wmargin = (float)((fabs(o.eA[4]) + fabs(o.eB[4])) * S + fabs(o.eC[4])) * 2e-6f;
    const double Fn = a.F * a.near_;
    const double dvA = 5100.0 * Fn * o.eA[4], dvB = 5100.0 * Fn * o.eB[4], dvC = 5100.0 * (Fn * o.eC[4] - a.F + (double)a.nd_ref);
    o.vA = (float)dvA; o.vB = (float)dvB; o.vC = (float)dvC;
    o.dvA = dvA; o.dvB = dvB; o.dvC = dvC;
    // float evaluation of V: ~4 roundings at the magnitude of its terms; it must stay well inside VAL_BOUND
    o.steep = ((fabs(dvA) + fabs(dvB)) * S + fabs(dvC)) * 2.4e-7 > 0.5 * VAL_BOUND;
    bool front = true, clipped = false;
    for (int k = 0; k < nv; k++) {
        if (!(ve[k][2] > 1e-6)) front = false;
        if (!(ve[k][2] >= a.near_) || ve[k][2] > a.far_) clipped = true;
    }
    o.clipped = clipped;
    o.c_lo = 0.f; o.c_hi = (float)(S - 1); o.r_lo = 0.f; o.r_hi = (float)(S - 1);
    if (front) {
        double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
        for (int k = 0; k < nv; k++) {
            const double x = ve[k][0] / (ve[k][2] * a.th), y = ve[k][1] / (ve[k][2] * a.th);
            xmin = fmin(xmin, x); xmax = fmax(xmax, x); ymin = fmin(ymin, y); ymax = fmax(ymax, y);
            vp[k][0] = (x + 1) * 0.5 * S - 0.5; vp[k][1] = (1 - y) * 0.5 * S - 0.5; vp[k][2] = 1.0 / ve[k][2];
        }
        o.c_lo = (float)((xmin + 1) * 0.5 * S - 0.5 - 1.0); o.c_hi = (float)((xmax + 1) * 0.5 * S - 0.5 + 1.0);
        o.r_lo = (float)((1 - ymax) * 0.5 * S - 0.5 - 1.0); o.r_hi = (float)((1 - ymin) * 0.5 * S - 0.5 + 1.0);
    }
    return front;
}

// optional counters for tools/raster_stats.py (a separate diagnostic build; never in the product library)
#ifdef TG_RASTER_STATS
__device__ unsigned long long g_rstats[48];
#define RSTAT(i, v) atomicAdd(&g_rstats[i], (unsigned long long)(v))
#else
#define RSTAT(i, v) ((void)0)
#endif

// exact (fp64) coverage: all edge functions >= -1e-12 (they are normalised to unit gradient), in front of the eye
__device__ __forceinline__ bool inside_exact(const PrimCoef& t, int c, int r, double& w)
{
    w = t.eA[4] * c + t.eB[4] * r + t.eC[4];
    bool in = w > 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) in = in && (t.eA[i] * c + t.eB[i] * r + t.eC[i] >= -1e-12);
    return in;
}

// t_s_camera's float32 arithmetic (tactile_sensor.py:268-284).  uint8(((clip(pen, 0, 0.05) / 0.05) * 255)):
// q = pen / 0.05f is computed as q0 = pen * 20, q = fma(fma(-0.05f, q0, pen), 20, q0); over all 1.03e9 floats
// in [0, 0.05f] the resulting uint8 equals the IEEE-division one (tools/check_quantize.c, tests/test_host.py).
__device__ __forceinline__ uint32_t quantize(float cur, float nd)
{
    float diff = cur - nd;
    const float eps = 1e-4f, maxpen = 0.05f, rcp = 20.0f;
    if (diff >= -eps && diff <= eps) diff = 0.0f;
    const float pen = fminf(fabsf(diff), maxpen);
    const float q0 = __fmul_rn(pen, rcp);
    const float q = __fmaf_rn(__fmaf_rn(-maxpen, q0, pen), rcp, q0);
    return (uint32_t)__float2uint_rz(__fmul_rn(q, 255.0f));
}

// the oracle's arithmetic for one pixel: nearest covering primitive in fp64, GL near/far clipping, float32 post-process
__device__ __forceinline__ uint32_t exact_pixel(const RasterArgs& a, const PrimCoef* pc, uint32_t cand, int c, int r, float nd, uint32_t basev)
{
    if (nd < 0.0f) return basev;
    const double w_near = 1.0 / a.near_, w_far = 1.0 / a.far_;
    double best = 0.0;
    while (cand) { // primitives that can touch this pixel's span (tile + span classification are conservative)
        const int t = __ffs(cand) - 1;
        cand &= cand - 1;
        double w;
        if (!inside_exact(pc[t], c, r, w)) continue;
        if (w <= w_near && w >= w_far && w > best) best = w;
    }
    if (best <= 0.0) return 0u;
    const float d = (float)(a.F - a.F * a.near_ * best);
    return quantize(fminf(nd, d), nd);
}

// what a warp's shading code needs (all shared-memory pointers are the warp's own)
struct WarpCtx {
    const PrimCoef* pc;
    const float* s_nodef;
    const uint8_t* s_base;
    void* s_queue;  // exact-path queue: uint32 entries (mask << 16 | pixel offset, <= 16 primitives) or uint64 (mask << 32 | offset)
    int* s_qcnt;
    int S, sh_S, row0;
};

template <class QT> __device__ __forceinline__ QT queue_entry(uint32_t cand, uint32_t off);
template <> __device__ __forceinline__ uint32_t queue_entry<uint32_t>(uint32_t cand, uint32_t off) { return (cand << 16) | off; }
template <> __device__ __forceinline__ unsigned long long queue_entry<unsigned long long>(uint32_t cand, uint32_t off) { return ((unsigned long long)cand << 32) | off; }
__device__ __forceinline__ void queue_decode(uint32_t q, uint32_t& cand, int& off) { cand = q >> 16; off = (int)(q & 0xffffu); }
__device__ __forceinline__ void queue_decode(unsigned long long q, uint32_t& cand, int& off) { cand = (uint32_t)(q >> 32); off = (int)(uint32_t)q; }

// shade one 16-pixel span (this lane's): float fast path; uncertain pixels go to the exact queue.
// off = pixel offset inside the band; in_m / part_m = primitives covering the whole span / crossing it
template <bool WITH_PART, class QT>
__device__ __forceinline__ void shade_span(const RasterArgs& a, const WarpCtx& x, int off, uint32_t in_m, uint32_t part_m, bool clip, uint8_t* obs_e)
{
    const PrimCoef* pc = x.pc;
    const int S = x.S;
    const int lr = off >> x.sh_S, c0 = off & (S - 1), r = x.row0 + lr;
    const float fr = (float)r, fc0 = (float)c0;
    float vb[16];
#pragma unroll
    for (int k = 0; k < 16; k++) vb[k] = -1e30f;
    uint32_t unc = 0; // pixels that need the exact path
    uint32_t m = in_m;
    while (m) {
        const int t = __ffs(m) - 1;
        m &= m - 1;
        if (pc[t].steep) {
            // grazing primitive: V is a small difference of large terms -> evaluate it in fp64, round the result
            const double dA = pc[t].dvA, d0 = dA * c0 + (pc[t].dvB * r + pc[t].dvC);
#pragma unroll
            for (int k = 0; k < 16; k++) vb[k] = fmaxf(vb[k], (float)(dA * k + d0));
            continue;
        }
        const float vA = pc[t].vA, v0 = fmaf(vA, fc0, fmaf(pc[t].vB, fr, pc[t].vC));
#pragma unroll
        for (int k = 0; k < 16; k++) vb[k] = fmaxf(vb[k], fmaf(vA, (float)k, v0));
    }
    if constexpr (WITH_PART) {
        m = part_m;
        while (m) {
            const int t = __ffs(m) - 1;
            m &= m - 1;
            const PrimCoef& c = pc[t];
            const float mg = c.margin, sc = c.margin / fmaxf(c.wmargin, 1e-30f); // w is compared on the edges' margin scale
            const bool steep = c.steep;
            float a0[5];
#pragma unroll
            for (int i = 0; i < 5; i++) a0[i] = fmaf(c.fA[i], fc0, fmaf(c.fB[i], fr, c.fC[i]));
            const float vA = c.vA, v0 = fmaf(vA, fc0, fmaf(c.vB, fr, c.vC));
            const double dA = c.dvA, d0 = steep ? dA * c0 + (c.dvB * r + c.dvC) : 0.0;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const float fk = (float)k;
                const float lo = fminf(fminf(fmaf(c.fA[0], fk, a0[0]), fmaf(c.fA[1], fk, a0[1])),
                                       fminf(fminf(fmaf(c.fA[2], fk, a0[2]), fmaf(c.fA[3], fk, a0[3])), fmaf(c.fA[4], fk, a0[4]) * sc));
                if (lo > mg) vb[k] = fmaxf(vb[k], steep ? (float)(dA * k + d0) : fmaf(vA, fk, v0));
                else if (lo >= -mg) unc |= 1u << k; // within the float margin of an edge
            }
            RSTAT(4, __popc(unc));
        }
    }
    float nd[16];
    {
        const float4* p = reinterpret_cast<const float4*>(x.s_nodef + off);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float4 v = p[k];
            nd[4 * k] = v.x; nd[4 * k + 1] = v.y; nd[4 * k + 2] = v.z; nd[4 * k + 3] = v.w;
        }
    }
    const uint4 bres = *reinterpret_cast<const uint4*>(x.s_base + off);
    uint32_t wds[4] = {bres.x, bres.y, bres.z, bres.w};
    uint32_t nd_skin = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        // val = 5100 (nodef - d).  nodef - nd_ref is exact (Sterbenz); uncovered pixels have vb = -1e30 -> 0;
        // border pixels (nd = -1) keep the baked byte
        const bool skin = nd[k] >= 0.0f;
        nd_skin |= (skin ? 1u : 0u) << k;
        const float val = skin ? fmaf(nd[k] - a.nd_ref, VAL_SCALE, vb[k]) : -1e30f;
        const uint32_t u = __float2uint_rz(fminf(fmaxf(val, 0.0f), 255.0f));
        // quantisation step within the error bound?  (val in (-B, 255 + B) and |val - round(val)| < B)
        if (fabsf(val - rintf(val)) < VAL_BOUND && val > -VAL_BOUND && val < 255.0f + VAL_BOUND) unc |= 1u << k;
        wds[k >> 2] |= u << (8 * (k & 3));
    }
    *reinterpret_cast<uint4*>(obs_e + off) = make_uint4(wds[0], wds[1], wds[2], wds[3]);
    if (clip) unc = 0xffffu; // near/far clipping in play in this tile: every pixel of the span takes the exact path
    unc &= nd_skin;
    RSTAT(5, __popc(unc)); RSTAT(6, __popc(unc) * __popc(in_m | part_m)); RSTAT(7, 1);
    RSTAT(8, __popc(part_m));
    QT* queue = reinterpret_cast<QT*>(x.s_queue);
    while (unc) {
        const int k = __ffs(unc) - 1;
        unc &= unc - 1;
        queue[atomicAdd(x.s_qcnt, 1)] = queue_entry<QT>(in_m | part_m, (uint32_t)(off + k));
    }
}

// EXACT PATCH-UP of the queued pixels (whole warp)
template <class QT>
__device__ __forceinline__ void flush_exact(const RasterArgs& a, const WarpCtx& x, uint8_t* obs_e, int lane)
{
    __syncwarp();
    const int qn = *x.s_qcnt;
    const QT* queue = reinterpret_cast<const QT*>(x.s_queue);
    for (int i = lane; i < qn; i += 32) {
        uint32_t cand; int off;
        queue_decode(queue[i], cand, off);
        obs_e[off] = (uint8_t)exact_pixel(a, x.pc, cand, off & (x.S - 1), x.row0 + (off >> x.sh_S), x.s_nodef[off], x.s_base[off]);
    }
    __syncwarp();
    if (lane == 0) *x.s_qcnt = 0;
    __syncwarp();
}

// `need`: optional device counter; a launch that finds it zero has nothing to render and returns at once (the scanline
// raster's fallback pass, tg_raster_scan.cuh)
__global__ void __launch_bounds__(RASTER_THREADS)
raster_kernel(const RasterArgs a, const int* __restrict__ need)
{
    if (need && *need == 0) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = a.S, band_rows = S / a.bands, band_px = band_rows * S;
    float* s_nodef = reinterpret_cast<float*>(smem_raw);
    uint8_t* s_base = smem_raw + (size_t)band_px * 4;
    uint32_t* s_skin = reinterpret_cast<uint32_t*>(smem_raw + (size_t)band_px * 5); // 1 bit per 16-px span: has a non-border pixel
    const int n_spans = band_px / 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t wbase = ((size_t)band_px * 5 + (size_t)((n_spans + 31) / 32) * 4 + 15) & ~size_t(15);
    const size_t per_warp = RASTER_PER_WARP_SMEM;
    PrimCoef* pc = reinterpret_cast<PrimCoef*>(smem_raw + wbase + per_warp * warp);
    SpanEntry* s_list = reinterpret_cast<SpanEntry*>(pc + RASTER_MAXPRIM);
    uint32_t* s_queue = reinterpret_cast<uint32_t*>(s_list + 2 * SPAN_LIST);
    int* s_qcnt = reinterpret_cast<int*>(s_queue + EXACT_QUEUE);
    const int tiles_x = S / TILE_COLS, tiles_y = band_rows / TILE_ROWS, n_tiles = tiles_x * tiles_y; // <= 32
    __shared__ __align__(8) uint64_t bar;

    const int sh_b = 31 - __clz(a.bands); // bands is a power of two
    const int band = blockIdx.x & (a.bands - 1);
    const int lane_cta = blockIdx.x >> sh_b, n_cta = gridDim.x >> sh_b;
    const int row0 = band * band_rows;

    // TMA bulk copies of this band's tables, once per CTA
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        const uint32_t bytes = (uint32_t)band_px * 5u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        tma_bulk_load(s_nodef, a.nodef + (size_t)row0 * S, (uint32_t)band_px * 4u, &bar);
        tma_bulk_load(s_base, a.base + (size_t)row0 * S, (uint32_t)band_px, &bar);
    }
    __syncthreads();
    {
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(ok)
                         : "r"(smem_u32(&bar)), "r"(0u)
                         : "memory");
        }
    }
    // span skin bitmap (once per CTA)
    for (int w0 = threadIdx.x; w0 < (n_spans + 31) / 32; w0 += RASTER_THREADS) {
        uint32_t bits = 0;
        for (int j = 0; j < 32 && w0 * 32 + j < n_spans; j++) {
            const float* p = s_nodef + (size_t)(w0 * 32 + j) * 16;
            bool skin = false;
            for (int k = 0; k < 16; k++) skin = skin || (p[k] >= 0.0f);
            bits |= (skin ? 1u : 0u) << j;
        }
        s_skin[w0] = bits;
    }
    __syncthreads();

    const uint32_t lt_mask = (1u << lane) - 1u;
    const int sh_S = 31 - __clz(S), sh_tx = 31 - __clz(tiles_x); // S and tiles_x are powers of two
    SpanEntry* l_in = s_list;             // spans covered by whole primitives only
    SpanEntry* l_pt = s_list + SPAN_LIST; // spans some primitive edge crosses

    WarpCtx ctx;
    ctx.pc = pc; ctx.s_nodef = s_nodef; ctx.s_base = s_base; ctx.s_queue = s_queue; ctx.s_qcnt = s_qcnt;
    ctx.S = S; ctx.sh_S = sh_S; ctx.row0 = row0;
    auto shade = [&](const SpanEntry en, uint8_t* obs_e, auto with_part) {
        shade_span<decltype(with_part)::value, uint32_t>(a, ctx, (int)en.off16 * 16, en.in_m, en.part_m, en.pad != 0, obs_e);
    };
    auto flush_queue = [&](uint8_t* obs_e) { flush_exact<uint32_t>(a, ctx, obs_e, lane); };

    // one env image (band slice) per warp iteration
    for (int e = lane_cta * RASTER_WARPS + warp; e < a.n; e += n_cta * RASTER_WARPS) {
        if (a.mask && !a.mask[e]) continue;
        __syncwarp();
        if (lane == 0) *s_qcnt = 0;
        {
            // per-primitive setup (lane = primitive), then the pairwise "entirely behind the other's plane" relation:
            // 1/z of a plane is affine on the screen, so plane d is in front of polygon t everywhere iff it is at t's
            // vertices.  Wherever d covers a whole tile, t cannot be seen there (grazing slivers, coplanar faces).
            double vp[4][3];
            bool front = false;
            const int nv = lane < a.nprim ? a.prim_nv[lane] : 0;
            if (lane < a.nprim) front = prim_setup(a, a.cam + (size_t)e * 12, a.stim + (size_t)e * 12, a.prims + 12 * lane, nv, pc[lane], vp);
            __syncwarp();
            uint32_t bh = 0;
            if (lane < a.nprim && front && pc[lane].valid) {
                for (int d = 0; d < a.nprim; d++) {
                    const PrimCoef& o = pc[d];
                    if (d == lane || !o.valid) continue;
                    bool ok = true;
                    for (int k = 0; k < nv; k++) {
                        const double wd = o.eA[4] * vp[k][0] + o.eB[4] * vp[k][1] + o.eC[4];
                        const double tol = 1e-11 * (fabs(o.eA[4] * vp[k][0]) + fabs(o.eB[4] * vp[k][1]) + fabs(o.eC[4]) + vp[k][2]);
                        ok = ok && (wd - vp[k][2] >= -tol);
                    }
                    if (ok) bh |= 1u << d;
                }
            }
            // coplanar pairs are behind each other: the lower index stays
            if (lane < a.nprim) pc[lane].behind = bh;
            __syncwarp();
            uint32_t m = bh;
            while (m) {
                const int d = __ffs(m) - 1;
                m &= m - 1;
                if (((pc[d].behind >> lane) & 1u) && d > lane) bh &= ~(1u << d);
            }
            __syncwarp();
            if (lane < a.nprim) pc[lane].behind = bh;
            __syncwarp();
        }
        // ---- tile classification: lane = tile (bbox reject, corner tests, exact occlusion cull)
        uint32_t my_in = 0, my_part = 0; // bit 31 of my_part: near/far clipping can matter inside this tile
        if (lane < n_tiles) {
            const float cl = (float)((lane & (tiles_x - 1)) * TILE_COLS), ch = cl + (TILE_COLS - 1);
            const float rl = (float)(row0 + (lane >> sh_tx) * TILE_ROWS), rh = rl + (TILE_ROWS - 1);
            const double dcl = cl, dch = ch, drl = rl, drh = rh;
            const double w_near = 1.0 / a.near_, w_far = 1.0 / a.far_;
            // pass 1: classify (bbox reject, corner tests with the float margins); a primitive with a vertex outside
            // [near, far] only matters here if its plane leaves the 1/z range over this tile (1/z is affine)
            bool tile_clipped = false;
            for (int t = 0; t < a.nprim; t++) {
                const PrimCoef& c = pc[t];
                if (!c.valid || c.c_hi < cl || c.c_lo > ch || c.r_hi < rl || c.r_lo > rh) continue;
                bool all_in = true, out = false;
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    const float kl = fmaf(c.fB[k], rl, c.fC[k]), kh = fmaf(c.fB[k], rh, c.fC[k]);
                    const float v00 = fmaf(c.fA[k], cl, kl), v01 = fmaf(c.fA[k], ch, kl), v10 = fmaf(c.fA[k], cl, kh), v11 = fmaf(c.fA[k], ch, kh);
                    const float lo = fminf(fminf(v00, v01), fminf(v10, v11)), hi = fmaxf(fmaxf(v00, v01), fmaxf(v10, v11));
                    const float mgk = k == 4 ? c.wmargin : c.margin;
                    all_in = all_in && (lo > mgk);
                    out = out || (hi < -mgk);
                }
                if (out) continue;
                if (all_in) my_in |= 1u << t; else my_part |= 1u << t;
                if (c.clipped) {
                    const double em = 1e-12 * ((fabs(c.eA[4]) + fabs(c.eB[4])) * S + fabs(c.eC[4]));
                    const double w0 = c.eA[4] * dcl + c.eB[4] * drl + c.eC[4], w1 = c.eA[4] * dch + c.eB[4] * drl + c.eC[4];
                    const double w2 = c.eA[4] * dcl + c.eB[4] * drh + c.eC[4], w3 = c.eA[4] * dch + c.eB[4] * drh + c.eC[4];
                    const double lo = fmin(fmin(w0, w1), fmin(w2, w3)) - em, hi = fmax(fmax(w0, w1), fmax(w2, w3)) + em;
                    if (!(lo > w_far && hi < w_near)) tile_clipped = true;
                }
            }
            if (!tile_clipped && my_in) {
                // pass 2: exact occlusion cull against the nearest covering primitive.  1/z at the corners in fp64
                // (grazing primitives have huge float margins); 1/z affine: nearer at 4 corners = nearer everywhere
                double dom[4] = {-1e300, -1e300, -1e300, -1e300};
                int dom_t = -1;
                uint32_t m = my_in;
                while (m) {
                    const int t = __ffs(m) - 1;
                    m &= m - 1;
                    const PrimCoef& c = pc[t];
                    const double em = 1e-12 * ((fabs(c.eA[4]) + fabs(c.eB[4])) * S + fabs(c.eC[4]));
                    const double w0 = c.eA[4] * dcl + c.eB[4] * drl + c.eC[4] - em;
                    if (w0 > dom[0]) {
                        dom[0] = w0; dom[1] = c.eA[4] * dch + c.eB[4] * drl + c.eC[4] - em;
                        dom[2] = c.eA[4] * dcl + c.eB[4] * drh + c.eC[4] - em; dom[3] = c.eA[4] * dch + c.eB[4] * drh + c.eC[4] - em;
                        dom_t = t;
                    }
                }
                uint32_t keep_in = 0, keep_part = 0, cand = my_in | my_part;
                const uint32_t cover = my_in;
                while (cand) {
                    const int t = __ffs(cand) - 1;
                    cand &= cand - 1;
                    const PrimCoef& c = pc[t];
                    const double em = 1e-12 * ((fabs(c.eA[4]) + fabs(c.eB[4])) * S + fabs(c.eC[4]));
                    const double w0 = c.eA[4] * dcl + c.eB[4] * drl + c.eC[4] + em, w1 = c.eA[4] * dch + c.eB[4] * drl + c.eC[4] + em;
                    const double w2 = c.eA[4] * dcl + c.eB[4] * drh + c.eC[4] + em, w3 = c.eA[4] * dch + c.eB[4] * drh + c.eC[4] + em;
                    // hidden: its plane is behind the dominator's over the tile, or the whole polygon is behind the plane
                    // of some primitive that covers the tile
                    const bool hidden = (w0 < dom[0] && w1 < dom[1] && w2 < dom[2] && w3 < dom[3]) || (c.behind & cover) != 0;
                    if (!hidden || t == dom_t) { if ((my_in >> t) & 1u) keep_in |= 1u << t; else keep_part |= 1u << t; }
                }
                my_in = keep_in; my_part = keep_part;
                RSTAT(1, 1);
            }
            RSTAT(0, 1); RSTAT(2, __popc(my_in)); RSTAT(3, __popc(my_part)); RSTAT(11, tile_clipped ? 1 : 0);
#ifdef TG_RASTER_STATS
            for (int t = 0; t < a.nprim; t++) { if ((my_part >> t) & 1u) RSTAT(16 + t, 1); if ((my_in >> t) & 1u) RSTAT(32 + t, 1); }
#endif
            if (tile_clipped) my_part |= 0x80000000u;
        }
        if (lane < a.nprim) { RSTAT(9, pc[lane].valid && pc[lane].steep ? 1 : 0); RSTAT(10, pc[lane].valid ? 1 : 0); }
        uint8_t* obs_e = a.obs + (size_t)e * S * S + (size_t)row0 * S;
        int cnt_in = 0, cnt_pt = 0; // entries in the two span lists (warp-uniform)
        for (int tile = 0; tile <= n_tiles; tile++) {
            const bool drain = tile == n_tiles;
            if (!drain) {
                // ---- BIN: copy the baked row, list the spans that need shading
                const uint32_t in_m = __shfl_sync(0xffffffffu, my_in, tile), part_w = __shfl_sync(0xffffffffu, my_part, tile);
                const uint32_t part_m = part_w & 0x7fffffffu, tile_clip = part_w >> 31;
                const int lr = (tile >> sh_tx) * TILE_ROWS + (lane >> TILE_SPR_SHIFT), c0 = (tile & (tiles_x - 1)) * TILE_COLS + (lane & ((1 << TILE_SPR_SHIFT) - 1)) * 16;
                const int off = (lr << sh_S) + c0, span = off >> 4;
                *reinterpret_cast<uint4*>(obs_e + off) = *reinterpret_cast<const uint4*>(s_base + off);
                if ((in_m | part_m) == 0) continue;
                const bool skin = (s_skin[span >> 5] >> (span & 31)) & 1u;
                uint32_t sp_in = in_m, sp_part = 0;
                if (skin && part_m) {
                    const float fr = (float)(row0 + lr), fc0 = (float)c0;
                    uint32_t m = part_m;
                    while (m) {
                        const int t = __ffs(m) - 1;
                        m &= m - 1;
                        const PrimCoef& c = pc[t];
                        float lo = 1e30f, hx = 1e30f; // margins subtracted: > 0 means certainly positive
#pragma unroll
                        for (int i = 0; i < 5; i++) {
                            const float mgi = i == 4 ? c.wmargin : c.margin;
                            const float a0 = fmaf(c.fA[i], fc0, fmaf(c.fB[i], fr, c.fC[i])), z0 = fmaf(c.fA[i], 15.0f, a0);
                            lo = fminf(lo, fminf(a0, z0) - mgi);
                            hx = fminf(hx, fmaxf(a0, z0) + mgi);
                        }
                        if (hx < 0.0f) continue;                     // one function is negative over the whole span
                        if (lo > 0.0f) sp_in |= 1u << t;             // span fully inside
                        else sp_part |= 1u << t;
                    }
                }
                const bool act_in = skin && sp_in && !sp_part, act_pt = skin && sp_part;
                const uint32_t bal_in = __ballot_sync(0xffffffffu, act_in), bal_pt = __ballot_sync(0xffffffffu, act_pt);
                SpanEntry en;
                en.off16 = (uint16_t)span; en.in_m = (uint16_t)sp_in; en.part_m = (uint16_t)sp_part; en.pad = (uint16_t)tile_clip;
                if (act_in) l_in[cnt_in + __popc(bal_in & lt_mask)] = en;
                if (act_pt) l_pt[cnt_pt + __popc(bal_pt & lt_mask)] = en;
                cnt_in += __popc(bal_in); cnt_pt += __popc(bal_pt);
                __syncwarp();
            }
            // ---- SHADE: full batches of 32 spans (or what is left when draining), one span per lane
            if (cnt_in >= 32 || (drain && cnt_in > 0)) {
                const int nb = min(cnt_in, 32);
                const SpanEntry en = l_in[lane], tl = l_in[32 + lane];
                __syncwarp();
                if (lane < nb) shade(en, obs_e, std::false_type{});
                if (lane < cnt_in - nb) l_in[lane] = tl;
                cnt_in -= nb;
                __syncwarp();
                if (*s_qcnt > EXACT_QUEUE - 512) flush_queue(obs_e); // the next batch can add up to 512 entries
            }
            if (cnt_pt >= 32 || (drain && cnt_pt > 0)) {
                const int nb = min(cnt_pt, 32);
                const SpanEntry en = l_pt[lane], tl = l_pt[32 + lane];
                __syncwarp();
                if (lane < nb) shade(en, obs_e, std::true_type{});
                if (lane < cnt_pt - nb) l_pt[lane] = tl;
                cnt_pt -= nb;
                __syncwarp();
                if (*s_qcnt > EXACT_QUEUE - 512) flush_queue(obs_e);
            }
            if (drain) flush_queue(obs_e);
        }
    }
}
