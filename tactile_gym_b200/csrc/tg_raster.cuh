// tg_raster.cuh - tactile depth raster + post-process, one uint8 [S][S] image per env.
//
// Replaces TactileSensor.get_imgs (pb.getCameraImage over the whole ~270k-triangle scene,
// sensors/tactile_sensor.py:212-259) + TactileSensor.t_s_camera (:261-294).
//
// What makes it cheap ("decisive simplification"): the camera is rigidly attached to the
// sensor, so everything but the stimulus is static in the camera frame and already baked into the
// reference's nodef_dep / border_mask / nodef_gray images.  Per env only the few stimulus triangles are
// z-tested against nodef_dep.
//
// Kernel shape (HBM-write bound; algorithmic bytes/env = S*S obs + 192 B camera/stimulus state):
//   * persistent CTAs, grid = 2 x #SMs; each CTA owns one row band (S*S/bands pixels) of a strided set of
//     envs, so the band of nodef_dep (f32) and of the pre-baked border image (u8) is fetched ONCE per CTA
//     by TMA bulk copies (cp.async.bulk + mbarrier) into shared memory and reused for every env;
//   * per env, ntri threads turn camera + stimulus pose into fp64 homogeneous edge equations
//     (b = M^-1 d, inside <=> all b_i >= 0, 1/z = sum b_i: exact per-pixel clipping, no vertex projection);
//   * each thread owns 16-pixel row spans: spans outside every triangle's screen bbox are a straight
//     shared-memory -> HBM copy of the baked row; others evaluate the edge equations per pixel;
//   * the post-process is done in float32 with numpy's operation order, output written as one 16-byte
//     store per span (a warp writes 512 contiguous bytes).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tactile_gym_b200.h"

#define RASTER_THREADS 512
#define RASTER_BATCH 4 // envs set up per barrier

struct TriCoef {
    double r0[3], r1[3], r2[3]; // rows of M^-1: b_i = r_i . (dx, dy, 1)
    int c_lo, c_hi, r_lo, r_hi; // screen bbox (inclusive); r_lo > r_hi -> culled
};

struct RasterArgs {
    int n, S, bands, ntri;
    double th;          // tan(fov/2)
    double F, near_, far_; // F = far/(far-near)
    const float* nodef;      // [S*S], border pixels = -1
    const uint8_t* base;     // [S*S], border pixels = (u8)nodef_gray, others 0
    const double* tris;      // [ntri][9] stimulus-frame triangles
    const double* cam;       // [N][12]
    const double* stim;      // [N][12]
    const uint8_t* mask;     // optional [N]
    uint8_t* obs;            // [N][S*S]
    uint8_t* term_obs;       // optional [N][S*S]: previous obs of masked envs is copied here first
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :
                 : "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tri_setup(const RasterArgs& a, const double* cam, const double* stim, const double* tl, TriCoef& o)
{
    // world vertices -> eye space (x right, y up, z forward)
    double ve[3][3];
    bool front = true;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double* v = tl + 3 * k;
        double w[3];
#pragma unroll
        for (int c = 0; c < 3; c++) w[c] = stim[3 * c] * v[0] + stim[3 * c + 1] * v[1] + stim[3 * c + 2] * v[2] + stim[9 + c] - cam[c];
        ve[k][0] = w[0] * cam[9] + w[1] * cam[10] + w[2] * cam[11];
        ve[k][1] = w[0] * cam[6] + w[1] * cam[7] + w[2] * cam[8];
        ve[k][2] = w[0] * cam[3] + w[1] * cam[4] + w[2] * cam[5];
        if (ve[k][2] <= 1e-6) front = false;
    }
    // M = [p0 p1 p2] columns; M^-1 rows = (p1 x p2, p2 x p0, p0 x p1) / det
    double c0[3], c1[3], c2[3];
    c0[0] = ve[1][1] * ve[2][2] - ve[1][2] * ve[2][1]; c0[1] = ve[1][2] * ve[2][0] - ve[1][0] * ve[2][2]; c0[2] = ve[1][0] * ve[2][1] - ve[1][1] * ve[2][0];
    c1[0] = ve[2][1] * ve[0][2] - ve[2][2] * ve[0][1]; c1[1] = ve[2][2] * ve[0][0] - ve[2][0] * ve[0][2]; c1[2] = ve[2][0] * ve[0][1] - ve[2][1] * ve[0][0];
    c2[0] = ve[0][1] * ve[1][2] - ve[0][2] * ve[1][1]; c2[1] = ve[0][2] * ve[1][0] - ve[0][0] * ve[1][2]; c2[2] = ve[0][0] * ve[1][1] - ve[0][1] * ve[1][0];
    const double det = ve[0][0] * c0[0] + ve[0][1] * c0[1] + ve[0][2] * c0[2];
    const int S = a.S;
    o.c_lo = 0; o.c_hi = S - 1; o.r_lo = 0; o.r_hi = S - 1;
    if (fabs(det) < 1e-300) { o.r_lo = 1; o.r_hi = 0; return; }
    const double inv = 1.0 / det;
#pragma unroll
    for (int c = 0; c < 3; c++) { o.r0[c] = c0[c] * inv; o.r1[c] = c1[c] * inv; o.r2[c] = c2[c] * inv; }
    if (front) {
        double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const double x = ve[k][0] / (ve[k][2] * a.th), y = ve[k][1] / (ve[k][2] * a.th);
            xmin = fmin(xmin, x); xmax = fmax(xmax, x); ymin = fmin(ymin, y); ymax = fmax(ymax, y);
        }
        const double cl = floor((xmin + 1) * 0.5 * S - 0.5) - 1, ch = ceil((xmax + 1) * 0.5 * S - 0.5) + 1;
        const double rl = floor((1 - ymax) * 0.5 * S - 0.5) - 1, rh = ceil((1 - ymin) * 0.5 * S - 0.5) + 1;
        o.c_lo = (int)fmax(cl, 0.0); o.c_hi = (int)fmin(ch, (double)(S - 1));
        o.r_lo = (int)fmax(rl, 0.0); o.r_hi = (int)fmin(rh, (double)(S - 1));
        if (ch < 0 || rh < 0 || cl > S - 1 || rl > S - 1) { o.r_lo = 1; o.r_hi = 0; }
    }
}

// t_s_camera's float32 arithmetic (tactile_sensor.py:268-284)
__device__ __forceinline__ uint32_t quantize(float cur, float nd)
{
    float diff = cur - nd;
    const float eps = 1e-4f, maxpen = 0.05f;
    if (diff >= -eps && diff <= eps) diff = 0.0f;
    float pen = fabsf(diff);
    pen = fminf(fmaxf(pen, 0.0f), maxpen);
    const float val = __fmul_rn(__fdiv_rn(pen, maxpen), 255.0f);
    return (uint32_t)__float2uint_rz(val);
}

__global__ void __launch_bounds__(RASTER_THREADS)
raster_kernel(const RasterArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = a.S, band_rows = S / a.bands, band_px = band_rows * S;
    float* s_nodef = reinterpret_cast<float*>(smem_raw);
    uint8_t* s_base = smem_raw + (size_t)band_px * 4;
    TriCoef* s_tri = reinterpret_cast<TriCoef*>(smem_raw + (size_t)band_px * 5);
    __shared__ __align__(8) uint64_t bar;
    __shared__ int s_env[RASTER_BATCH];

    const int band = blockIdx.x % a.bands;
    const int lane_cta = blockIdx.x / a.bands, n_cta = gridDim.x / a.bands;
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

    const int spans_per_row = S / 16, n_spans = band_rows * spans_per_row;
    const double inv_S = 1.0 / S;

    for (int e0 = lane_cta * RASTER_BATCH; e0 < a.n; e0 += n_cta * RASTER_BATCH) {
        __syncthreads(); // previous batch done with s_tri / s_env
        if (threadIdx.x < RASTER_BATCH) {
            const int e = e0 + threadIdx.x;
            s_env[threadIdx.x] = (e < a.n && (!a.mask || a.mask[e])) ? e : -1;
        }
        if (threadIdx.x < RASTER_BATCH * a.ntri) {
            const int bi = threadIdx.x / a.ntri, t = threadIdx.x % a.ntri, e = e0 + bi;
            if (e < a.n && (!a.mask || a.mask[e]))
                tri_setup(a, a.cam + (size_t)e * 12, a.stim + (size_t)e * 12, a.tris + 9 * t, s_tri[bi * a.ntri + t]);
        }
        __syncthreads();
#pragma unroll 1
        for (int bi = 0; bi < RASTER_BATCH; bi++) {
            const int e = s_env[bi];
            if (e < 0) continue;
            const TriCoef* tc = s_tri + bi * a.ntri;
            uint8_t* out = a.obs + (size_t)e * S * S + (size_t)row0 * S;
            uint8_t* tout = a.term_obs ? a.term_obs + (size_t)e * S * S + (size_t)row0 * S : nullptr;
            for (int sp = threadIdx.x; sp < n_spans; sp += RASTER_THREADS) {
                const int lr = sp / spans_per_row, c0 = (sp % spans_per_row) * 16, r = row0 + lr;
                const int off = lr * S + c0;
                if (tout) *reinterpret_cast<uint4*>(tout + off) = *reinterpret_cast<const uint4*>(out + off);
                uint4 res = *reinterpret_cast<const uint4*>(s_base + off);
                // which triangles can touch this span?
                uint32_t hit = 0;
                for (int t = 0; t < a.ntri; t++)
                    if (r >= tc[t].r_lo && r <= tc[t].r_hi && c0 + 15 >= tc[t].c_lo && c0 <= tc[t].c_hi) hit |= 1u << t;
                if (hit) {
                    float nd[16], cur[16];
                    {
                        const float4* p = reinterpret_cast<const float4*>(s_nodef + off);
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const float4 v = p[k];
                            nd[4 * k] = v.x; nd[4 * k + 1] = v.y; nd[4 * k + 2] = v.z; nd[4 * k + 3] = v.w;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 16; k++) cur[k] = nd[k];
                    const double dy = (1.0 - (r + 0.5) * inv_S * 2.0) * a.th;
                    while (hit) {
                        const int t = __ffs(hit) - 1;
                        hit &= hit - 1;
                        const TriCoef c = tc[t];
                        const double k0 = c.r0[1] * dy + c.r0[2], k1 = c.r1[1] * dy + c.r1[2], k2 = c.r2[1] * dy + c.r2[2];
#pragma unroll
                        for (int k = 0; k < 16; k++) {
                            const double dx = ((c0 + k + 0.5) * inv_S * 2.0 - 1.0) * a.th;
                            const double b0 = c.r0[0] * dx + k0, b1 = c.r1[0] * dx + k1, b2 = c.r2[0] * dx + k2;
                            const double w = b0 + b1 + b2;
                            const double tol = -1e-12 * w;
                            if (w > 0.0 && b0 >= tol && b1 >= tol && b2 >= tol && w * a.near_ <= 1.0 && w * a.far_ >= 1.0) {
                                const float d = (float)(a.F * (1.0 - a.near_ * w));
                                cur[k] = fminf(cur[k], d);
                            }
                        }
                    }
                    uint32_t wds[4];
#pragma unroll
                    for (int k4 = 0; k4 < 4; k4++) {
                        uint32_t wd = 0;
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const float n0 = nd[4 * k4 + k];
                            const uint32_t basev = ((k4 == 0 ? res.x : k4 == 1 ? res.y : k4 == 2 ? res.z : res.w) >> (8 * k)) & 0xffu;
                            const uint32_t qv = n0 < 0.0f ? basev : quantize(cur[4 * k4 + k], n0);
                            wd |= qv << (8 * k);
                        }
                        wds[k4] = wd;
                    }
                    res = make_uint4(wds[0], wds[1], wds[2], wds[3]);
                }
                *reinterpret_cast<uint4*>(out + off) = res;
            }
        }
    }
}
