// tg_raster_hf.cuh - tactile raster over a per-env HEIGHTFIELD stimulus (surface_follow).
//
// Replaces, for surface_follow, pb.getCameraImage + t_s_camera (sensors/tactile_sensor.py:212-294) where the only
// moving geometry is the 64 x 64 heightfield of base_surface_env.py:402-424 (7,938 triangles, new every episode).
// Same arithmetic as raster_kernel (tg_raster.cuh): certified float fast path + exact fp64 patch-up; what differs is
// where the primitives come from.  A tile of the image only sees the few heightfield cells under its pixel rays
// between the near plane and the skin (nothing behind the skin can show), so every tile builds its OWN primitive list:
//   1. the 4 corner rays of the tile at z = near and z = deepest skin depth in the tile -> x/y box -> cell range
//   2. lane = triangle: vertices from the env's heights ([EXT] Bullet's mesh: tg_surface.cuh:hf_vertex, diagonal
//      (x+1, y)-(x, y+1)), eye space, screen-space coefficients (prim_from_eye)
//   3. lane = triangle: covers the tile / misses it / crosses it (corner tests with the float margins)
//   4. lane = 16-pixel span: span-level refinement, shade (shade_span), then the warp patches the queued pixels exactly
// A tile that would need more than HF_MAXPRIM triangles is redone span by span (a 16 x 1 pixel region sees <= 8).
#pragma once
#include "tg_raster.cuh"
#include "tg_surface.cuh"

#define HF_THREADS 320   // 10 warps: what fits beside the band tables in shared memory (12.5 KB per warp)
#define HF_WARPS (HF_THREADS / 32)
#define HF_UPE 8          // work units per env image (a unit = every 8th tile, diagonally)
#define HF_MAXPRIM 32
#define HF_QUEUE 512 // a region adds at most 32 spans x 16 pixels
#define HF_PER_WARP_SMEM (sizeof(PrimCoef) * HF_MAXPRIM + sizeof(unsigned long long) * HF_QUEUE + 16)

__global__ void __launch_bounds__(HF_THREADS, 1)
raster_hf_kernel(const RasterArgs a, int* __restrict__ error_flag)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = a.S, band_rows = S / a.bands, band_px = band_rows * S;
    float* s_nodef = reinterpret_cast<float*>(smem_raw);
    uint8_t* s_base = smem_raw + (size_t)band_px * 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t wbase = ((size_t)band_px * 5 + 15) & ~size_t(15);
    PrimCoef* pc = reinterpret_cast<PrimCoef*>(smem_raw + wbase + HF_PER_WARP_SMEM * warp);
    unsigned long long* s_queue = reinterpret_cast<unsigned long long*>(pc + HF_MAXPRIM);
    int* s_qcnt = reinterpret_cast<int*>(s_queue + HF_QUEUE);
    __shared__ __align__(8) uint64_t bar;

    const int sh_b = 31 - __clz(a.bands);
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
    if (lane == 0) *s_qcnt = 0;
    __syncwarp();

    const int sh_S = 31 - __clz(S);
    WarpCtx ctx;
    ctx.pc = pc; ctx.s_nodef = s_nodef; ctx.s_base = s_base; ctx.s_queue = s_queue; ctx.s_qcnt = s_qcnt;
    ctx.S = S; ctx.sh_S = sh_S; ctx.row0 = row0;
    const double kx = a.th * 2.0 / S, x0 = a.th * (1.0 / S - 1.0), y0 = a.th * (1.0 - 1.0 / S);
    const double half = (SURF_N - 1) / 2.0;
    const int tr = a.hf_tile_rows, tc = a.hf_tile_cols;
    const int tiles_x = S / tc, n_tiles = tiles_x * (band_rows / tr);

    // Work unit of a warp: (env, 1 / HF_UPE of the band's tiles).  One env per warp left 14 % of the warps without work at
    // config 3's 1024 envs and made every warp walk 32 tiles in a row; with units the tiles of an image are rendered by up to
    // HF_UPE warps side by side.  Tile t of an image belongs to part (t + t / tiles_x) mod HF_UPE - a diagonal pattern, so that
    // every part gets its share of border tiles (cheap) and centre tiles (expensive) and a static round-robin over the units balances.
    const int upe = min(HF_UPE, n_tiles);
    const int units = a.n * upe;
    for (int u = lane_cta * HF_WARPS + warp; u < units; u += n_cta * HF_WARPS) {
        const int e = u / upe, part = u - e * upe;
        if (a.mask && !a.mask[e]) continue;
        const double* cam = a.cam + (size_t)e * 12;
        const int buf = a.hf_cur[e] ^ (a.hf_flip ? 1 : 0);
        const double* H = a.hf + ((size_t)e * 2 + (size_t)buf) * SURF_PTS;
        const double zc = a.hf_meta[((size_t)e * 2 + (size_t)buf) * SURF_META];
        const double hmax_env = a.hf_meta[((size_t)e * 2 + (size_t)buf) * SURF_META + 7];
        uint8_t* obs_e = a.obs + (size_t)e * S * S + (size_t)row0 * S;

        // one rectangular pixel region (band-local rows): false = it needs more than HF_MAXPRIM triangles
        auto region = [&](int rl0, int nrows, int cl0, int ncols) -> bool {
            const int sh_spr = 31 - __clz(ncols >> 4), nspans = nrows << sh_spr;
            const int lr = rl0 + (lane >> sh_spr), c0 = cl0 + ((lane & ((1 << sh_spr) - 1)) << 4), off = (lr << sh_S) + c0;
            bool skin = false;
            float dmax = -1.0f;
            if (lane < nspans) {
                *reinterpret_cast<uint4*>(obs_e + off) = *reinterpret_cast<const uint4*>(s_base + off);
#pragma unroll
                for (int k = 0; k < 16; k++) dmax = fmaxf(dmax, s_nodef[off + k]);
                skin = dmax >= 0.0f;
            }
            if (__ballot_sync(0xffffffffu, skin) == 0u) return true; // border only
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
            // ---- 1. cells under the region's rays.  Only the ray segments below the terrain's top and in front of the
            // deepest skin depth can meet visible terrain; the top starts as the env's highest point and is then
            // tightened to the highest vertex of the cells found so far (sound: above that height, inside those cells,
            // there is nothing to hit), which shrinks the segment - and the cell range - to the local relief.
            const double zmax = a.near_ * a.F / (a.F - (double)dmax) * (1.0 + 1e-6);
            const double cc = (lane & 1) ? (double)(cl0 + ncols) - 0.5 : (double)cl0 - 0.5;
            const double rr = (lane & 2) ? (double)(row0 + rl0 + nrows) - 0.5 : (double)(row0 + rl0) - 0.5;
            const double rdx = kx * cc + x0, rdy = y0 - kx * rr;
            const double gz = cam[11] * rdx + cam[8] * rdy + cam[5]; // world-z change per unit eye depth along this corner ray
            double zt = a.surf_pos[2] + hmax_env - zc + 1e-7;
            int i0 = 0, i1 = -1, j0 = 0, j1 = -1;
#pragma unroll 1
            for (int it = 0; it < 4; it++) {
                // eye depth from which the corner ray is below world height zt (looking up / already below: from the near plane)
                double zlo = gz < 0.0 ? fmax((zt - cam[2]) / gz, a.near_) : a.near_;
                zlo = fmin(zlo, __shfl_xor_sync(0xffffffffu, zlo, 1));
                zlo = fmin(zlo, __shfl_xor_sync(0xffffffffu, zlo, 2));
                zlo = __shfl_sync(0xffffffffu, zlo, 0) * (1.0 - 1e-9);
                if (zlo > zmax) return true; // the terrain stays behind the skin everywhere in this region
                const double z = (lane & 4) ? zmax : zlo;
                const double ex = rdx * z, ey = rdy * z;
                double xmin = cam[0] + cam[9] * ex + cam[6] * ey + cam[3] * z, xmax = xmin;
                double ymin = cam[1] + cam[10] * ex + cam[7] * ey + cam[4] * z, ymax = ymin;
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) {
                    xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, o)); xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
                    ymin = fmin(ymin, __shfl_xor_sync(0xffffffffu, ymin, o)); ymax = fmax(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
                }
                xmin = __shfl_sync(0xffffffffu, xmin, 0); xmax = __shfl_sync(0xffffffffu, xmax, 0);
                ymin = __shfl_sync(0xffffffffu, ymin, 0); ymax = __shfl_sync(0xffffffffu, ymax, 0);
                // vertex j sits at x = surf_pos.x + float32((j - 31.5) grid): widen by 1e-6 cells for that rounding
                const double fj0 = floor((xmin - a.surf_pos[0]) / a.surf_grid + half - 1e-6), fj1 = floor((xmax - a.surf_pos[0]) / a.surf_grid + half + 1e-6);
                const double fi0 = floor((ymin - a.surf_pos[1]) / a.surf_grid + half - 1e-6), fi1 = floor((ymax - a.surf_pos[1]) / a.surf_grid + half + 1e-6);
                if (fj1 < 0.0 || fi1 < 0.0 || fj0 > (double)(SURF_N - 2) || fi0 > (double)(SURF_N - 2)) return true; // off the grid
                j0 = (int)fmax(fj0, 0.0); j1 = (int)fmin(fj1, (double)(SURF_N - 2));
                i0 = (int)fmax(fi0, 0.0); i1 = (int)fmin(fi1, (double)(SURF_N - 2));
                if (2 * (j1 - j0 + 1) * (i1 - i0 + 1) <= HF_MAXPRIM / 2 || it == 3) break;
                // highest vertex of these cells -> tighter top
                const int nvj = j1 - j0 + 2, nvert = nvj * (i1 - i0 + 2);
                float hm = -3.0e38f;
                for (int v = lane; v < nvert; v += 32) hm = fmaxf(hm, (float)H[(i0 + v / nvj) * SURF_N + j0 + v % nvj]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) hm = fmaxf(hm, __shfl_xor_sync(0xffffffffu, hm, o));
                const double zt_new = a.surf_pos[2] + (double)hm - zc + 1e-7;
                if (!(zt_new < zt)) break;
                zt = zt_new;
            }
            const int ncj = j1 - j0 + 1, nprim = 2 * ncj * (i1 - i0 + 1);
            if (nprim > HF_MAXPRIM) return false;
            // ---- 2. lane = triangle
            if (lane < nprim) {
                const int cell = lane >> 1, ci = i0 + cell / ncj, cj = j0 + cell % ncj;
                double vw[3][3], ve[4][3], vp[4][3];
                if (lane & 1) {
                    hf_vertex(a.surf_pos, a.surf_grid, H, zc, ci, cj + 1, vw[0]);
                    hf_vertex(a.surf_pos, a.surf_grid, H, zc, ci + 1, cj, vw[1]);
                    hf_vertex(a.surf_pos, a.surf_grid, H, zc, ci + 1, cj + 1, vw[2]);
                } else {
                    hf_vertex(a.surf_pos, a.surf_grid, H, zc, ci, cj, vw[0]);
                    hf_vertex(a.surf_pos, a.surf_grid, H, zc, ci + 1, cj, vw[1]);
                    hf_vertex(a.surf_pos, a.surf_grid, H, zc, ci, cj + 1, vw[2]);
                }
#pragma unroll
                for (int k = 0; k < 3; k++) world_to_eye(cam, vw[k], ve[k]);
#pragma unroll
                for (int c = 0; c < 3; c++) ve[3][c] = ve[2][c];
                prim_from_eye(a, ve, 3, pc[lane], vp);
            }
            __syncwarp();
            // ---- 3. lane = triangle against the region rectangle
            bool is_in = false, is_part = false, clip = false;
            if (lane < nprim && pc[lane].valid) {
                const PrimCoef& c = pc[lane];
                const float cl = (float)cl0, ch = (float)(cl0 + ncols - 1), rl = (float)(row0 + rl0), rh = (float)(row0 + rl0 + nrows - 1);
                bool all_in = true, out = c.c_hi < cl || c.c_lo > ch || c.r_hi < rl || c.r_lo > rh;
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    const float kl = fmaf(c.fB[k], rl, c.fC[k]), kh = fmaf(c.fB[k], rh, c.fC[k]);
                    const float v00 = fmaf(c.fA[k], cl, kl), v01 = fmaf(c.fA[k], ch, kl), v10 = fmaf(c.fA[k], cl, kh), v11 = fmaf(c.fA[k], ch, kh);
                    const float lo = fminf(fminf(v00, v01), fminf(v10, v11)), hi = fmaxf(fmaxf(v00, v01), fmaxf(v10, v11));
                    const float mgk = k == 4 ? c.wmargin : c.margin;
                    all_in = all_in && (lo > mgk);
                    out = out || (hi < -mgk);
                }
                if (!out) {
                    is_in = all_in; is_part = !all_in;
                    if (c.clipped) {
                        const double em = 1e-12 * ((fabs(c.eA[4]) + fabs(c.eB[4])) * S + fabs(c.eC[4]));
                        const double w0 = c.eA[4] * cl + c.eB[4] * rl + c.eC[4], w1 = c.eA[4] * ch + c.eB[4] * rl + c.eC[4];
                        const double w2 = c.eA[4] * cl + c.eB[4] * rh + c.eC[4], w3 = c.eA[4] * ch + c.eB[4] * rh + c.eC[4];
                        const double lo = fmin(fmin(w0, w1), fmin(w2, w3)) - em, hi = fmax(fmax(w0, w1), fmax(w2, w3)) + em;
                        clip = !(lo > 1.0 / a.far_ && hi < 1.0 / a.near_);
                    }
                }
            }
            const uint32_t in_m = __ballot_sync(0xffffffffu, is_in), part_m = __ballot_sync(0xffffffffu, is_part);
            const bool tile_clip = __ballot_sync(0xffffffffu, clip) != 0u;
            if ((in_m | part_m) == 0u) return true;
            // ---- 4. lane = span
            if (lane < nspans && skin) {
                uint32_t sp_in = in_m, sp_part = 0;
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
                    if (hx < 0.0f) continue;         // one function is negative over the whole span
                    if (lo > 0.0f) sp_in |= 1u << t; // span fully inside
                    else sp_part |= 1u << t;
                }
                if (sp_part) shade_span<true, unsigned long long>(a, ctx, off, sp_in, sp_part, tile_clip, obs_e);
                else if (sp_in) shade_span<false, unsigned long long>(a, ctx, off, sp_in, 0u, tile_clip, obs_e);
            }
            flush_exact<unsigned long long>(a, ctx, obs_e, lane);
            return true;
        };

        for (int tile = 0; tile < n_tiles; tile++) {
            if ((tile + tile / tiles_x) % upe != part) continue;
            const int rl0 = (tile / tiles_x) * tr, cl0 = (tile % tiles_x) * tc;
            if (region(rl0, tr, cl0, tc)) continue;
            // too many triangles under this tile: span by span
            for (int sr = 0; sr < tr; sr++)
                for (int sc = 0; sc < tc; sc += 16)
                    if (!region(rl0 + sr, 1, cl0 + sc, 16) && lane == 0) *error_flag = 2;
        }
    }
}
