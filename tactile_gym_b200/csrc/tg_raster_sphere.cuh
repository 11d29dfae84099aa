// tg_raster_sphere.cuh - tactile raster when the stimulus is a SPHERE (object_roll's marble).
//
// Replaces pb.getCameraImage + t_s_camera (sensors/tactile_sensor.py:212-294) for object_roll
// (rl_envs/nonprehensile_manipulation/object_roll/object_roll_env.py; sphere.urdf: <sphere radius="0.0025"/> x the episode's
// globalScaling).  [EXT] pybullet draws a <sphere> visual as a tessellated mesh of unknown resolution; the analytic sphere
// is rendered (oracle/tg_oracle.c:or_tactile_image_sphere; < 0.2 output LSB from a 32-segment tessellation).
//
// The marble covers a few hundred pixels, so the kernel is a streaming copy of the baked border / zero image (HBM-write
// bound: S*S bytes per env) with an fp64 ray / sphere intersection for the pixels inside the sphere's screen box: no float
// fast path is needed.  One warp per env at a time, one 16-pixel span per lane, 16-byte stores.
#pragma once
#include "tg_raster.cuh"

#define SPH_THREADS 256

__global__ void __launch_bounds__(SPH_THREADS)
raster_sphere_kernel(const RasterArgs a)
{
    const int S = a.S, px = S * S, spans = px >> 4;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, nwarp = (gridDim.x * blockDim.x) >> 5;
    for (int e = warp; e < a.n; e += nwarp) {
        if (a.mask && !a.mask[e]) continue;
        const double* cam = a.cam + (size_t)e * 12;
        const double* st = a.stim + (size_t)e * 12;
        const double R = st[0];
        const double d0[3] = {st[9] - cam[0], st[10] - cam[1], st[11] - cam[2]};
        // sphere centre in eye space: x right, y up, z forward
        const double s[3] = {d0[0] * cam[9] + d0[1] * cam[10] + d0[2] * cam[11], d0[0] * cam[6] + d0[1] * cam[7] + d0[2] * cam[8],
                             d0[0] * cam[3] + d0[1] * cam[4] + d0[2] * cam[5]};
        const double ss = s[0] * s[0] + s[1] * s[1] + s[2] * s[2] - R * R;
        // conservative screen box of the sphere (pixels outside cannot hit it)
        int c0 = 0, c1 = S - 1, r0 = 0, r1 = S - 1;
        if (s[2] - R > 1e-6) {
            const double zn = (s[2] - R) * a.th, zf = (s[2] + R) * a.th;
            const double xl = (s[0] - R) / ((s[0] - R) < 0 ? zn : zf), xh = (s[0] + R) / ((s[0] + R) > 0 ? zn : zf);
            const double yl = (s[1] - R) / ((s[1] - R) < 0 ? zn : zf), yh = (s[1] + R) / ((s[1] + R) > 0 ? zn : zf);
            c0 = max(0, (int)floor((xl + 1.0) * 0.5 * S - 0.5) - 1); c1 = min(S - 1, (int)ceil((xh + 1.0) * 0.5 * S - 0.5) + 1);
            r0 = max(0, (int)floor((1.0 - yh) * 0.5 * S - 0.5) - 1); r1 = min(S - 1, (int)ceil((1.0 - yl) * 0.5 * S - 0.5) + 1);
        }
        uint8_t* obs_e = a.obs + (size_t)e * px;
        for (int sp = lane; sp < spans; sp += 32) {
            const int off = sp << 4, r = off / S, cb = off - r * S;
            uint4 v = *reinterpret_cast<const uint4*>(a.base + off);
            if (r >= r0 && r <= r1 && cb + 15 >= c0 && cb <= c1) {
                uint32_t w[4] = {v.x, v.y, v.z, v.w};
                const double yn = 1.0 - (r + 0.5) / S * 2.0;
#pragma unroll 1
                for (int k = 0; k < 16; k++) {
                    const int c = cb + k;
                    const float nd = a.nodef[off + k];
                    if (nd < 0.0f || c < c0 || c > c1) continue;   // border pixel (baked) / outside the box
                    const double xn = (c + 0.5) / S * 2.0 - 1.0;
                    const double d[3] = {xn * a.th, yn * a.th, 1.0};
                    const double dd = d[0] * d[0] + d[1] * d[1] + 1.0, ds = d[0] * s[0] + d[1] * s[1] + s[2];
                    const double disc = ds * ds - dd * ss;
                    uint32_t q = 0;
                    if (disc >= 0.0) {
                        const double z = (ds - sqrt(disc)) / dd;
                        if (z >= a.near_ && z <= a.far_) {
                            const float dep = (float)(a.far_ / (a.far_ - a.near_) * (1.0 - a.near_ / z));
                            q = quantize(fminf(nd, dep), nd);
                        }
                    }
                    w[k >> 2] = (w[k >> 2] & ~(0xffu << ((k & 3) * 8))) | (q << ((k & 3) * 8));
                }
                v = make_uint4(w[0], w[1], w[2], w[3]);
            }
            *reinterpret_cast<uint4*>(obs_e + off) = v;
        }
    }
}
