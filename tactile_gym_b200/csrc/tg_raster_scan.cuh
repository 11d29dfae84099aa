// tg_raster_scan.cuh - scanline raster for stimuli made of CONVEX parts (edge box, cube, pole = plate + post).
//
// Same job and same tables as raster_kernel (tg_raster.cuh): per env, z-test the stimulus against the baked nodef_dep and apply
// t_s_camera's post-process (sensors/tactile_sensor.py:212-294).  What changes is how coverage is found.  raster_kernel treats
// every polygon on its own: tile / span classification against five edge functions per primitive, a float fast path with error
// margins, an fp64 patch-up queue for the pixels the margins cannot decide - 16.7 k warp-instructions per 128 x 128 image
// (profiles/r01_raster_kernel_final.md).  A convex part needs none of that:
//   * only its FRONT faces (camera outside the face's half-space) can be the nearest surface, and on the screen they tile the
//     part's silhouette without overlap - no depth test between them, the back faces are never looked at;
//   * a face is a convex polygon, so on an image row it covers ONE column interval, whose ends are where the row meets the
//     face's edge lines: per face and row two integers, computed exactly (fp64, the oracle's `E >= -1e-12` rule) once per env;
//   * a pixel inside a face's interval takes that face's depth - ONE fp64 FMA for 1/z (affine in the pixel), then literally the
//     oracle's arithmetic (fp64 depth -> float32 -> t_s_camera's float32 post-process): the bytes equal the CPU oracle's by
//     construction, so there is no float-margin analysis, no uncertainty queue and no second pass;
//   * several parts (the pole): nearest wins per pixel, max over parts of 1/z.
// Every output row segment is written exactly once (16-byte stores), baked border bytes included.
// Envs the shortcut does not cover - a front face cut by the NEAR plane inside the image, or the eye inside a part: GL then shows
// back faces - are flagged in `fallback` and rendered by raster_kernel in a masked second launch that exits at once when nothing
// was flagged.  (Faces reaching behind the eye or beyond the far plane are fine, see scan_setup_kernel: a tilted pole's plate does.)
#pragma once
#include "tg_raster.cuh"

#define SCAN_THREADS 1024  // 32 warps per SM: the render kernel is kept under 64 registers (the fp64 face set-up lives in its own kernel)
#define SCAN_WARPS (SCAN_THREADS / 32)
#define SCAN_MAXPOOLS 32 // unit counters per band
#define SCAN_MAXPARTS 4  // convex parts of a stimulus (the pole has 2)
#define SCAN_MAXFRONT 8 // front faces per env kept in the interval table (a box shows <= 3, the pole <= 6)

struct ScanFace {
    double w[3];                // 1/z_eye = w[0] c + w[1] r + w[2] on this face's plane (PrimCoef index 4)
    double ea[4], eb[4];        // edge i crosses row r at column ea[i] r + eb[i] (tolerance folded in); for an edge parallel to
                                // the rows (dir 0) they hold its (eB, eC): inside <=> eB r + eC >= -1e-12
    signed char dir[4];         // +1: columns >= crossing are inside, -1: columns <= crossing, 0: row-parallel edge, 2: unused slot
    int part;
};

// what scan_setup_kernel leaves per env for the render kernel (800 bytes, fetched per work unit with 16-byte cp.async's)
struct __align__(16) ScanEnv {
    int nf, c_lo, c_hi, r_lo, r_hi, pad[3]; // front faces; columns / rows any of them can touch (nf < 0: rendered by raster_kernel)
    ScanFace face[SCAN_MAXFRONT];
};
static_assert(sizeof(ScanFace) == 96 && sizeof(ScanEnv) == 32 + 96 * SCAN_MAXFRONT && sizeof(ScanEnv) % 16 == 0, "ScanEnv layout");

#define SCAN_UNIT_ROWS 32 // most rows a work unit of the render kernel can have (16 by default, TG_SCAN_UNIT_ROWS)

// per warp of the render kernel: two ScanEnv buffers (the next unit's is prefetched), the faces' row intervals, the rows' shade
// masks, the half-span list
__host__ __device__ inline size_t scan_per_warp_smem(int S, int unit_rows)
{
    return (2 * sizeof(ScanEnv) + (size_t)SCAN_MAXFRONT * unit_rows * 2 + (size_t)unit_rows * 4 + (size_t)unit_rows * S / 8 * 2 + 15) & ~size_t(15);
}
// band tables (nodef f32, baked bytes, half-span skin bitmap) in front of the warps' tables
__host__ __device__ inline size_t scan_tables_smem(int band_px) { return ((size_t)band_px * 5 + (size_t)((band_px / 8 + 31) / 32) * 4 + 15) & ~size_t(15); }

// one front face, one row: the inclusive column interval [lo, hi] it covers (lo > hi: none)
__device__ __forceinline__ void scan_interval(const ScanFace& f, int r, int S, int& lo, int& hi)
{
    lo = 0; hi = S - 1;
    const double dr = (double)r;
    const int dirs = *reinterpret_cast<const int*>(f.dir);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int d = (int)(signed char)(dirs >> (8 * i));
        if (d == 2) continue;
        const double x = fma(f.ea[i], dr, f.eb[i]);
        if (d == 0) { if (x < -1e-12) { lo = 1; hi = 0; } continue; }
        // ceil / floor inside the conversion, which saturates: a crossing far off the image leaves the interval empty or untouched
        if (d > 0) lo = max(lo, __double2int_ru(x));
        else hi = min(hi, __double2int_rd(x));
    }
}

// Pre-pass, `lpe` lanes per env (8 / 16 / 32, the next power of two above the primitive count), lane = primitive: eye-space
// vertices, plane, edge lines, front-facing test (fp64, a few hundred instructions per env - kept out of the render kernel so
// that one stays small in registers and code).
__global__ void __launch_bounds__(128)
scan_setup_kernel(const RasterArgs a, const int* __restrict__ prim_part, const double* __restrict__ part_cen, ScanEnv* __restrict__ out,
                  uint8_t* __restrict__ fallback, int* __restrict__ fb_count, int* __restrict__ fb_next, int* __restrict__ ctr, int lpe)
{
    // house-keeping for the launches behind this one (stream order): the render kernel's unit counters and the fallback counter
    // of the NEXT raster pass start from zero (this pass counts in `fb_count`, which the previous pass's set-up cleared)
    if (blockIdx.x == 0) { ctr[threadIdx.x] = 0; if (threadIdx.x == 0) *fb_next = 0; }   // 128 threads = 4 bands x SCAN_MAXPOOLS counters
    const int lane = threadIdx.x & 31, sub = lane & (lpe - 1), gbase = lane - sub;
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) / lpe;
    const uint32_t gmask = (lpe == 32 ? 0xffffffffu : ((1u << lpe) - 1u)) << gbase;
    const int S = a.S;
    const bool live = e < a.n;
    const bool masked = live && a.mask && !a.mask[e];
    bool bad = false, front = false;
    int my_part = -1, why = 0;   // why: 1 near-plane cut, 2 eye inside a part, 3 too many front faces, 4 test hook (profiling: the value stored in `fallback`)
    ScanFace mine;
    int c_lo = S, c_hi = -1, r_lo = S, r_hi = -1;
    if (live && !masked && sub < a.nprim) {
        const double* cam = a.cam + (size_t)e * 12;
        const double* stim = a.stim + (size_t)e * 12;
        const int nv = a.prim_nv[sub];
        double ve[4][3], vp[4][3];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const double* v = a.prims + 12 * sub + 3 * (k < nv ? k : nv - 1);
            double w[3];
#pragma unroll
            for (int c = 0; c < 3; c++) w[c] = stim[3 * c] * v[0] + stim[3 * c + 1] * v[1] + stim[3 * c + 2] * v[2] + stim[9 + c] - cam[c];
            ve[k][0] = w[0] * cam[9] + w[1] * cam[10] + w[2] * cam[11];
            ve[k][1] = w[0] * cam[6] + w[1] * cam[7] + w[2] * cam[8];
            ve[k][2] = w[0] * cam[3] + w[1] * cam[4] + w[2] * cam[5];
        }
        PrimCoef pc;
        const bool infront = prim_from_eye(a, ve, nv, pc, vp);
        my_part = prim_part[sub];
        // (a face whose plane passes through the eye - pc.valid == 0 - is seen edge-on: no area, no 1/z form, skipped as in raster_kernel)
        if (pc.valid) {
            // camera outside this face's half-space <=> the face is a front face.  Plane n . x = h through the face, the
            // part's centroid on the inner side: outside <=> (n . 0 - h) = -h and sc = (n . cen - h) have opposite signs,
            // i.e. h and sc have the same sign
            const int part = prim_part[sub];
            const double* pcn = part_cen + 3 * part;
            double cw[3], ce[3];
#pragma unroll
            for (int c = 0; c < 3; c++) cw[c] = stim[3 * c] * pcn[0] + stim[3 * c + 1] * pcn[1] + stim[3 * c + 2] * pcn[2] + stim[9 + c];
            world_to_eye(cam, cw, ce);
            const double e1[3] = {ve[1][0] - ve[0][0], ve[1][1] - ve[0][1], ve[1][2] - ve[0][2]};
            const double e2[3] = {ve[2][0] - ve[0][0], ve[2][1] - ve[0][1], ve[2][2] - ve[0][2]};
            const double nrm[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
            const double h = nrm[0] * ve[0][0] + nrm[1] * ve[0][1] + nrm[2] * ve[0][2];
            const double sc = nrm[0] * ce[0] + nrm[1] * ce[1] + nrm[2] * ce[2] - h;
            front = (h > 0.0) == (sc > 0.0) && sc != 0.0;
            // Clipping.  The edge functions are homogeneous (built from the eye-space vertices, no projection), so a face with
            // vertices behind the eye or beyond the far plane needs nothing special: where its plane is behind the eye or
            // beyond `far`, the window depth d = F - F near / z comes out >= 1 > nodef and the pixel stays 0, as in the oracle,
            // which does not count it as covered.  The NEAR plane is different: a front face nearer than `near` is cut open
            // there and GL shows the part's inside (its back faces), which this kernel never looks at - such envs go to
            // raster_kernel.  Two necessary conditions for such a cut inside the image, both cheap: the face's PLANE is nearer than
            // `near` on some pixel of the image (1/z is affine in the pixel: test the four corner pixels), and - when all vertices
            // are in front of the eye, so that 1/z is affine on the face with its extremes at the vertices - some VERTEX is.
            if (front) {
                const double cS = (double)(S - 1);
                const double w00 = pc.eC[4], w01 = pc.eA[4] * cS + pc.eC[4], w10 = pc.eB[4] * cS + pc.eC[4], w11 = pc.eA[4] * cS + pc.eB[4] * cS + pc.eC[4];
                bool near_cut = !(fmax(fmax(w00, w01), fmax(w10, w11)) < (1.0 - 1e-9) / a.near_);
                if (near_cut && infront) {
                    bool vertex_near = false;
                    for (int k = 0; k < nv; k++) vertex_near = vertex_near || !(ve[k][2] >= a.near_);
                    near_cut = vertex_near;
                }
                bad = near_cut;
                if (bad) why = 1;
            }
            mine.w[0] = pc.eA[4]; mine.w[1] = pc.eB[4]; mine.w[2] = pc.eC[4];
            mine.part = part;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                mine.dir[i] = 2; mine.ea[i] = 0; mine.eb[i] = 0;
                if (i < nv) {
                    const double A = pc.eA[i], B = pc.eB[i], C = pc.eC[i];
                    if (fabs(A) * (double)S < 1e-9 * (fabs(B) * (double)S + fabs(C) + 1e-300)) {
                        mine.dir[i] = 0; mine.ea[i] = B; mine.eb[i] = C;         // edge line parallel to the rows
                    } else {
                        // A c + B r + C >= -1e-12  <=>  c >= (-1e-12 - C - B r) / A  (A > 0), <= for A < 0
                        const double inv = 1.0 / A;
                        mine.dir[i] = A > 0.0 ? 1 : -1;
                        mine.ea[i] = -B * inv; mine.eb[i] = (-1e-12 - C) * inv;
                    }
                }
            }
            if (front) {
                c_lo = max(0, (int)floor(fmin(fmax((double)pc.c_lo, -1.0), (double)S)));
                c_hi = min(S - 1, (int)ceil(fmin(fmax((double)pc.c_hi, -1.0), (double)S)));
                r_lo = max(0, (int)floor(fmin(fmax((double)pc.r_lo, -1.0), (double)S)));
                r_hi = min(S - 1, (int)ceil(fmin(fmax((double)pc.r_hi, -1.0), (double)S)));
            }
        }
    }
    if (a.scan_test_fallback && (e & 1) && live && !masked) { bad = true; why = 4; }
    // a convex part shows at least one front face to any eye outside it: a part without one has the eye inside (its back faces
    // would be what GL draws) - raster_kernel's case
    for (int p = 0; p < SCAN_MAXPARTS; p++) {
        const uint32_t has = __ballot_sync(0xffffffffu, my_part == p) & gmask, shows = __ballot_sync(0xffffffffu, my_part == p && front) & gmask;
        if (has != 0u && shows == 0u) { bad = true; why = max(why, 2); }
    }
    const uint32_t bad_m = __ballot_sync(0xffffffffu, bad) & gmask;
    const uint32_t front_m = __ballot_sync(0xffffffffu, front) & gmask;
    const int nf = __popc(front_m);
    const bool give_up = bad_m != 0u || nf > SCAN_MAXFRONT;   // raster_kernel renders this env (masked second launch)
    if (nf > SCAN_MAXFRONT) why = 3;
    for (int d = lpe >> 1; d > 0; d >>= 1) why = max(why, __shfl_xor_sync(0xffffffffu, why, d));
    if (front && !give_up) out[e].face[__popc(front_m & ((1u << lane) - 1u))] = mine;
    // the rows / columns any front face can touch (union of the conservative screen boxes)
    for (int d = lpe >> 1; d > 0; d >>= 1) {
        c_lo = min(c_lo, __shfl_xor_sync(0xffffffffu, c_lo, d)); c_hi = max(c_hi, __shfl_xor_sync(0xffffffffu, c_hi, d));
        r_lo = min(r_lo, __shfl_xor_sync(0xffffffffu, r_lo, d)); r_hi = max(r_hi, __shfl_xor_sync(0xffffffffu, r_hi, d));
    }
    if (live && sub == 0) {
        if (masked) { fallback[e] = 0; out[e].nf = -1; }
        else if (give_up) { fallback[e] = (uint8_t)max(why, 1); atomicAdd(fb_count, 1); out[e].nf = -1; }
        else { fallback[e] = 0; out[e].nf = nf; out[e].c_lo = c_lo; out[e].c_hi = c_hi; out[e].r_lo = r_lo; out[e].r_hi = r_hi; }
    }
}

// Render kernel.  A CTA owns ONE row band of the image (128 x 128 and smaller: the whole image; 256 x 256: a quarter) and
// fetches that band's static tables (nodef_dep f32, baked border bytes, the half-span skin bitmap) once, by TMA bulk copies
// into shared memory.  After that its 32 warps run on their own: a work unit is (env, `unit_rows` consecutive rows of the band),
// handed out through one global counter per band (`ctr[band]`, zeroed by the set-up kernel).  Units the stimulus does not touch
// are a plain copy and cost a fraction of the others; the counter evens that out.  Both round trips a unit needs - the counter
// and the env's ScanEnv record - are software-pipelined: while unit i is rendered, unit i+1's record is on its way into the
// warp's second buffer (cp.async) and the counter is being asked for unit i+2.
// unit_rows (16 or 32) and band_rows / unit_rows are powers of two.
template <int sh_unit>
__global__ void __launch_bounds__(SCAN_THREADS, 1)
raster_scan_kernel(const RasterArgs a, const ScanEnv* __restrict__ envs, const uint32_t* __restrict__ skin8, int* __restrict__ ctr, int pools)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = a.S, bands = a.bands, band_rows = S / bands, band_px = band_rows * S;
    constexpr int unit_rows = 1 << sh_unit;
    const int sh_parts = (31 - __clz(band_rows)) - sh_unit;
    float* s_nodef = reinterpret_cast<float*>(smem_raw);
    uint8_t* s_base = smem_raw + (size_t)band_px * 4;
    uint32_t* s_skin = reinterpret_cast<uint32_t*>(smem_raw + (size_t)band_px * 5);   // 1 bit per 8-pixel half span of the band: has a non-border pixel
    const int skin_words = (band_px / 8 + 31) / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* wsm = smem_raw + scan_tables_smem(band_px) + scan_per_warp_smem(S, unit_rows) * warp;
    ScanEnv* s_env = reinterpret_cast<ScanEnv*>(wsm);                                             // [2]
    uchar2* s_iv = reinterpret_cast<uchar2*>(wsm + 2 * sizeof(ScanEnv));                          // [SCAN_MAXFRONT][unit_rows] (lo, hi)
    uint32_t* s_rowm = reinterpret_cast<uint32_t*>(s_iv + (size_t)SCAN_MAXFRONT * unit_rows);     // [unit_rows] half spans of the row to shade
    uint16_t* s_list = reinterpret_cast<uint16_t*>(s_rowm + unit_rows);                           // [unit_rows * S / 8] the same as a list
    __shared__ __align__(8) uint64_t bar;
    __shared__ int s_one[32];

    const int band = blockIdx.x % bands;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        const uint32_t bytes = (uint32_t)band_px * 5u + (uint32_t)skin_words * 4u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        tma_bulk_load(s_nodef, a.nodef + (size_t)band * band_px, (uint32_t)band_px * 4u, &bar);
        tma_bulk_load(s_base, a.base + (size_t)band * band_px, (uint32_t)band_px, &bar);
        tma_bulk_load(s_skin, skin8 + (size_t)band * skin_words, (uint32_t)skin_words * 4u, &bar);
    }
    // a warp's next unit: lane 0 asks the band's counter; the result register is only read a round later, so the atomic's
    // latency is hidden behind the current unit.  ptxas rewrites atomic adds it can prove warp-uniform into its aggregated form
    // (vote, one atomic, a SHUFFLE OF THE RESULT right behind it - which waits for the whole round trip); an addend it cannot
    // see through (a 1 read back from shared memory) under a real branch keeps the plain instruction.
    // One counter for the whole band would see ~0.8 atomics per ns at 4096 envs - which is as fast as one address goes, and was
    // measured to pin the kernel's duration.  So the band's units are cut into `pools` contiguous slices, each with its own
    // counter and its own ~8 CTAs (the CTAs of a band take the pools round robin): still hundreds of envs per pool to even out.
    const int units_band = a.n << sh_parts;
    const int pool = (blockIdx.x / bands) % pools;
    const int u_first = (int)(((long long)units_band * pool) / pools), units = (int)(((long long)units_band * (pool + 1)) / pools);
    int* my_ctr = ctr + band * SCAN_MAXPOOLS + pool;
    if (warp == 0) s_one[lane] = S > 0;
    __syncthreads();      // s_one, and the mbarrier's init, are visible to every warp
    const int one = *reinterpret_cast<volatile int*>(&s_one[lane]);
    int u_req = 0;
    auto request = [&]() {
        if (lane == 0) asm volatile("atom.global.add.u32 %0, [%1], %2;\n" : "+r"(u_req) : "l"(my_ctr), "r"(one) : "memory");
    };
    // unit u's env record -> buffer `b` of this warp: 50 16-byte pieces
    auto fetch = [&](int u, int b) {
        if (u < units) {
            const unsigned char* src = reinterpret_cast<const unsigned char*>(envs + (u >> sh_parts));
            const uint32_t dst = smem_u32(s_env + b);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst + lane * 16u), "l"(src + lane * 16) : "memory");
            if (lane < (int)(sizeof(ScanEnv) / 16) - 32)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst + (lane + 32) * 16u), "l"(src + (lane + 32) * 16) : "memory");
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    request();
    int u = u_first + __shfl_sync(0xffffffffu, u_req, 0);     // start-up: the one exposed round trip
    request();
    fetch(u, 0);
    const int sh_S = 31 - __clz(S), sh_H = sh_S - 3;      // S / 8 = 1 << sh_H half spans per row (4 .. 32)
    const double Fn = a.F * a.near_;
    const int spans = (unit_rows << sh_S) >> 4;           // 16-pixel spans of one unit
    const uint32_t row_all = sh_H == 5 ? 0xffffffffu : ((1u << (1 << sh_H)) - 1u);
    bool tables = false;                                  // this thread has seen the band tables arrive
    int buf = 0;

    for (; u < units; buf ^= 1) {
        const int u_next = u_first + __shfl_sync(0xffffffffu, u_req, 0);
        request();
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncwarp();
        fetch(u_next, buf ^ 1);
        const int e = u >> sh_parts, part = u & ((1 << sh_parts) - 1);
        u = u_next;
        const ScanEnv& se = s_env[buf];
        const int nf = se.nf;
        if (nf < 0) continue;              // masked-out envs and the ones handed to raster_kernel
        const int trow0 = part << sh_unit, row0 = band * band_rows + trow0;     // first row of the unit in the band's tables / in the image
        const int br0 = max(se.r_lo, row0), br1 = min(se.r_hi, row0 + unit_rows - 1);
        uint8_t* obs_e = a.obs + (((size_t)e << (2 * sh_S)) + ((size_t)row0 << sh_S));
        const float* t_nodef = s_nodef + ((size_t)trow0 << sh_S);
        const uint8_t* t_base = s_base + ((size_t)trow0 << sh_S);
        // ---- row intervals of the front faces inside this unit (needs no table: runs under the table fetch in the first round)
        if (br1 >= br0) {
            for (int idx = lane; idx < (nf << sh_unit); idx += 32) {
                const int f = idx >> sh_unit, lr = idx & (unit_rows - 1), r = row0 + lr;
                int lo = 1, hi = 0;
                if (r >= br0 && r <= br1) scan_interval(se.face[f], r, S, lo, hi);
                s_iv[f * unit_rows + lr] = lo <= hi ? make_uchar2((unsigned char)lo, (unsigned char)hi) : make_uchar2(1, 0);
            }
        }
        if (!tables) {
            uint32_t ok = 0;
            while (!ok) {
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                             : "=r"(ok)
                             : "r"(smem_u32(&bar)), "r"(0u)
                             : "memory");
            }
            tables = true;
        }
        if (br1 < br0) {
            // the stimulus does not reach these rows: the baked bytes
            for (int sp = lane; sp < spans; sp += 32) *reinterpret_cast<uint4*>(obs_e + (sp << 4)) = *reinterpret_cast<const uint4*>(t_base + (sp << 4));
            continue;
        }
        __syncwarp();
        // ---- per row: the half spans between the leftmost and the rightmost covered column that have skin pixels
        if (lane < unit_rows) {
            int ulo = S, uhi = -1;
            for (int f = 0; f < nf; f++) {
                const uchar2 iv = s_iv[f * unit_rows + lane];
                if (iv.x <= iv.y) { ulo = min(ulo, (int)iv.x); uhi = max(uhi, (int)iv.y); }
            }
            uint32_t m = 0;
            if (ulo <= uhi) {
                const int hb = ((trow0 + lane) << sh_H);        // first half span of the row in the band's bitmap
                const uint32_t skin = (s_skin[hb >> 5] >> (hb & 31)) & row_all;
                const int h_lo = ulo >> 3, h_hi = uhi >> 3;
                m = skin & (0xffffffffu >> (31 - h_hi)) & (0xffffffffu << h_lo);
            }
            s_rowm[lane] = m;
        }
        __syncwarp();
        // ---- pass A: every 16-pixel span of the unit, one per lane.  8-pixel halves that are not to be shaded get their baked
        // bytes at once; the others are compacted (ballot + popc) into the warp's list so that pass B runs with all lanes busy
        int cnt = 0;
        for (int s0 = 0; s0 < spans; s0 += 32) {
            const int span = s0 + lane;
            uint32_t bits = 0;
            if (span < spans) {
                const int off = span << 4, lr = off >> sh_S, si = span & ((1 << (sh_H - 1)) - 1);
                bits = (s_rowm[lr] >> (2 * si)) & 3u;
                const uint4 bb = *reinterpret_cast<const uint4*>(t_base + off);
                if (bits == 0u) *reinterpret_cast<uint4*>(obs_e + off) = bb;
                else if (bits == 2u) *reinterpret_cast<uint2*>(obs_e + off) = make_uint2(bb.x, bb.y);
                else if (bits == 1u) *reinterpret_cast<uint2*>(obs_e + off + 8) = make_uint2(bb.z, bb.w);
            }
            const uint32_t b0 = __ballot_sync(0xffffffffu, bits & 1u), b1 = __ballot_sync(0xffffffffu, bits & 2u);
            const uint32_t below = (1u << lane) - 1u;
            if (bits & 1u) s_list[cnt + __popc(b0 & below)] = (uint16_t)(2 * span);
            cnt += __popc(b0);
            if (bits & 2u) s_list[cnt + __popc(b1 & below)] = (uint16_t)(2 * span + 1);
            cnt += __popc(b1);
        }
        __syncwarp();
        // ---- pass B: the listed half spans, 32 at a time, one per lane
        for (int i0 = 0; i0 < cnt; i0 += 32) {
            if (i0 + lane >= cnt) continue;
            const int off = (int)s_list[i0 + lane] << 3, lr = off >> sh_S, cb = off & (S - 1), r = row0 + lr;
            const float4 n0 = *reinterpret_cast<const float4*>(t_nodef + off), n1 = *reinterpret_cast<const float4*>(t_nodef + off + 4);
            const float nd[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
            // pen = nodef - d of the NEAREST front face covering the pixel (several parts overlap on the screen; the faces of one
            // part do not): the max over the covering faces of nodef - (float)d, since both roundings are monotonic.  Per face the
            // window depth is affine in the column: d = F - F near / z = (F - Fn w0) - Fn wA k, w = the oracle's 1/z = eA c + eB r + eC
            float pen[8];
#pragma unroll
            for (int k = 0; k < 8; k++) pen[k] = -1.0f;
            for (int f = 0; f < nf; f++) {
                const uchar2 iv = s_iv[f * unit_rows + lr];
                const int l = max((int)iv.x - cb, 0), h = min((int)iv.y - cb, 7);
                if (l > h || iv.x > iv.y) continue;
                const uint32_t m = (0xffu >> (7 - h)) & (0xffu << l);
                const double wA = se.face[f].w[0];
                const double w0 = wA * (double)cb + (se.face[f].w[1] * (double)r + se.face[f].w[2]);
                const double dA = -Fn * wA, d0 = a.F - Fn * w0;
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if ((m >> k) & 1u) pen[k] = fmaxf(pen[k], nd[k] - (float)fma(dA, (double)k, d0));
            }
            const uint2 bres = *reinterpret_cast<const uint2*>(t_base + off);
            uint32_t q8[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                // exact_pixel's arithmetic (tg_raster.cuh) for a covered skin pixel: d as float32, cur = min(nodef, d), then
                // t_s_camera's float32 post-process.  cur - nodef = -(nodef - d) when d < nodef, else 0: pen = nodef - d where that
                // exceeds the 1e-4 dead zone.  Border pixels carry nodef = -1 and d >= 0 (no vertex is nearer than the near
                // plane here), so their pen stays negative and they keep the baked byte; uncovered pixels stay 0.
                const float p = pen[k] > 1e-4f ? fminf(pen[k], 0.05f) : 0.0f;
                const float q0 = __fmul_rn(p, 20.0f);
                const float q = __fmaf_rn(__fmaf_rn(-0.05f, q0, p), 20.0f, q0);
                q8[k] = __float2uint_rz(__fmul_rn(q, 255.0f));
            }
            uint2 o;
            o.x = bres.x | q8[0] | (q8[1] << 8) | (q8[2] << 16) | (q8[3] << 24);
            o.y = bres.y | q8[4] | (q8[5] << 8) | (q8[6] << 16) | (q8[7] << 24);
            *reinterpret_cast<uint2*>(obs_e + off) = o;
        }
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    if (!tables) {       // never leave while the bulk copies into this CTA's shared memory are in flight
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(ok)
                         : "r"(smem_u32(&bar)), "r"(0u)
                         : "memory");
        }
    }
}
