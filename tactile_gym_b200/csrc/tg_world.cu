// tg_world.cu - C ABI of libtactile_gym_b200.so (include/tactile_gym_b200.h): world lifetime, buffers, launches.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: ranges are no-ops unless a tool (nsys, ncu --nvtx) injects itself

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "tg_env.cuh"
#include "tg_g8.cuh"
#include "tg_raster.cuh"
#include "tg_raster_hf.cuh"
#include "tg_raster_scan.cuh"
#include "tg_raster_sphere.cuh"

static thread_local std::string g_err;

// kernels are specialised per arm topology
#define TOPO_DISPATCH(w, CALL)                                         \
    do {                                                               \
        if ((w)->cfg.arm.topo == TG_TOPO_MG400) { using Topo = TopoMG400; CALL; } \
        else { using Topo = TopoChain6; CALL; }                        \
    } while (0)

static int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return fail(TG_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct TgWorld {
    TgConfig cfg;
    int device = 0;
    int n = 0, nb = 0, S = 0;
    int sm_count = 0;
    EnvBuffers eb{};
    RasterArgs ra{};
    struct Alloc { void* p; size_t bytes; };
    std::vector<Alloc> allocs;        // every device buffer of the world, in creation order (tg_checkpoint_* walks this list)
    double* d_draws = nullptr;
    int* d_draw_avail = nullptr;
    uint32_t* d_mt = nullptr;         // device RNG states [N][624] (allocated by the first tg_set_rng_state)
    int* d_mt_pos = nullptr;
    int draw_capacity = 0;            // doubles allocated behind d_draws
    int epoch = 0;                    // bumped by every launch that may touch the standby slots
    unsigned char* d_done_internal = nullptr;
    float* d_reward_internal = nullptr;
    size_t raster_smem = 0;
    // scanline raster (convex parts): tables, the per-env fallback mask + counter, launch shape
    int* d_prim_part = nullptr; double* d_part_cen = nullptr; uint8_t* d_fallback = nullptr; int* d_fb_count = nullptr;
    ScanEnv* d_scan_envs = nullptr; uint32_t* d_skin8 = nullptr; unsigned raster_pass = 0;
    // d_fb_count: [0], [1] envs handed to raster_kernel (even / odd passes), [2 + band * SCAN_MAXPOOLS + pool] the render kernel's unit counters
    // scan_max_pools: unit-counter pools per band - measured no faster than one counter (same-box A/B), kept as the TG_SCAN_POOLS knob
    size_t scan_smem = 0; int scan_grid = 0, scan_lpe = 32, scan_unit_rows = 16, scan_max_pools = 1; bool scan_ok = false;
    size_t push_smem = 0;
    int raster_grid = 0;
    int standby_blocks = 0;
    bool use_g8 = false;              // 8 lanes per env (tg_g8.cuh) instead of one thread per env
    int g8_blocks = 0;
    long long launches = 0;
    // tg_step_host: device staging for the actions, a copy stream and one event per observation chunk
    double* cam_local = nullptr;      // surface_follow-v2 (vertical): the cameras in the heightfield's frame, [N][12]
    float* d_actions_stage = nullptr;
    // tg_step_host's compacted terminal observations: indices + count, image / feature rows (grown on demand; not world state)
    int* d_term_idx = nullptr; uint8_t* d_term_stage = nullptr; float* d_term_feat_stage = nullptr; int term_stage_cap = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t chunk_ev[TG_HOST_MAX_CHUNKS] = {};
    cudaEvent_t step_ev = nullptr, copy_done_ev = nullptr;
};

template <class Tp>
static int dalloc(TgWorld* w, Tp** p, size_t count)
{
    void* v = nullptr;
    cudaError_t e = cudaMalloc(&v, count * sizeof(Tp));
    if (e != cudaSuccess) return fail(TG_ENOMEM, "cudaMalloc(%zu) failed: %s", count * sizeof(Tp), cudaGetErrorString(e));
    cudaMemset(v, 0, count * sizeof(Tp));
    w->allocs.push_back({v, count * sizeof(Tp)});
    *p = static_cast<Tp*>(v);
    return TG_OK;
}

// NVTX range over one C-ABI call: shows up as tg_step / tg_step_host / tg_reset ... on a profiler's timeline
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

extern "C" int tg_version(void) { return TG_VERSION; }
extern "C" const char* tg_last_error(void) { return g_err.c_str(); }

extern "C" int tg_create(const TgConfig* cfg, int device, TgWorld** out)
{
    if (!cfg || !out) return fail(TG_EINVAL, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(TG_ENODEV, "no CUDA device: tactile_gym_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(TG_EINVAL, "device %d out of range (%d devices)", device, ndev);
    if (cfg->n_envs <= 0) return fail(TG_EINVAL, "n_envs must be > 0");
    const int S = cfg->sensor.image_size;
    if (S != 64 && S != 128 && S != 256) return fail(TG_EINVAL, "image_size must be 64, 128 or 256 (got %d)", S);
    if (cfg->arm.topo == TG_TOPO_CHAIN6 && cfg->arm.nb != 6) return fail(TG_EINVAL, "CHAIN6 topology needs nb == 6");
    if (cfg->arm.topo == TG_TOPO_MG400 && cfg->arm.nb != 8) return fail(TG_EINVAL, "MG400 topology needs nb == 8");
    if (cfg->arm.topo != TG_TOPO_CHAIN6 && cfg->arm.topo != TG_TOPO_MG400) return fail(TG_EINVAL, "unknown topology %d", cfg->arm.topo);
    if (cfg->task.task != TG_TASK_EDGE_FOLLOW && cfg->task.task != TG_TASK_OBJECT_BALANCE && cfg->task.task != TG_TASK_SURFACE_FOLLOW &&
        cfg->task.task != TG_TASK_OBJECT_PUSH && cfg->task.task != TG_TASK_OBJECT_ROLL) return fail(TG_EUNSUPPORTED, "task %d not built yet", cfg->task.task);
    if (cfg->task.task == TG_TASK_OBJECT_ROLL) {
        if (cfg->task.push_shape != 1 || !(cfg->task.roll_radius > 0 && cfg->task.obj_mass > 0 && cfg->task.roll_cyl_radius > 0))
            return fail(TG_EINVAL, "object_roll needs push_shape = 1, a marble radius / mass and the tip cylinder");
        if (cfg->task.n_draws != 6) return fail(TG_EINVAL, "object_roll consumes 6 draws per reset");
        if (cfg->arm.topo != TG_TOPO_CHAIN6) return fail(TG_EUNSUPPORTED, "object_roll is built for the UR5 (the reference has no MG400 rest pose for it)");
    }
    if (cfg->task.task == TG_TASK_OBJECT_PUSH) {
        if (!cfg->h_tip_hull || cfg->n_tip_hull <= 0) return fail(TG_EINVAL, "object_push needs the tip core hull (h_tip_hull)");
        if (!(cfg->task.push_half[0] > 0 && cfg->task.push_half[1] > 0 && cfg->task.push_half[2] > 0 && cfg->task.push_inertia_per_mass[0] > 0 &&
              cfg->task.push_inertia_per_mass[1] > 0 && cfg->task.push_inertia_per_mass[2] > 0))
            return fail(TG_EINVAL, "object_push needs a cube with positive extents and inertia");
        if (cfg->task.push_mode < TG_PUSH_WORK || cfg->task.push_mode > TG_PUSH_TCP_TXTYRZ) return fail(TG_EINVAL, "unknown push_mode %d", cfg->task.push_mode);
        if (cfg->task.n_draws != 3) return fail(TG_EINVAL, "object_push consumes 3 draws per reset (init_obj_ang, obj_mass, seed | direction)");
    }
    if (cfg->task.task == TG_TASK_OBJECT_BALANCE && !(cfg->task.obj_mass > 0 && cfg->task.obj_inertia[0] > 0 && cfg->task.obj_inertia[1] > 0 && cfg->task.obj_inertia[2] > 0))
        return fail(TG_EINVAL, "object_balance needs a free object with positive mass and inertia");
    if (cfg->task.n_draws < 0 || cfg->task.n_draws > TG_MAXDRAW) return fail(TG_EINVAL, "n_draws must be in 0..%d", TG_MAXDRAW);
    if (cfg->task.task == TG_TASK_SURFACE_FOLLOW || cfg->task.task == TG_TASK_OBJECT_ROLL) {
        if (cfg->sensor.n_prim != 0) return fail(TG_EINVAL, "surface_follow / object_roll draw a per-env heightfield / sphere: n_prim must be 0");
    } else if (cfg->sensor.n_prim <= 0 || cfg->sensor.n_prim > RASTER_MAXPRIM) return fail(TG_EINVAL, "n_prim must be in 1..%d", RASTER_MAXPRIM);
    if (!cfg->sensor.h_nodef_dep || !cfg->sensor.h_nodef_gray || !cfg->sensor.h_border_mask || !cfg->h_rest_q ||
        (cfg->sensor.n_prim > 0 && (!cfg->sensor.h_prims || !cfg->sensor.h_prim_nv)))
        return fail(TG_EINVAL, "null table pointer in config");
    CK(cudaSetDevice(device));

    TgWorld* w = new TgWorld();
    struct Guard { TgWorld* w; ~Guard() { if (w) tg_destroy(w); } } guard{w};   // every early return below frees the partial world
    w->cfg = *cfg;
    w->device = device;
    w->n = cfg->n_envs; w->nb = cfg->arm.nb; w->S = S;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    w->sm_count = prop.multiProcessorCount;
    const int n = w->n, nb = w->nb;
    int lanes = cfg->lanes_per_warp;
    {
        // the 8-lanes-per-env kernel: motor-row tasks under TCP_velocity_control with gravity compensation.  One warp carries 4
        // envs (against up to 32 of the one-thread kernel), so it wins while the device is not yet full of warps: measured on
        // B200 (edge_follow, L2 flushed) 0.180 against 0.267 ms at N = 4096, 0.294 against 0.276 ms at N = 8192
        const bool can = (cfg->task.task == TG_TASK_EDGE_FOLLOW || cfg->task.task == TG_TASK_SURFACE_FOLLOW ||
                          (cfg->task.task == TG_TASK_OBJECT_BALANCE && cfg->arm.topo == TG_TOPO_CHAIN6)) &&
                         cfg->task.control_mode == 0 && cfg->phys.gravity_comp == 1;
        if (lanes == -8 && !can) return fail(TG_EUNSUPPORTED, "lanes_per_warp = -8 (8 lanes per env) is built for edge_follow / surface_follow under TCP_velocity_control");
        w->use_g8 = can && (lanes == -8 || (lanes == 0 && n <= 6144));
        if (const char* ev = getenv("TG_G8")) w->use_g8 = can && atoi(ev) != 0;     // tuning hook
        w->g8_blocks = (n + 15) / 16;
    }
    if (lanes != 1 && lanes != 2 && lanes != 4 && lanes != 8 && lanes != 16 && lanes != 32) {
        // The per-env work is one long dependent fp64 chain and a warp costs the same issue slots whether 1 or
        // 32 lanes are active, so the kernel time is the chain latency as long as every SM sub-partition has
        // about one warp: give each of the 4 x #SM schedulers a warp before packing more lanes into a warp
        // (measured on B200, N = 4096: 1 lane/warp 1.28 ms, 2: 0.66, 4: 0.34, 8: 0.22, 32: 0.22).
        const long target_warps = (long)w->sm_count * 4;
        lanes = 32;
        while (lanes > 1 && (long)(n + lanes / 2 - 1) / (lanes / 2) <= target_warps) lanes >>= 1;
    }
    EnvBuffers& b = w->eb;
    b.n = n; b.lanes = lanes;
    int rc;
    double* rest = nullptr;
    if ((rc = dalloc(w, &b.q, (size_t)nb * n)) || (rc = dalloc(w, &b.qd, (size_t)nb * n)) || (rc = dalloc(w, &b.embed, n)) ||
        (rc = dalloc(w, &b.edge_ang, n)) || (rc = dalloc(w, &b.steps, n)) || (rc = dalloc(w, &b.reset_substeps, n)) ||
        (rc = dalloc(w, &b.reset_count, n)) || (rc = dalloc(w, &b.cam, (size_t)12 * n)) || (rc = dalloc(w, &b.stim, (size_t)12 * n)) ||
        (rc = dalloc(w, &b.tcp, (size_t)7 * n)) || (rc = dalloc(w, &rest, nb)) || (rc = dalloc(w, &w->d_done_internal, n)) ||
        (rc = dalloc(w, &w->d_reward_internal, n))) {
        return rc;
    }
    CK(cudaMemcpy(rest, cfg->h_rest_q, sizeof(double) * nb, cudaMemcpyHostToDevice));
    b.rest_q = rest;
    b.draws = nullptr; b.draw_rounds = 0;
    // standby reset pipeline (tg_env.cuh): needs episodes of at least 2 steps
    b.pipeline = cfg->task.max_steps >= 2 ? 1 : 0;
    {
        // Quantum sizes of the resumable reset.  A standby thread's quantum runs BESIDE the step warps of the same launch, as one
        // thread's dependent fp64 chain: it must stay well below the step's own duration or it IS the launch's duration (measured,
        // edge_follow 4096 envs: 6 move substeps + 8 IK iterations per quantum -> 0.324 ms per vec-step, 2 + 3 -> 0.232 ms).  So
        // the quanta are as small as the episode length allows: the rebuild (<= 100 IK iterations, then a move of typically
        // 7 - 110 substeps; surface_follow: 4,096 heights first) has to finish within about half an episode, otherwise the env
        // completes it inline when it needs it (exact either way, counted by tg_pipeline_stalls).
        // object_balance's episodes end when the pole falls, typically after a few dozen steps whatever max_steps says: it keeps
        // the larger quanta (measured at config 5: 0.383 ms per vec-step with 6 / 8, 0.407 ms with 2 / 3 - the rebuilds were late).
        const double ms = cfg->task.task == TG_TASK_OBJECT_BALANCE ? 60.0 : std::max(2, cfg->task.max_steps);
        b.ik_chunk = std::min(100, std::max(3, (int)ceil(100.0 / (0.25 * ms))));
        b.reset_chunk = std::min(1000, std::max(2, (int)ceil(120.0 / (0.3 * ms))));
        b.surf_chunk = SURF_CHUNK;
    }
    if (const char* ev = getenv("TG_SURF_CHUNK")) b.surf_chunk = std::max(1, atoi(ev));
    if (const char* ev = getenv("TG_IK_CHUNK")) b.ik_chunk = std::max(1, atoi(ev));       // tuning hooks
    if (const char* ev = getenv("TG_RESET_CHUNK")) b.reset_chunk = std::max(1, atoi(ev));
    b.step_blocks = (n + 4 * lanes - 1) / (4 * lanes);
    if ((rc = dalloc(w, &b.sb_q, (size_t)nb * n)) || (rc = dalloc(w, &b.sb_qd, (size_t)nb * n)) || (rc = dalloc(w, &b.sb_embed, n)) ||
        (rc = dalloc(w, &b.sb_ang, n)) || (rc = dalloc(w, &b.sb_cam, (size_t)12 * n)) || (rc = dalloc(w, &b.sb_stim, (size_t)12 * n)) ||
        (rc = dalloc(w, &b.sb_tcp, (size_t)7 * n)) || (rc = dalloc(w, &b.sb_substeps, n)) || (rc = dalloc(w, &b.sb_ready, n)) ||
        (rc = dalloc(w, &b.term_cam, (size_t)12 * n)) || (rc = dalloc(w, &b.term_stim, (size_t)12 * n)) || (rc = dalloc(w, &b.error_flag, 1)) ||
        (rc = dalloc(w, &b.stall_count, 1)) || (rc = dalloc(w, &b.nan_count, 1)) || (rc = dalloc(w, &b.sb_targ, (size_t)nb * n)) || (rc = dalloc(w, &b.sb_cv, n)) || (rc = dalloc(w, &b.sb_ik, n)) ||
        (rc = dalloc(w, &b.sb_draw, (size_t)TG_MAXDRAW * n))) {
        return rc;
    }
    w->standby_blocks = b.pipeline ? std::max(8, std::min(64, w->sm_count / 2)) : 0;
    if (cfg->task.task == TG_TASK_OBJECT_ROLL) {
        w->standby_blocks = 0;
        w->push_smem = sizeof(double) * PushLayout<TopoChain6>::SLOTS * PUSH_BLOCK;
        CK(cudaFuncSetAttribute(step_kernel<TopoChain6, TG_TASK_OBJECT_ROLL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w->push_smem));
        CK(cudaFuncSetAttribute(step_kernel<TopoChain6, TG_TASK_OBJECT_ROLL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w->push_smem));
    }
    if (cfg->task.task == TG_TASK_OBJECT_PUSH) {
        // object_push steps PUSH_BLOCK envs per block with its constraint rows in dynamic shared memory (tg_push.cuh);
        // standby rebuilds ride in the step threads, not in extra blocks
        w->standby_blocks = 0;
        if (cfg->arm.topo == TG_TOPO_MG400) {
            w->push_smem = sizeof(double) * PushLayout<TopoMG400>::SLOTS * PUSH_BLOCK;
            CK(cudaFuncSetAttribute(step_kernel<TopoMG400, TG_TASK_OBJECT_PUSH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w->push_smem));
            CK(cudaFuncSetAttribute(step_kernel<TopoMG400, TG_TASK_OBJECT_PUSH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w->push_smem));
        } else {
            w->push_smem = sizeof(double) * PushLayout<TopoChain6>::SLOTS * PUSH_BLOCK;
            CK(cudaFuncSetAttribute(step_kernel<TopoChain6, TG_TASK_OBJECT_PUSH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w->push_smem));
            CK(cudaFuncSetAttribute(step_kernel<TopoChain6, TG_TASK_OBJECT_PUSH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w->push_smem));
        }
    }
    if (cfg->task.task == TG_TASK_SURFACE_FOLLOW) {
        if (!b.pipeline) { return fail(TG_EINVAL, "surface_follow needs max_steps >= 2"); }
        if ((rc = dalloc(w, &b.height, (size_t)2 * SURF_PTS * n)) || (rc = dalloc(w, &b.hf_meta, (size_t)2 * SURF_META * n)) ||
            (rc = dalloc(w, &b.hf_cur, n)) || (rc = dalloc(w, &b.sb_perm, (size_t)256 * n)) || (rc = dalloc(w, &b.sb_surf_it, n)) ||
            (rc = dalloc(w, &b.sb_hmm, (size_t)2 * n)) || (rc = dalloc(w, &b.accum, n)) || (rc = dalloc(w, &w->cam_local, (size_t)12 * n))) {
            return rc;
        }
    }
    if (cfg->task.task == TG_TASK_OBJECT_ROLL) {
        if ((rc = dalloc(w, &b.traj, (size_t)PUSH_TRAJ_SZ * n)) || (rc = dalloc(w, &b.sb_traj, (size_t)PUSH_TRAJ_SZ * n)) || (rc = dalloc(w, &b.goal, n)) ||
            (rc = dalloc(w, &b.sb_goal, n))) {
            return rc;
        }
        b.hull = nullptr; b.n_hull = 0;
    }
    if (cfg->task.task == TG_TASK_OBJECT_PUSH) {
        double* hull = nullptr;
        if ((rc = dalloc(w, &b.traj, (size_t)PUSH_TRAJ_SZ * n)) || (rc = dalloc(w, &b.sb_traj, (size_t)PUSH_TRAJ_SZ * n)) || (rc = dalloc(w, &b.goal, n)) ||
            (rc = dalloc(w, &b.sb_goal, n)) || (rc = dalloc(w, &hull, (size_t)3 * cfg->n_tip_hull))) {
            return rc;
        }
        CK(cudaMemcpy(hull, cfg->h_tip_hull, sizeof(double) * 3 * cfg->n_tip_hull, cudaMemcpyHostToDevice));
        b.hull = hull; b.n_hull = cfg->n_tip_hull;
    }
    if (cfg->task.task == TG_TASK_OBJECT_BALANCE || cfg->task.task == TG_TASK_OBJECT_PUSH || cfg->task.task == TG_TASK_OBJECT_ROLL) {
        if ((rc = dalloc(w, &b.obj, (size_t)13 * n)) || (rc = dalloc(w, &b.obj_ext, (size_t)4 * n)) || (rc = dalloc(w, &b.grav, n)) ||
            (rc = dalloc(w, &b.sb_obj, (size_t)13 * n)) || (rc = dalloc(w, &b.sb_obj_ext, (size_t)4 * n)) || (rc = dalloc(w, &b.sb_grav, n))) {
            return rc;
        }
    }

    // raster tables: border pixels get nodef = -1 and the baked grey value (tactile_sensor.py:289-292)
    {
        const size_t px = (size_t)S * S;
        std::vector<float> nd(px);
        std::vector<uint8_t> base(px);
        for (size_t i = 0; i < px; i++) {
            const bool border = cfg->sensor.border_on && cfg->sensor.h_border_mask[i] == 1;
            nd[i] = border ? -1.0f : cfg->sensor.h_nodef_dep[i];
            base[i] = border ? (uint8_t)cfg->sensor.h_nodef_gray[i] : 0;
        }
        float* dn; uint8_t* db; double* dt; int* dnv;
        const int np = cfg->sensor.n_prim;
        for (int i = 0; i < np; i++)
            if (cfg->sensor.h_prim_nv[i] != 3 && cfg->sensor.h_prim_nv[i] != 4) { return fail(TG_EINVAL, "primitive %d has %d vertices", i, cfg->sensor.h_prim_nv[i]); }
        if ((rc = dalloc(w, &dn, px)) || (rc = dalloc(w, &db, px)) || (rc = dalloc(w, &dt, (size_t)std::max(np, 1) * 12)) || (rc = dalloc(w, &dnv, std::max(np, 1)))) { return rc; }
        CK(cudaMemcpy(dn, nd.data(), px * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(db, base.data(), px, cudaMemcpyHostToDevice));
        if (np > 0) {
            CK(cudaMemcpy(dt, cfg->sensor.h_prims, sizeof(double) * 12 * np, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(dnv, cfg->sensor.h_prim_nv, sizeof(int) * np, cudaMemcpyHostToDevice));
        }
        float ndmin = 1e30f, ndmax = -1e30f;
        for (size_t i = 0; i < px; i++) if (nd[i] >= 0.0f) { ndmin = std::min(ndmin, nd[i]); ndmax = std::max(ndmax, nd[i]); }
        RasterArgs& r = w->ra;
        r.nd_ref = ndmin <= ndmax ? 0.5f * (ndmin + ndmax) : 0.5f;
        if (ndmin <= ndmax && !(ndmin >= 0.5f * r.nd_ref && ndmax <= 2.0f * r.nd_ref)) { return fail(TG_EUNSUPPORTED, "nodef_dep range [%g, %g] too wide for the float path", ndmin, ndmax); }
        r.n = n; r.S = S; r.bands = S == 256 ? 4 : 1; r.nprim = np; r.prim_nv = dnv;
        r.th = tan(cfg->sensor.fov_deg * (M_PI / 180.0) / 2.0);
        r.near_ = cfg->sensor.near_; r.far_ = cfg->sensor.far_;
        r.F = cfg->sensor.far_ / (cfg->sensor.far_ - cfg->sensor.near_);
        r.nodef = dn; r.base = db; r.prims = dt; r.cam = b.cam; r.stim = b.stim; r.mask = nullptr; r.obs = nullptr;
        const size_t band_px = px / r.bands;
        w->raster_smem = band_px * 5 + ((band_px / 16 + 31) / 32) * 4 + 16 + RASTER_PER_WARP_SMEM * RASTER_WARPS;
        CK(cudaFuncSetAttribute(raster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w->raster_smem));
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, raster_kernel, RASTER_THREADS, w->raster_smem));
        if (per_sm < 1) { return fail(TG_ECUDA, "raster kernel does not fit on an SM (smem %zu)", w->raster_smem); }
        int grid = w->sm_count * per_sm;
        grid -= grid % r.bands;
        const int need = ((n + RASTER_WARPS - 1) / RASTER_WARPS) * r.bands;
        if (grid > need) grid = need;
        w->raster_grid = grid;
        r.hf = nullptr; r.hf_cur = nullptr; r.hf_meta = nullptr; r.hf_flip = 0; r.hf_tile_rows = 16; r.hf_tile_cols = 32;
        r.scan_test_fallback = getenv("TG_SCAN_TEST_FALLBACK") ? 1 : 0;
        if (np > 0 && cfg->sensor.n_parts > 0 && cfg->sensor.n_parts <= SCAN_MAXPARTS && cfg->sensor.h_prim_part && cfg->sensor.h_part_centroid && !getenv("TG_NO_SCAN")) {
            // convex parts: the scanline raster renders, raster_kernel only takes the envs it flags
            for (int i = 0; i < np; i++)
                if (cfg->sensor.h_prim_part[i] < 0 || cfg->sensor.h_prim_part[i] >= cfg->sensor.n_parts) return fail(TG_EINVAL, "h_prim_part[%d] out of range", i);
            if ((rc = dalloc(w, &w->d_prim_part, np)) || (rc = dalloc(w, &w->d_part_cen, (size_t)3 * cfg->sensor.n_parts)) ||
                (rc = dalloc(w, &w->d_fallback, n)) || (rc = dalloc(w, &w->d_fb_count, 2 + 4 * SCAN_MAXPOOLS)) || (rc = dalloc(w, &w->d_scan_envs, n))) return rc;
            CK(cudaMemcpy(w->d_prim_part, cfg->sensor.h_prim_part, sizeof(int) * np, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(w->d_part_cen, cfg->sensor.h_part_centroid, sizeof(double) * 3 * cfg->sensor.n_parts, cudaMemcpyHostToDevice));
            // band tables + half-span skin bitmap + the warps' own tables
            // rows per work unit: 32 at <= 128 x 128 (fewer, fatter units: - 3.5 % at 4096 / 8192 envs), 16 for the 64-row bands of 256 x 256
            // (32 there: + 8 %); same-box A/B, tools/raster_time.py
            w->scan_unit_rows = S <= 128 ? 32 : 16;
            if (const char* ur = getenv("TG_SCAN_UNIT_ROWS")) w->scan_unit_rows = atoi(ur) == 32 ? 32 : 16;   // experiment knob
            if (const char* mp = getenv("TG_SCAN_POOLS")) w->scan_max_pools = std::max(1, std::min(SCAN_MAXPOOLS, atoi(mp)));
            const int unit_rows = std::min(w->scan_unit_rows, S / r.bands);
            w->scan_smem = scan_tables_smem((int)band_px) + scan_per_warp_smem(S, unit_rows) * SCAN_WARPS;
            {
                // 1 bit per 8-pixel half span, band by band: has a non-border pixel
                const size_t words = (band_px / 8 + 31) / 32;
                std::vector<uint32_t> bits(words * r.bands, 0u);
                for (size_t i = 0; i < px; i++)
                    if (nd[i] >= 0.0f) { const size_t bnd = i / band_px, h = (i % band_px) / 8; bits[bnd * words + h / 32] |= 1u << (h % 32); }
                if ((rc = dalloc(w, &w->d_skin8, bits.size()))) return rc;
                CK(cudaMemcpy(w->d_skin8, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice));
            }
            w->scan_lpe = np <= 8 ? 8 : (np <= 16 ? 16 : 32);
            int ps = 0;
            CK(cudaFuncSetAttribute(raster_scan_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w->scan_smem));
            CK(cudaFuncSetAttribute(raster_scan_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w->scan_smem));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ps, raster_scan_kernel<4>, SCAN_THREADS, w->scan_smem));
            if (ps >= 1) {
                w->scan_grid = w->sm_count * ps;
                w->scan_grid -= w->scan_grid % r.bands;   // the same number of CTAs on every band
                w->scan_ok = true;
            }
        }
        if (cfg->task.task == TG_TASK_SURFACE_FOLLOW) {
            // heightfield stimulus: per-tile primitive lists (tg_raster_hf.cuh).  Tile size: its footprint on the surface
            // (at the deepest skin depth) should stay within ~1.5 grid cells so that a tile sees <= 32 triangles
            r.hf = b.height; r.hf_cur = b.hf_cur; r.hf_meta = b.hf_meta;
            for (int c = 0; c < 3; c++) r.surf_pos[c] = cfg->task.surf_pos[c];
            r.surf_grid = cfg->task.surf_grid;
            const double zmax = r.near_ * r.F / (r.F - (double)ndmax), pix = 2.0 * zmax * r.th / S, lim = 1.5 * r.surf_grid;
            r.hf_tile_cols = 32 * pix <= lim ? 32 : 16;
            r.hf_tile_rows = 16 * pix <= lim ? 16 : (8 * pix <= lim ? 8 : 4);
            w->raster_smem = ((band_px * 5 + 15) & ~size_t(15)) + HF_PER_WARP_SMEM * HF_WARPS;
            CK(cudaFuncSetAttribute(raster_hf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w->raster_smem));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, raster_hf_kernel, HF_THREADS, w->raster_smem));
            if (per_sm < 1) { return fail(TG_ECUDA, "heightfield raster kernel does not fit on an SM (smem %zu)", w->raster_smem); }
            grid = w->sm_count * per_sm;
            grid -= grid % r.bands;
            const int need_hf = ((n * HF_UPE + HF_WARPS - 1) / HF_WARPS) * r.bands;
            if (grid > need_hf) grid = need_hf;
            w->raster_grid = grid;
        }
    }
    if ((rc = dalloc(w, &w->d_draw_avail, n))) return rc;
    b.draw_avail = w->d_draw_avail;
    guard.w = nullptr;
    *out = w;
    return TG_OK;
}

extern "C" int tg_destroy(TgWorld* w)
{
    if (!w) return TG_OK;
    cudaSetDevice(w->device);
    for (const TgWorld::Alloc& a : w->allocs) cudaFree(a.p);
    if (w->d_draws) cudaFree(w->d_draws);
    if (w->d_term_idx) cudaFree(w->d_term_idx);
    if (w->d_term_stage) cudaFree(w->d_term_stage);
    if (w->d_term_feat_stage) cudaFree(w->d_term_feat_stage);
    if (w->copy_stream) cudaStreamDestroy(w->copy_stream);
    for (cudaEvent_t ev : w->chunk_ev) if (ev) cudaEventDestroy(ev);
    if (w->step_ev) cudaEventDestroy(w->step_ev);
    if (w->copy_done_ev) cudaEventDestroy(w->copy_done_ev);
    delete w;
    return TG_OK;
}

// Reset draws live in a device ring [N][rounds][n_draws]: the k-th reset of env e since tg_set_draws reads slot k % rounds and
// needs k < draw_avail[e].  tg_set_draws starts a sequence (synchronous); tg_draws_poll / tg_draws_upload keep it fed without
// ever synchronising (the host overwrites only slots whose draws were consumed).
extern "C" int tg_set_draws(TgWorld* w, const double* h_draws, int rounds)
{
    if (!w || !h_draws || rounds <= 0) return fail(TG_EINVAL, "bad arguments");
    if (w->cfg.task.task == TG_TASK_SURFACE_FOLLOW && w->cfg.task.surf_mode == 4)
        return fail(TG_EUNSUPPORTED, "noise_mode 'random' draws 1,024 heights per reset: it runs on the device RNG (tg_set_rng_state) only");
    CK(cudaSetDevice(w->device));
    CK(cudaDeviceSynchronize());
    const size_t cnt = (size_t)w->n * rounds * w->cfg.task.n_draws;
    if ((size_t)w->draw_capacity < cnt) {
        if (w->d_draws) { cudaFree(w->d_draws); w->d_draws = nullptr; w->draw_capacity = 0; }
        CK(cudaMalloc(&w->d_draws, cnt * sizeof(double)));
        w->draw_capacity = (int)cnt;
    }
    CK(cudaMemcpy(w->d_draws, h_draws, cnt * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemset(w->eb.reset_count, 0, sizeof(int) * w->n));
    std::vector<int> avail(w->n, rounds);
    CK(cudaMemcpy(w->d_draw_avail, avail.data(), sizeof(int) * w->n, cudaMemcpyHostToDevice));
    w->eb.draws = w->d_draws; w->eb.draw_rounds = rounds;
    w->eb.mt_active = 0;
    {
        int flag = 0;   // a new sequence: forget that the old one ran dry
        CK(cudaMemcpy(&flag, w->eb.error_flag, sizeof(int), cudaMemcpyDeviceToHost));
        flag &= ~2;
        CK(cudaMemcpy(w->eb.error_flag, &flag, sizeof(int), cudaMemcpyHostToDevice));
    }
    if (w->eb.pipeline) {
        // a new draw sequence starts: standbys computed from the old one are recomputed now
        CK(cudaMemset(w->eb.sb_ready, 0, sizeof(int) * w->n));
        w->eb.epoch = ++w->epoch;
        TOPO_DISPATCH(w, (standby_kernel<Topo><<<(w->n + 127) / 128, 128>>>(w->cfg.arm, w->cfg.phys, w->cfg.task, w->eb)));
        w->launches++;
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
    }
    return TG_OK;
}

extern "C" int tg_set_rng_state(TgWorld* w, const uint32_t* h_key, const int32_t* h_pos)
{
    if (!w || !h_key || !h_pos) return fail(TG_EINVAL, "bad arguments");
    CK(cudaSetDevice(w->device));
    CK(cudaDeviceSynchronize());
    for (int d = 0; d < w->cfg.task.n_draws; d++)
        if (w->cfg.task.draw_kind[d] < TG_DRAW_CONST || w->cfg.task.draw_kind[d] > TG_DRAW_CHOICE_RAND) return fail(TG_EINVAL, "draw_kind[%d] = %d", d, w->cfg.task.draw_kind[d]);
    if (!w->d_mt) {
        int rc;
        if ((rc = dalloc(w, &w->d_mt, (size_t)w->n * MT_N)) || (rc = dalloc(w, &w->d_mt_pos, w->n))) return rc;
    }
    CK(cudaMemcpy(w->d_mt, h_key, sizeof(uint32_t) * MT_N * w->n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(w->d_mt_pos, h_pos, sizeof(int) * w->n, cudaMemcpyHostToDevice));
    CK(cudaMemset(w->eb.reset_count, 0, sizeof(int) * w->n));
    w->eb.mt = w->d_mt; w->eb.mt_pos = w->d_mt_pos; w->eb.mt_active = 1;
    {
        int flag = 0;
        CK(cudaMemcpy(&flag, w->eb.error_flag, sizeof(int), cudaMemcpyDeviceToHost));
        flag &= ~2;
        CK(cudaMemcpy(w->eb.error_flag, &flag, sizeof(int), cudaMemcpyHostToDevice));
    }
    if (w->eb.pipeline) {
        CK(cudaMemset(w->eb.sb_ready, 0, sizeof(int) * w->n));
        w->eb.epoch = ++w->epoch;
        TOPO_DISPATCH(w, (standby_kernel<Topo><<<(w->n + 127) / 128, 128>>>(w->cfg.arm, w->cfg.phys, w->cfg.task, w->eb)));
        w->launches++;
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
    }
    return TG_OK;
}

extern "C" int tg_draws_poll(TgWorld* w, int32_t* h_counts, void* stream)
{
    if (!w || !h_counts) return fail(TG_EINVAL, "bad arguments");
    CK(cudaSetDevice(w->device));
    CK(cudaMemcpyAsync(h_counts, w->eb.reset_count, sizeof(int) * w->n, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaMemcpyAsync(h_counts + w->n, w->eb.error_flag, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return TG_OK;
}

extern "C" int tg_draws_upload(TgWorld* w, const double* h_ring, const int32_t* h_avail, void* stream)
{
    if (!w || !h_ring || !h_avail) return fail(TG_EINVAL, "bad arguments");
    if (!w->d_draws || w->eb.draw_rounds <= 0) return fail(TG_EINVAL, "tg_draws_upload needs a ring started by tg_set_draws");
    CK(cudaSetDevice(w->device));
    const size_t cnt = (size_t)w->n * w->eb.draw_rounds * w->cfg.task.n_draws;
    // ring first, counters second: a reset that sees the new draw_avail finds the new draws (same stream, in order; slots
    // already valid are rewritten with identical values)
    CK(cudaMemcpyAsync(w->d_draws, h_ring, cnt * sizeof(double), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    CK(cudaMemcpyAsync(w->d_draw_avail, h_avail, sizeof(int) * w->n, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return TG_OK;
}

extern "C" int tg_pipeline_error(TgWorld* w, void* stream)
{
    if (!w) return fail(TG_EINVAL, "bad arguments");
    CK(cudaSetDevice(w->device));
    int flag = 0;
    CK(cudaMemcpyAsync(&flag, w->eb.error_flag, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return flag;   // bit 0: reset pipeline / raster; bit 1: a reset found its draws exhausted
}

extern "C" int tg_pipeline_stalls(TgWorld* w, void* stream)
{
    if (!w) return fail(TG_EINVAL, "bad arguments");
    CK(cudaSetDevice(w->device));
    int cnt = 0;
    CK(cudaMemcpyAsync(&cnt, w->eb.stall_count, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return cnt;
}

extern "C" int tg_scan_fallbacks(TgWorld* w, void* stream)
{
    if (!w) return fail(TG_EINVAL, "bad arguments");
    if (!w->scan_ok || w->raster_pass == 0) return 0;
    CK(cudaSetDevice(w->device));
    int cnt = 0;
    CK(cudaMemcpyAsync(&cnt, w->d_fb_count + ((w->raster_pass - 1u) & 1u), sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return cnt;
}

extern "C" int tg_scan_fallback_reasons(TgWorld* w, int32_t* h_counts, void* stream)
{
    if (!w || !h_counts) return fail(TG_EINVAL, "bad arguments");
    for (int i = 0; i < 8; i++) h_counts[i] = 0;
    if (!w->scan_ok) return TG_OK;
    CK(cudaSetDevice(w->device));
    std::vector<uint8_t> fb(w->n);
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    CK(cudaMemcpy(fb.data(), w->d_fallback, w->n, cudaMemcpyDeviceToHost));
    for (int e = 0; e < w->n; e++) h_counts[fb[e] & 7]++;
    return TG_OK;
}

extern "C" int tg_nan_resets(TgWorld* w, void* stream)
{
    if (!w) return fail(TG_EINVAL, "bad arguments");
    CK(cudaSetDevice(w->device));
    int cnt = 0;
    CK(cudaMemcpyAsync(&cnt, w->eb.nan_count, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return cnt;
}

// ---- full checkpoint: every device buffer of the world (live state, standby slots and partial rebuilds, both heightfields,
// RNG states, draw ring, counters) plus the host-side launch counters the kernels are handed.  The caller's own output
// buffers (obs / reward / done / features) are not part of the world; the Python layer saves those it needs.
struct CkptHeader {
    uint32_t magic, version;
    int32_t n, nb, task, S, n_allocs, draw_rounds;
    int64_t draw_doubles;
    int32_t epoch; uint32_t raster_pass;
    int64_t launches;
};
static const uint32_t CKPT_MAGIC = 0x54474350u; // "TGCP"

extern "C" size_t tg_checkpoint_bytes(const TgWorld* w)
{
    if (!w) return 0;
    size_t tot = sizeof(CkptHeader) + sizeof(uint64_t) * w->allocs.size();
    for (const TgWorld::Alloc& a : w->allocs) tot += a.bytes;
    if (w->d_draws && w->eb.draw_rounds > 0) tot += sizeof(double) * (size_t)w->n * w->eb.draw_rounds * w->cfg.task.n_draws;
    return tot;
}

extern "C" int tg_checkpoint_save(TgWorld* w, void* host, size_t bytes, void* stream)
{
    NvtxRange nvtx_("tg_checkpoint_save");
    if (!w || !host) return fail(TG_EINVAL, "bad arguments");
    if (bytes < tg_checkpoint_bytes(w)) return fail(TG_EINVAL, "checkpoint buffer too small: %zu < %zu", bytes, tg_checkpoint_bytes(w));
    CK(cudaSetDevice(w->device));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    if (w->copy_stream) CK(cudaStreamSynchronize(w->copy_stream));
    unsigned char* o = static_cast<unsigned char*>(host);
    CkptHeader h{};
    h.magic = CKPT_MAGIC; h.version = (uint32_t)TG_VERSION; h.n = w->n; h.nb = w->nb; h.task = w->cfg.task.task; h.S = w->S;
    h.n_allocs = (int32_t)w->allocs.size(); h.draw_rounds = w->d_draws ? w->eb.draw_rounds : 0;
    h.draw_doubles = h.draw_rounds > 0 ? (int64_t)w->n * h.draw_rounds * w->cfg.task.n_draws : 0;
    h.epoch = w->epoch; h.raster_pass = w->raster_pass; h.launches = w->launches;
    memcpy(o, &h, sizeof(h)); o += sizeof(h);
    for (const TgWorld::Alloc& a : w->allocs) { const uint64_t b = a.bytes; memcpy(o, &b, sizeof(b)); o += sizeof(b); }
    for (const TgWorld::Alloc& a : w->allocs) { CK(cudaMemcpy(o, a.p, a.bytes, cudaMemcpyDeviceToHost)); o += a.bytes; }
    if (h.draw_doubles > 0) CK(cudaMemcpy(o, w->d_draws, sizeof(double) * (size_t)h.draw_doubles, cudaMemcpyDeviceToHost));
    return TG_OK;
}

extern "C" int tg_checkpoint_load(TgWorld* w, const void* host, size_t bytes, void* stream)
{
    NvtxRange nvtx_("tg_checkpoint_load");
    if (!w || !host || bytes < sizeof(CkptHeader)) return fail(TG_EINVAL, "bad arguments");
    CK(cudaSetDevice(w->device));
    const unsigned char* o = static_cast<const unsigned char*>(host);
    CkptHeader h;
    memcpy(&h, o, sizeof(h)); o += sizeof(h);
    if (h.magic != CKPT_MAGIC || h.version != (uint32_t)TG_VERSION) return fail(TG_EINVAL, "not a checkpoint of this library version");
    if (h.n != w->n || h.nb != w->nb || h.task != w->cfg.task.task || h.S != w->S || h.n_allocs != (int32_t)w->allocs.size())
        return fail(TG_EINVAL, "checkpoint of a different world (envs %d / %d, task %d / %d, image %d / %d, buffers %d / %zu)", h.n, w->n, h.task,
                    w->cfg.task.task, h.S, w->S, h.n_allocs, w->allocs.size());
    size_t need = sizeof(CkptHeader) + sizeof(uint64_t) * w->allocs.size() + sizeof(double) * (size_t)h.draw_doubles;
    for (size_t i = 0; i < w->allocs.size(); i++) {
        uint64_t b;
        memcpy(&b, o + sizeof(uint64_t) * i, sizeof(b));
        if (b != w->allocs[i].bytes) return fail(TG_EINVAL, "checkpoint buffer %zu has %llu bytes, the world's has %zu", i, (unsigned long long)b, w->allocs[i].bytes);
        need += b;
    }
    if (bytes < need) return fail(TG_EINVAL, "checkpoint truncated: %zu < %zu", bytes, need);
    o += sizeof(uint64_t) * w->allocs.size();
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    if (w->copy_stream) CK(cudaStreamSynchronize(w->copy_stream));
    for (const TgWorld::Alloc& a : w->allocs) { CK(cudaMemcpy(a.p, o, a.bytes, cudaMemcpyHostToDevice)); o += a.bytes; }
    if (h.draw_doubles > 0) {
        if (w->draw_capacity < h.draw_doubles) {
            if (w->d_draws) cudaFree(w->d_draws);
            w->d_draws = nullptr; w->draw_capacity = 0;
            CK(cudaMalloc(&w->d_draws, sizeof(double) * (size_t)h.draw_doubles));
            w->draw_capacity = (int)h.draw_doubles;
        }
        CK(cudaMemcpy(w->d_draws, o, sizeof(double) * (size_t)h.draw_doubles, cudaMemcpyHostToDevice));
        w->eb.draws = w->d_draws; w->eb.draw_rounds = h.draw_rounds;
    }
    w->epoch = h.epoch; w->eb.epoch = h.epoch; w->raster_pass = h.raster_pass; w->launches = h.launches;
    return TG_OK;
}

extern "C" int tg_get_reset_counts(TgWorld* w, int32_t* h_counts, void* stream)
{
    if (!w || !h_counts) return fail(TG_EINVAL, "bad arguments");
    CK(cudaSetDevice(w->device));
    CK(cudaMemcpyAsync(h_counts, w->eb.reset_count, sizeof(int) * w->n, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return TG_OK;
}

static dim3 env_grid(const TgWorld* w) { return dim3(w->eb.step_blocks); } // 128 threads = 4 warps = 4 * lanes envs

// surface_follow-v2's upright heightfield: depth along a camera ray does not change under a rigid motion, so instead of rotating the
// 7,938-triangle surface the camera is brought into the heightfield's own frame (about surf_pos; world (a, b, c) -> local (c, b, -a),
// the inverse of euler (0, -pi/2, 0)) and raster_hf_kernel runs unchanged
__global__ void surf_cam_local_kernel(int n, const double* __restrict__ cam, double* __restrict__ out, double sx, double sy, double sz)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const double* c = cam + (size_t)e * 12;
    double* o = out + (size_t)e * 12;
    const double ex = c[0] - sx, ey = c[1] - sy, ez = c[2] - sz;
    o[0] = sx + ez; o[1] = sy + ey; o[2] = sz - ex;
#pragma unroll
    for (int k = 1; k < 4; k++) { o[3 * k] = c[3 * k + 2]; o[3 * k + 1] = c[3 * k + 1]; o[3 * k + 2] = -c[3 * k]; }
}

static int launch_raster(TgWorld* w, uint8_t* d_obs, const uint8_t* mask, cudaStream_t st, bool terminal_state = false, int e0 = 0, int e1 = -1)
{
    RasterArgs r = w->ra;
    r.obs = d_obs; r.mask = mask;
    if (e1 < 0) e1 = w->n;
    const int cnt = e1 - e0;
    if (cnt <= 0) return TG_OK;
    if (terminal_state) { r.cam = w->eb.term_cam; r.stim = w->eb.term_stim; r.hf_flip = 1; }
    // envs [e0, e1): the kernels index their per-env arrays from 0, so the range is a shift of the base pointers
    r.n = cnt;
    r.obs += (size_t)e0 * w->S * w->S; r.cam += (size_t)e0 * 12; r.stim += (size_t)e0 * 12;
    if (r.mask) r.mask += e0;
    if (r.hf) { r.hf += (size_t)e0 * 2 * SURF_PTS; r.hf_cur += e0; r.hf_meta += (size_t)e0 * 2 * SURF_META; }
    if (r.hf && w->cfg.task.surf_vertical) {
        double* local = w->cam_local + (size_t)e0 * 12;
        surf_cam_local_kernel<<<(cnt + 127) / 128, 128, 0, st>>>(cnt, r.cam, local, w->cfg.task.surf_pos[0], w->cfg.task.surf_pos[1], w->cfg.task.surf_pos[2]);
        w->launches++;
        CK(cudaGetLastError());
        r.cam = local;
    }
    if (w->cfg.task.task == TG_TASK_OBJECT_ROLL) raster_sphere_kernel<<<std::min((cnt + 7) / 8, 8 * w->sm_count), SPH_THREADS, 0, st>>>(r);
    else {
        const int wp = r.hf ? HF_WARPS : RASTER_WARPS;
        const int grid = std::min(w->raster_grid, (((r.hf ? cnt * HF_UPE : cnt) + wp - 1) / wp) * r.bands);
        if (r.hf) raster_hf_kernel<<<grid, HF_THREADS, w->raster_smem, st>>>(r, w->eb.error_flag);
        else if (w->scan_ok) {
            // convex stimulus: scanline raster, then raster_kernel for the envs it flagged (returns at once when there are none)
            const int band_rows = r.S / r.bands, unit_rows = std::min(w->scan_unit_rows, band_rows), units_per_band = cnt * (band_rows / unit_rows);
            int sh_unit = 0;
            while ((1 << sh_unit) < unit_rows) sh_unit++;
            const int sgrid = std::min(w->scan_grid, ((units_per_band + SCAN_WARPS - 1) / SCAN_WARPS) * r.bands);
            // d_fb_count: [0], [1] the fallback counters of even / odd raster passes, [2 + band] the render kernel's unit counters.
            // No memset: a pass's set-up kernel clears the unit counters and the NEXT pass's fallback counter (stream order)
            int* fb = w->d_fb_count + (w->raster_pass & 1u);
            int* fb_next = w->d_fb_count + ((w->raster_pass + 1u) & 1u);
            w->raster_pass++;
            const int per_blk = 128 / w->scan_lpe;
            scan_setup_kernel<<<(cnt + per_blk - 1) / per_blk, 128, 0, st>>>(r, w->d_prim_part, w->d_part_cen, w->d_scan_envs + e0, w->d_fallback + e0, fb, fb_next, w->d_fb_count + 2, w->scan_lpe);
            const int pools = std::max(1, std::min(w->scan_max_pools, (sgrid / r.bands) / 8));
            if (sh_unit == 5) raster_scan_kernel<5><<<sgrid, SCAN_THREADS, w->scan_smem, st>>>(r, w->d_scan_envs + e0, w->d_skin8, w->d_fb_count + 2, pools);
            else raster_scan_kernel<4><<<sgrid, SCAN_THREADS, w->scan_smem, st>>>(r, w->d_scan_envs + e0, w->d_skin8, w->d_fb_count + 2, pools);
            w->launches += 2;
            CK(cudaGetLastError());
            RasterArgs r2 = r;
            r2.mask = w->d_fallback + e0;
            raster_kernel<<<grid, RASTER_THREADS, w->raster_smem, st>>>(r2, fb);
        } else raster_kernel<<<grid, RASTER_THREADS, w->raster_smem, st>>>(r, nullptr);
    }
    w->launches++;
    CK(cudaGetLastError());
    return TG_OK;
}

static int launch_reset(TgWorld* w, const uint8_t* mask, cudaStream_t st)
{
    w->eb.epoch = ++w->epoch;
    TOPO_DISPATCH(w, (reset_kernel<Topo><<<env_grid(w), 128, 0, st>>>(w->cfg.arm, w->cfg.phys, w->cfg.task, w->eb, mask)));
    w->launches++;
    CK(cudaGetLastError());
    return TG_OK;
}

// step kernels are specialised per (topology, task)
#define STEP_LAUNCH(Topo, TASK, grid, block, smem) \
    step_kernel<Topo, TASK><<<grid, block, smem, st>>>(w->cfg.arm, w->cfg.phys, w->cfg.task, w->eb, d_actions, d_reward, d_done, autoreset)
#define STEP_LAUNCH_P(Topo, TASK) \
    step_kernel<Topo, TASK, true><<<pgrid, PUSH_THREADS, w->push_smem, st>>>(w->cfg.arm, w->cfg.phys, w->cfg.task, w->eb, d_actions, d_reward, d_done, autoreset)
#define STEP_DISPATCH(Topo)                                                                               \
    switch (w->cfg.task.task) {                                                                           \
    case TG_TASK_OBJECT_BALANCE: STEP_LAUNCH(Topo, TG_TASK_OBJECT_BALANCE, grid, 128, 0); break;          \
    case TG_TASK_SURFACE_FOLLOW: STEP_LAUNCH(Topo, TG_TASK_SURFACE_FOLLOW, grid, 128, 0); break;          \
    case TG_TASK_OBJECT_PUSH: if (posctl) STEP_LAUNCH_P(Topo, TG_TASK_OBJECT_PUSH); else STEP_LAUNCH(Topo, TG_TASK_OBJECT_PUSH, pgrid, PUSH_THREADS, w->push_smem); break; \
    case TG_TASK_OBJECT_ROLL: if (posctl) STEP_LAUNCH_P(Topo, TG_TASK_OBJECT_ROLL); else STEP_LAUNCH(Topo, TG_TASK_OBJECT_ROLL, pgrid, PUSH_THREADS, w->push_smem); break; \
    default: STEP_LAUNCH(Topo, TG_TASK_EDGE_FOLLOW, grid, 128, 0); break;                                 \
    }

static int launch_step(TgWorld* w, const float* d_actions, float* d_reward, uint8_t* d_done, int autoreset, cudaStream_t st)
{
    w->eb.epoch = ++w->epoch;
    if (w->use_g8) {
        EnvBuffers eb = w->eb;
        eb.step_blocks = w->g8_blocks;
        const dim3 grid(w->g8_blocks + w->standby_blocks);
#define G8_LAUNCH(Topo, TASK) step_kernel_g8<Topo, TASK><<<grid, 128, 0, st>>>(w->cfg.arm, w->cfg.phys, w->cfg.task, eb, d_actions, d_reward, d_done, autoreset)
        const bool surf = w->cfg.task.task == TG_TASK_SURFACE_FOLLOW;
        if (w->cfg.task.task == TG_TASK_OBJECT_BALANCE) G8_LAUNCH(TopoChain6, TG_TASK_OBJECT_BALANCE);
        else if (w->cfg.arm.topo == TG_TOPO_MG400) { if (surf) G8_LAUNCH(TopoMG400, TG_TASK_SURFACE_FOLLOW); else G8_LAUNCH(TopoMG400, TG_TASK_EDGE_FOLLOW); }
        else { if (surf) G8_LAUNCH(TopoChain6, TG_TASK_SURFACE_FOLLOW); else G8_LAUNCH(TopoChain6, TG_TASK_EDGE_FOLLOW); }
#undef G8_LAUNCH
        w->launches++;
        CK(cudaGetLastError());
        return TG_OK;
    }
    const dim3 grid(w->eb.step_blocks + w->standby_blocks); // the extra blocks recompute consumed standbys meanwhile
    const dim3 pgrid((w->n + PUSH_BLOCK - 1) / PUSH_BLOCK);  // object_push: PUSH_BLOCK envs per block, rows in shared memory
    const bool posctl = w->cfg.task.control_mode == 1;
    if (w->cfg.arm.topo == TG_TOPO_MG400) { STEP_DISPATCH(TopoMG400) } else { STEP_DISPATCH(TopoChain6) }
    w->launches++;
    CK(cudaGetLastError());
    return TG_OK;
}

extern "C" int tg_bind_features(TgWorld* w, float* d_feat, float* d_term_feat)
{
    if (!w) return fail(TG_EINVAL, "bad arguments");
    if (d_feat && w->cfg.task.task != TG_TASK_OBJECT_PUSH && w->cfg.task.task != TG_TASK_OBJECT_ROLL && w->cfg.task.task != TG_TASK_SURFACE_FOLLOW)
        return fail(TG_EUNSUPPORTED, "only object_push, object_roll and surface_follow have an extended feature");
    w->eb.feat = d_feat;
    w->eb.term_feat = d_feat ? d_term_feat : nullptr;
    return TG_OK;
}

extern "C" int tg_bind_oracle_obs(TgWorld* w, float* d_oracle, float* d_term_oracle)
{
    if (!w) return fail(TG_EINVAL, "bad arguments");
    w->eb.oracle = d_oracle;
    w->eb.term_oracle = d_oracle ? d_term_oracle : nullptr;
    return TG_OK;
}

extern "C" int tg_reset(TgWorld* w, const uint8_t* d_mask, uint8_t* d_obs, void* stream)
{
    NvtxRange nvtx_("tg_reset");
    if (!w || !d_obs) return fail(TG_EINVAL, "bad arguments");
    CK(cudaSetDevice(w->device));
    int rc;
    if ((rc = launch_reset(w, d_mask, (cudaStream_t)stream))) return rc;
    return launch_raster(w, d_obs, d_mask, (cudaStream_t)stream);
}

extern "C" int tg_reset_only(TgWorld* w, const uint8_t* d_mask, void* stream)
{
    NvtxRange nvtx_("tg_reset_only");
    if (!w) return fail(TG_EINVAL, "bad arguments");
    CK(cudaSetDevice(w->device));
    return launch_reset(w, d_mask, (cudaStream_t)stream);
}

extern "C" int tg_step(TgWorld* w, const float* d_actions, uint8_t* d_obs, float* d_reward, uint8_t* d_done, uint8_t* d_term_obs, void* stream)
{
    NvtxRange nvtx_("tg_step");
    if (!w || !d_actions || !d_obs || !d_reward || !d_done) return fail(TG_EINVAL, "bad arguments");
    CK(cudaSetDevice(w->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (w->eb.pipeline) {
        // finished envs swap their standby start-of-episode state in inside step_kernel: 2 launches per step
        if ((rc = launch_step(w, d_actions, d_reward, d_done, 1, st))) return rc;
        if ((rc = launch_raster(w, d_obs, nullptr, st))) return rc;            // first obs of the new episode for done envs
        if (d_term_obs) return launch_raster(w, d_term_obs, d_done, st, true);  // their terminal obs, on request
        return TG_OK;
    }
    // sequential path (episodes shorter than 2 steps): terminal obs, reset, then every env's observation
    if ((rc = launch_step(w, d_actions, d_reward, d_done, 1, st))) return rc;
    if (d_term_obs && (rc = launch_raster(w, d_term_obs, d_done, st))) return rc;
    if ((rc = launch_reset(w, d_done, st))) return rc;
    return launch_raster(w, d_obs, nullptr, st);
}

// One env step with HOST buffers (the call a numpy VecEnv makes): the observation leaves in chunks, each chunk's
// device->host copy (copy stream) overlapping the next chunk's raster and the terminal-observation raster (caller's stream).
// ascending list of the envs with done != 0 (one block: per-thread segment counts, block scan, ordered write): idx[0] = how many,
// idx[1 + j] = the j-th (j < cap)
__global__ void __launch_bounds__(1024)
done_compact_kernel(const uint8_t* __restrict__ done, int n, int cap, int* __restrict__ idx)
{
    __shared__ int s_warp[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int per = (n + 1023) / 1024, b0 = min(n, t * per), b1 = min(n, b0 + per);
    int cnt = 0;
    for (int e = b0; e < b1; e++) cnt += done[e] != 0;
    int inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += v; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, w, d); if (lane >= d) w += v; }
        s_warp[lane] = w;
    }
    __syncthreads();
    int pos = inc - cnt + (warp > 0 ? s_warp[warp - 1] : 0);
    for (int e = b0; e < b1; e++)
        if (done[e] != 0) { if (pos < cap) idx[1 + pos] = e; pos++; }
    if (t == 1023) idx[0] = s_warp[31];
}

// row j of dst = row idx[1 + j] of src, j < min(idx[0], cap); 16-byte pieces
__global__ void __launch_bounds__(256)
gather_rows_kernel(const uint4* __restrict__ src, const int* __restrict__ idx, int cap, int row_vec, uint4* __restrict__ dst)
{
    const int j = blockIdx.x;
    if (j >= min(idx[0], cap)) return;
    const uint4* s = src + (size_t)idx[1 + j] * row_vec;
    uint4* d = dst + (size_t)j * row_vec;
    for (int k = threadIdx.x; k < row_vec; k += blockDim.x) d[k] = s[k];
}

extern "C" int tg_step_host(TgWorld* w, const TgHostStep* hs, void* stream)
{
    NvtxRange nvtx_("tg_step_host");
    if (!w || !hs || !hs->h_actions || !hs->d_reward || !hs->d_done || !hs->h_reward || !hs->h_done) return fail(TG_EINVAL, "bad arguments");
    if ((hs->h_obs != nullptr) != (hs->d_obs != nullptr)) return fail(TG_EINVAL, "h_obs and d_obs go together");
    if (!hs->h_obs && !hs->h_oracle) return fail(TG_EINVAL, "no observation requested (h_obs and h_oracle are both NULL)");
    if ((hs->h_feat != nullptr) != (hs->d_feat != nullptr)) return fail(TG_EINVAL, "h_feat and d_feat go together");
    if (!w->eb.pipeline) return fail(TG_EUNSUPPORTED, "tg_step_host needs the standby reset pipeline (max_steps >= 2)");
    CK(cudaSetDevice(w->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int n = w->n;
    if (!w->copy_stream) {
        int rc;
        if ((rc = dalloc(w, &w->d_actions_stage, (size_t)n * w->cfg.task.act_dim))) return rc;
        CK(cudaStreamCreateWithFlags(&w->copy_stream, cudaStreamNonBlocking));
        for (cudaEvent_t& ev : w->chunk_ev) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&w->step_ev, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&w->copy_done_ev, cudaEventDisableTiming));
    }
    int chunks = hs->chunks <= 0 ? 4 : hs->chunks;
    if (chunks > TG_HOST_MAX_CHUNKS) chunks = TG_HOST_MAX_CHUNKS;
    if (chunks > n) chunks = n;
    const size_t img = (size_t)w->S * w->S;
    int rc;
    CK(cudaMemcpyAsync(w->d_actions_stage, hs->h_actions, sizeof(float) * n * w->cfg.task.act_dim, cudaMemcpyHostToDevice, st));
    if ((rc = launch_step(w, w->d_actions_stage, hs->d_reward, hs->d_done, 1, st))) return rc;
    CK(cudaEventRecord(w->step_ev, st));
    CK(cudaStreamWaitEvent(w->copy_stream, w->step_ev, 0));
    CK(cudaMemcpyAsync(hs->h_reward, hs->d_reward, sizeof(float) * n, cudaMemcpyDeviceToHost, w->copy_stream));
    CK(cudaMemcpyAsync(hs->h_done, hs->d_done, n, cudaMemcpyDeviceToHost, w->copy_stream));
    if (hs->h_feat) CK(cudaMemcpyAsync(hs->h_feat, hs->d_feat, sizeof(float) * TG_PUSH_NFEAT * n, cudaMemcpyDeviceToHost, w->copy_stream));
    if (hs->h_oracle && w->eb.oracle) CK(cudaMemcpyAsync(hs->h_oracle, w->eb.oracle, sizeof(float) * TG_ORACLE_NOBS * n, cudaMemcpyDeviceToHost, w->copy_stream));
    for (int c = 0; hs->d_obs && c < chunks; c++) {   // observation_mode "oracle" renders nothing
        const int e0 = (int)((long long)n * c / chunks), e1 = (int)((long long)n * (c + 1) / chunks);
        if ((rc = launch_raster(w, hs->d_obs, nullptr, st, false, e0, e1))) return rc;
        CK(cudaEventRecord(w->chunk_ev[c], st));
        CK(cudaStreamWaitEvent(w->copy_stream, w->chunk_ev[c], 0));
        CK(cudaMemcpyAsync(hs->h_obs + img * e0, hs->d_obs + img * e0, img * (size_t)(e1 - e0), cudaMemcpyDeviceToHost, w->copy_stream));
    }
    if (hs->d_obs && hs->d_term_obs && (rc = launch_raster(w, hs->d_term_obs, hs->d_done, st, true))) return rc;
    if (hs->d_obs && hs->d_term_obs && hs->term_cap > 0 && hs->h_term_idx && hs->h_term_obs) {
        // the finished envs' terminal images (and feature rows), compacted and sent with the rest: no second round trip for the caller
        const int cap = std::min(hs->term_cap, n);
        if (w->term_stage_cap < cap) {
            // (rare: the caller grows term_cap geometrically) - not part of the world's state, so outside `allocs`
            CK(cudaStreamSynchronize(st)); CK(cudaStreamSynchronize(w->copy_stream));
            if (w->d_term_idx) cudaFree(w->d_term_idx);
            if (w->d_term_stage) cudaFree(w->d_term_stage);
            if (w->d_term_feat_stage) cudaFree(w->d_term_feat_stage);
            w->d_term_idx = nullptr; w->d_term_stage = nullptr; w->d_term_feat_stage = nullptr; w->term_stage_cap = 0;
            CK(cudaMalloc(&w->d_term_idx, sizeof(int) * ((size_t)cap + 1)));
            CK(cudaMalloc(&w->d_term_stage, img * (size_t)cap));
            CK(cudaMalloc(&w->d_term_feat_stage, sizeof(float) * TG_PUSH_NFEAT * (size_t)cap));
            w->term_stage_cap = cap;
        }
        done_compact_kernel<<<1, 1024, 0, st>>>(hs->d_done, n, cap, w->d_term_idx);
        gather_rows_kernel<<<cap, 256, 0, st>>>(reinterpret_cast<const uint4*>(hs->d_term_obs), w->d_term_idx, cap, (int)(img / 16), reinterpret_cast<uint4*>(w->d_term_stage));
        w->launches += 2;
        const bool feats = hs->h_term_feat && w->eb.term_feat;
        if (feats) {
            gather_rows_kernel<<<cap, 32, 0, st>>>(reinterpret_cast<const uint4*>(w->eb.term_feat), w->d_term_idx, cap, (int)(sizeof(float) * TG_PUSH_NFEAT / 16), reinterpret_cast<uint4*>(w->d_term_feat_stage));
            w->launches++;
        }
        CK(cudaGetLastError());
        // small copies, on the caller's stream: they run beside the observation chunks on the copy stream instead of queueing
        // behind them (the stream ends with a wait on the copy stream anyway)
        CK(cudaMemcpyAsync(hs->h_term_idx, w->d_term_idx, sizeof(int) * ((size_t)cap + 1), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(hs->h_term_obs, w->d_term_stage, img * (size_t)cap, cudaMemcpyDeviceToHost, st));
        if (feats) CK(cudaMemcpyAsync(hs->h_term_feat, w->d_term_feat_stage, sizeof(float) * TG_PUSH_NFEAT * (size_t)cap, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaEventRecord(w->copy_done_ev, w->copy_stream));
    CK(cudaStreamWaitEvent(st, w->copy_done_ev, 0));   // the caller's stream is complete only when the host buffers are
    return TG_OK;
}

extern "C" int tg_physics_only(TgWorld* w, const float* d_actions, float* d_reward, uint8_t* d_done, void* stream)
{
    NvtxRange nvtx_("tg_physics_only");
    if (!w || !d_actions) return fail(TG_EINVAL, "bad arguments");
    CK(cudaSetDevice(w->device));
    return launch_step(w, d_actions, d_reward ? d_reward : w->d_reward_internal, d_done ? d_done : w->d_done_internal, 0, (cudaStream_t)stream);
}

extern "C" int tg_raster_only(TgWorld* w, uint8_t* d_obs, void* stream)
{
    NvtxRange nvtx_("tg_raster_only");
    if (!w || !d_obs) return fail(TG_EINVAL, "bad arguments");
    CK(cudaSetDevice(w->device));
    return launch_raster(w, d_obs, nullptr, (cudaStream_t)stream);
}

extern "C" int tg_state_size(const TgWorld* w) { return w ? 2 * w->nb + 7 + 4 + 14 + 1 : 0; }

extern "C" int tg_get_state(TgWorld* w, double* h, void* stream)
{
    if (!w || !h) return fail(TG_EINVAL, "bad arguments");
    CK(cudaSetDevice(w->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int n = w->n, nb = w->nb, sz = tg_state_size(w);
    std::vector<double> q((size_t)nb * n), qd((size_t)nb * n), tcp((size_t)7 * n), emb(n), ang(n);
    std::vector<int> steps(n), rs(n);
    CK(cudaStreamSynchronize(st));
    CK(cudaMemcpy(q.data(), w->eb.q, sizeof(double) * nb * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(qd.data(), w->eb.qd, sizeof(double) * nb * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(tcp.data(), w->eb.tcp, sizeof(double) * 7 * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(emb.data(), w->eb.embed, sizeof(double) * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ang.data(), w->eb.edge_ang, sizeof(double) * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(steps.data(), w->eb.steps, sizeof(int) * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rs.data(), w->eb.reset_substeps, sizeof(int) * n, cudaMemcpyDeviceToHost));
    for (int e = 0; e < n; e++) {
        double* o = h + (size_t)e * sz;
        for (int i = 0; i < nb; i++) { o[i] = q[(size_t)i * n + e]; o[nb + i] = qd[(size_t)i * n + e]; }
        for (int c = 0; c < 7; c++) o[2 * nb + c] = tcp[(size_t)e * 7 + c];
        o[2 * nb + 7] = emb[e]; o[2 * nb + 8] = ang[e]; o[2 * nb + 9] = steps[e]; o[2 * nb + 10] = rs[e];
        for (int c = 0; c < 15; c++) o[2 * nb + 11 + c] = 0.0;
    }
    if (w->eb.goal) {
        std::vector<int> gl(n);
        CK(cudaMemcpy(gl.data(), w->eb.goal, sizeof(int) * n, cudaMemcpyDeviceToHost));
        for (int e = 0; e < n; e++) h[(size_t)e * sz + 2 * nb + 25] = gl[e];
    }
    if (w->eb.obj) {
        std::vector<double> ob((size_t)13 * n), gr(n);
        CK(cudaMemcpy(ob.data(), w->eb.obj, sizeof(double) * 13 * n, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(gr.data(), w->eb.grav, sizeof(double) * n, cudaMemcpyDeviceToHost));
        for (int e = 0; e < n; e++) {
            double* o = h + (size_t)e * sz + 2 * nb + 11;
            for (int c = 0; c < 13; c++) o[c] = ob[(size_t)e * 13 + c];
            o[13] = gr[e];
        }
    }
    return TG_OK;
}

extern "C" int tg_set_state(TgWorld* w, const double* h, void* stream)
{
    if (!w || !h) return fail(TG_EINVAL, "bad arguments");
    CK(cudaSetDevice(w->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int n = w->n, nb = w->nb, sz = tg_state_size(w);
    std::vector<double> q((size_t)nb * n), qd((size_t)nb * n), emb(n), ang(n);
    std::vector<int> steps(n);
    for (int e = 0; e < n; e++) {
        const double* o = h + (size_t)e * sz;
        for (int i = 0; i < nb; i++) { q[(size_t)i * n + e] = o[i]; qd[(size_t)i * n + e] = o[nb + i]; }
        emb[e] = o[2 * nb + 7]; ang[e] = o[2 * nb + 8]; steps[e] = (int)o[2 * nb + 9];
    }
    CK(cudaStreamSynchronize(st));
    CK(cudaMemcpy(w->eb.q, q.data(), sizeof(double) * nb * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(w->eb.qd, qd.data(), sizeof(double) * nb * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(w->eb.embed, emb.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(w->eb.edge_ang, ang.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(w->eb.steps, steps.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
    if (w->eb.obj) {
        std::vector<double> ob((size_t)13 * n), gr(n);
        for (int e = 0; e < n; e++) {
            const double* o = h + (size_t)e * sz + 2 * nb + 11;
            for (int c = 0; c < 13; c++) ob[(size_t)e * 13 + c] = o[c];
            gr[e] = o[13];
        }
        CK(cudaMemcpy(w->eb.obj, ob.data(), sizeof(double) * 13 * n, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(w->eb.grav, gr.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
    }
    if (w->eb.goal) {
        std::vector<int> gl(n);
        for (int e = 0; e < n; e++) gl[e] = (int)h[(size_t)e * sz + 2 * nb + 25];
        CK(cudaMemcpy(w->eb.goal, gl.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
    }
    return TG_OK;
}

extern "C" int tg_get_camera(TgWorld* w, double* h_cam, void* stream)
{
    if (!w || !h_cam) return fail(TG_EINVAL, "bad arguments");
    CK(cudaSetDevice(w->device));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    CK(cudaMemcpy(h_cam, w->eb.cam, sizeof(double) * 12 * w->n, cudaMemcpyDeviceToHost));
    return TG_OK;
}

// ------------------------------------------------------------ test hooks
template <class F>
static int with_tmp(TgWorld* w, int n, int nb, const double* h_a, const double* h_b, size_t out_count, double* h_out, bool out_is_a, F launch)
{
    double *da = nullptr, *db = nullptr, *dout = nullptr;
    CK(cudaSetDevice(w->device));
    CK(cudaMalloc(&da, sizeof(double) * n * nb));
    CK(cudaMemcpy(da, h_a, sizeof(double) * n * nb, cudaMemcpyHostToDevice));
    if (h_b) { CK(cudaMalloc(&db, sizeof(double) * n * nb)); CK(cudaMemcpy(db, h_b, sizeof(double) * n * nb, cudaMemcpyHostToDevice)); }
    if (!out_is_a) CK(cudaMalloc(&dout, sizeof(double) * out_count));
    launch(da, db, dout);
    w->launches++;
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess && !out_is_a) e = cudaMemcpy(h_out, dout, sizeof(double) * out_count, cudaMemcpyDeviceToHost);
    cudaFree(da); if (db) cudaFree(db); if (dout) cudaFree(dout);
    if (e != cudaSuccess) return fail(TG_ECUDA, "test hook failed: %s", cudaGetErrorString(e));
    return TG_OK;
}

extern "C" int tg_test_inverse_dynamics(TgWorld* w, int n, const double* h_q, const double* h_qd, double* h_tau)
{
    if (!w || n <= 0) return fail(TG_EINVAL, "bad arguments");
    return with_tmp(w, n, w->nb, h_q, h_qd, (size_t)n * w->nb, h_tau, false, [&](double* q, double* qd, double* out) {
        TOPO_DISPATCH(w, (test_id_kernel<Topo><<<(n + 63) / 64, 64>>>(w->cfg.arm, w->cfg.phys, n, q, qd, out)));
    });
}

extern "C" int tg_test_mass_matrix(TgWorld* w, int n, const double* h_q, double* h_M)
{
    if (!w || n <= 0) return fail(TG_EINVAL, "bad arguments");
    return with_tmp(w, n, w->nb, h_q, nullptr, (size_t)n * w->nb * w->nb, h_M, false, [&](double* q, double*, double* out) {
        TOPO_DISPATCH(w, (test_mass_kernel<Topo><<<(n + 63) / 64, 64>>>(w->cfg.arm, n, q, out)));
    });
}

extern "C" int tg_test_substep(TgWorld* w, int n, int nsteps, double* h_q, double* h_qd, const double* h_target_vel)
{
    if (!w || n <= 0) return fail(TG_EINVAL, "bad arguments");
    CK(cudaSetDevice(w->device));
    const size_t bytes = sizeof(double) * n * w->nb;
    double *q, *qd, *tv;
    CK(cudaMalloc(&q, bytes)); CK(cudaMalloc(&qd, bytes)); CK(cudaMalloc(&tv, bytes));
    CK(cudaMemcpy(q, h_q, bytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(qd, h_qd, bytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(tv, h_target_vel, bytes, cudaMemcpyHostToDevice));
    TOPO_DISPATCH(w, (test_substep_kernel<Topo><<<(n + 31) / 32, 32>>>(w->cfg.arm, w->cfg.phys, n, nsteps, q, qd, tv)));
    w->launches++;
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(h_q, q, bytes, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(h_qd, qd, bytes, cudaMemcpyDeviceToHost);
    cudaFree(q); cudaFree(qd); cudaFree(tv);
    if (e != cudaSuccess) return fail(TG_ECUDA, "test_substep failed: %s", cudaGetErrorString(e));
    return TG_OK;
}

extern "C" int tg_test_substep_g8(TgWorld* w, int n, int nsteps, double* h_q, double* h_qd, const double* h_target_vel)
{
    if (!w || n <= 0) return fail(TG_EINVAL, "bad arguments");
    CK(cudaSetDevice(w->device));
    const size_t bytes = sizeof(double) * n * w->nb;
    double *q, *qd, *tv;
    CK(cudaMalloc(&q, bytes)); CK(cudaMalloc(&qd, bytes)); CK(cudaMalloc(&tv, bytes));
    CK(cudaMemcpy(q, h_q, bytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(qd, h_qd, bytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(tv, h_target_vel, bytes, cudaMemcpyHostToDevice));
    TOPO_DISPATCH(w, (test_substep_g8_kernel<Topo><<<(n + 15) / 16, 128>>>(w->cfg.arm, w->cfg.phys, n, nsteps, q, qd, tv)));
    w->launches++;
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(h_q, q, bytes, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(h_qd, qd, bytes, cudaMemcpyDeviceToHost);
    cudaFree(q); cudaFree(qd); cudaFree(tv);
    if (e != cudaSuccess) return fail(TG_ECUDA, "test_substep_g8 failed: %s", cudaGetErrorString(e));
    return TG_OK;
}

extern "C" long long tg_launch_count(const TgWorld* w) { return w ? w->launches : 0; }

#ifdef TG_RASTER_STATS
// diagnostic build only (tools/raster_stats.py): read and clear the raster counters
extern "C" int tg_debug_raster_stats(unsigned long long* out48)
{
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyFromSymbol(out48, g_rstats, sizeof(unsigned long long) * 48));
    unsigned long long z[48] = {0};
    CK(cudaMemcpyToSymbol(g_rstats, z, sizeof(z)));
    return TG_OK;
}
#endif
