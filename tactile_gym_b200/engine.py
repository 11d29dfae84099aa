"""Batched world: N env instances of one task on one GPU, driven through the C ABI.

`TactileWorld` owns the TgWorld handle and the torch.cuda tensors the kernels write into
(obs uint8 [N,S,S,1], reward f32 [N], done u8 [N]).  Everything is enqueued on torch's current stream;
`step()` does not synchronise.  Random draws for resets are produced on the host by one numpy
RandomState per env - seeded exactly like the reference's gym seeding - and uploaded in rounds.
"""
import ctypes as C
import os

import numpy as np

from . import _lib as L
from . import scene, seeding

DRAW_ROUNDS = 64        # slots of the per-env draw ring on the device
DRAW_CHECK_EVERY = 32   # steps / resets between (asynchronous) read-backs of how many draws were consumed


def _sparse_flag(env_modes):
    """reward_mode -> TgTask.sparse_reward; anything but 'dense' / 'sparse' is the reference's sys.exit (ValueError here)"""
    mode = env_modes.get("reward_mode", "dense")
    if mode not in ("dense", "sparse"):
        raise ValueError("Incorrect reward_mode specified: {}".format(mode))
    return 1 if mode == "sparse" else 0


def _program(task, entries):
    """The reset draws as a program over numpy's MT19937 stream for the device RNG (TgTask.draw_kind, csrc/tg_rng.cuh): one
    (kind, lo, hi) per draw, in the reference's call order.  Constants take TgTask.draw_default[d]."""
    assert len(entries) == task.n_draws, (len(entries), task.n_draws)
    for d, (kind, lo, hi) in enumerate(entries):
        task.draw_kind[d], task.draw_lo[d], task.draw_hi[d] = kind, lo, hi


_CONST = (L.TG_DRAW_CONST, 0.0, 0.0)


def _uni(lo, hi):
    return (L.TG_DRAW_UNIFORM, lo, hi)


def _control_mode(env_modes, arm_type):
    """control_mode -> (TgTask.control_mode, pos_max_steps).  robot.py:156-186; _max_blocking_pos_move_steps = 10 in every env."""
    mode = env_modes["control_mode"]
    if mode == "TCP_velocity_control":
        return 0, 10
    if mode == "TCP_position_control":
        return 1, 10    # ur5 and mg400 (the latter slaves joints 5..7 in the IK result too, mg400.py:167-172)
    raise ValueError("Incorrect control mode specified: {}".format(mode))


def edge_follow_config(env_modes, image_size, max_steps, n_envs, lanes_per_warp=0):
    """EdgeFollowEnv.__init__ (rl_envs/exploration/edge_follow/edge_follow_env.py:23-134) as a TgConfig.
    Returns (cfg, keepalive) - keepalive holds the numpy arrays the config points into."""
    arm_type = env_modes["arm_type"]
    if "tactile_sensor_name" not in env_modes:
        raise KeyError("env_modes['tactile_sensor_name'] is required (edge_follow_env.py:59)")
    sensor = env_modes["tactile_sensor_name"]
    movement_mode = env_modes["movement_mode"]
    control_mode = env_modes["control_mode"]
    noise_mode = env_modes["noise_mode"]
    if arm_type not in ("ur5", "mg400"):
        raise ValueError("Incorrect arm type specified {}".format(arm_type))
    typ = "standard"
    S = int(image_size[0])
    mj = scene.load_model_json(arm_type, sensor, typ)
    sj = scene.load_sensor_json(sensor, typ)
    cam = sj["types"][typ]
    arm, control_links = scene.reduce_model(mj, sensor, cam["cam_pos"], cam["cam_rpy"])

    cfg = L.TgConfig()
    cfg.n_envs = n_envs
    cfg.lanes_per_warp = lanes_per_warp
    cfg.arm = arm
    cfg.phys = scene.default_physics(substeps=int(np.floor((1.0 / 10.0) / (1.0 / 240.0))))

    t = cfg.task
    t.task = L.TG_TASK_EDGE_FOLLOW
    t.sparse_reward = _sparse_flag(env_modes)                                                 # edge_follow_env.py:430-438
    t.max_steps = int(max_steps)
    idx = {"xy": [0, 1], "xyz": [0, 1, 2], "xyRz": [0, 1, 5], "xyzRz": [0, 1, 2, 5]}[movement_mode]
    t.act_dim = len(idx)
    for k in range(6):
        t.act_index[k] = idx[k] if k < len(idx) else -1
    t.act_min, t.act_max = -0.25, 0.25
    t.control_mode, t.pos_max_steps = _control_mode(env_modes, arm_type)
    max_pos_vel, max_ang_vel = 0.01, 5.0 * (np.pi / 180)          # edge_follow_env.py:155-165
    if t.control_mode == 1:                                        # :143-153: m / rad per step
        max_pos_vel, max_ang_vel = 0.001, 1 * (np.pi / 180)
    hi = [max_pos_vel] * 3 + [0.0, 0.0, max_ang_vel]
    for k in range(6):
        t.act_lo[k], t.act_hi[k] = -hi[k], hi[k]
    lims = np.zeros((6, 2))
    if arm_type == "mg400":                                        # :72-80
        edge_pos, edge_len, stim = [0.33, 0.0, 0.0], 0.105, "short_edge"
        lims[0], lims[1], lims[2], lims[5] = (-0.15, 0.15), (-0.11, 0.11), (-0.1, 0.1), (-np.pi, np.pi)
    else:                                                          # :81-88
        edge_pos, edge_len, stim = [0.65, 0.0, 0.0], 0.175, "long_edge"
        lims[0], lims[1], lims[2], lims[5] = (-0.175, 0.175), (-0.175, 0.175), (-0.1, 0.1), (-np.pi, np.pi)
    edge_height = 0.035
    wf_pos = [edge_pos[0], edge_pos[1], edge_height]               # :106-107
    wf_rpy = [-np.pi, 0.0, np.pi / 2]
    for k in range(3):
        t.workframe_pos[k], t.workframe_rpy[k] = wf_pos[k], wf_rpy[k]
        t.edge_pos[k] = edge_pos[k]
        t.init_rpy[k] = 0.0
    for k in range(6):
        t.tcp_lims[k][0], t.tcp_lims[k][1] = lims[k]
    t.edge_len, t.edge_height, t.termination_dist = edge_len, edge_height, 0.01
    embed_default = 0.0035                                         # :91-96
    if noise_mode == "rand_height":                                # :291-297
        t.embed_lo, t.embed_hi = {"tactip": (0.0015, 0.0065), "digit": (0.0011, 0.0028), "digitac": (0.0015, 0.0045)}[sensor]
    else:
        t.embed_lo = t.embed_hi = embed_default
    t.n_draws = 2
    t.draw_default[0], t.draw_default[1] = embed_default, 0.0
    _program(t, [_uni(t.embed_lo, t.embed_hi) if t.embed_lo != t.embed_hi else _CONST, _uni(-np.pi, np.pi)])   # :293-297, :240

    dep, gray, mask = scene.load_refimg(sensor, typ, S)
    tris = scene.load_stimulus(stim)
    rest = scene.load_rest_pose("edge_follow", arm_type, sensor, typ, control_links)
    s = cfg.sensor
    s.image_size, s.border_on = S, 1                               # turn_off_border=False (:119)
    s.fov_deg, s.near_, s.far_ = sj["fov"], sj["near"], sj["far"]
    s.h_nodef_dep = dep.ctypes.data_as(C.POINTER(C.c_float))
    s.h_nodef_gray = gray.ctypes.data_as(C.POINTER(C.c_float))
    s.h_border_mask = mask.ctypes.data_as(C.POINTER(C.c_uint8))
    prims, prim_nv = scene.merge_coplanar(tris)
    s.n_prim = len(prims)
    s.h_prims = prims.ctypes.data_as(C.POINTER(C.c_double))
    s.h_prim_nv = prim_nv.ctypes.data_as(C.POINTER(C.c_int32))
    parts, part_cen = scene.convex_parts(prims, prim_nv)           # boxes: the scanline raster applies
    if parts is not None:
        part_cen = np.ascontiguousarray(part_cen, dtype=np.float64)
        s.h_prim_part = parts.ctypes.data_as(C.POINTER(C.c_int32))
        s.h_part_centroid = part_cen.ctypes.data_as(C.POINTER(C.c_double))
        s.n_parts = len(part_cen)
    cfg.h_rest_q = rest.ctypes.data_as(C.POINTER(C.c_double))
    return cfg, (dep, gray, mask, tris, rest, prims, prim_nv, parts, part_cen)


def _uniform53(a, b):
    """numpy RandomState.random_sample from two raw 32-bit MT19937 outputs (genrand_res53)"""
    return ((a >> np.uint32(5)).astype(np.float64) * 67108864.0 + (b >> np.uint32(6)).astype(np.float64)) / 9007199254740992.0


def edge_follow_draws(task):
    """reset_task + update_edge draws (edge_follow_env.py:293-297, 240): [embed_dist,] edge_ang per reset"""
    lo, hi = task.embed_lo, task.embed_hi

    def draw(rng, rounds):
        out = np.empty((rounds, 2))
        if lo != hi:
            u = rng.random_sample(2 * rounds)
            out[:, 0] = lo + (hi - lo) * u[0::2]
            out[:, 1] = -np.pi + (np.pi - (-np.pi)) * u[1::2]
        else:
            u = rng.random_sample(rounds)
            out[:, 0] = lo
            out[:, 1] = -np.pi + (np.pi - (-np.pi)) * u
        return out

    return draw


def object_balance_draws(rand_gravity, rand_embed_dist, embed_lo, embed_hi, embed_default):
    """ObjectBalanceEnv draws per reset, in the reference's call order (object_balance_env.py:300-313, 366-371):
    [uniform(-1,-0.1)] [uniform(embed range)] choice([-1,1]) rand() choice([-1,1]) rand().
    The calls are emulated on the raw 32-bit MT19937 stream so that a whole batch of resets is drawn at once:
    uniform / rand take two outputs (53-bit double), choice takes one (randint(0, 2) = output & 1)."""
    per = (2 if rand_gravity else 0) + (2 if rand_embed_dist else 0) + 6

    def draw(rng, rounds):
        raw = rng.randint(0, 2 ** 32, size=per * rounds, dtype=np.uint32).reshape(rounds, per)
        k = 0
        out = np.empty((rounds, 4))
        if rand_gravity:
            out[:, 0] = -1.0 + (-0.1 - (-1.0)) * _uniform53(raw[:, k], raw[:, k + 1]); k += 2
        else:
            out[:, 0] = -0.1
        if rand_embed_dist:
            out[:, 1] = embed_lo + (embed_hi - embed_lo) * _uniform53(raw[:, k], raw[:, k + 1]); k += 2
        else:
            out[:, 1] = embed_default
        for col in (2, 3):
            sign = np.where((raw[:, k] & np.uint32(1)) == 0, -1.0, 1.0); k += 1
            out[:, col] = sign * _uniform53(raw[:, k], raw[:, k + 1]); k += 2
        return out

    return draw


def object_balance_config(env_modes, image_size, max_steps, n_envs, lanes_per_warp=0):
    """ObjectBalanceEnv.__init__ (rl_envs/nonprehensile_manipulation/object_balance/object_balance_env.py:22-124,
    object_mode 'pole') as a TgConfig.  Returns (cfg, keepalive, draw_fn)."""
    import json
    import os

    arm_type, sensor = env_modes["arm_type"], env_modes["tactile_sensor_name"]
    if env_modes.get("object_mode", "pole") != "pole":
        raise NotImplementedError("object_mode %r: only 'pole' is built" % env_modes.get("object_mode"))
    if arm_type != "ur5":
        raise ValueError("object_balance has rest poses for the ur5 only among the built arms (rest_poses.py)")
    typ, S = "standard", int(image_size[0])
    mj = scene.load_model_json(arm_type, sensor, typ)
    sj = scene.load_sensor_json(sensor, typ)
    cam = sj["types"][typ]
    arm, control_links = scene.reduce_model(mj, sensor, cam["cam_pos"], cam["cam_rpy"])
    cfg = L.TgConfig()
    cfg.n_envs, cfg.lanes_per_warp, cfg.arm = n_envs, lanes_per_warp, arm
    cfg.phys = scene.default_physics(substeps=int(np.floor((1.0 / 20.0) / (1.0 / 240.0))))   # :33-35 -> 12
    t = cfg.task
    t.task, t.max_steps = L.TG_TASK_OBJECT_BALANCE, int(max_steps)
    t.sparse_reward = _sparse_flag(env_modes)                                                 # object_balance_env.py:508-518
    idx = {"xy": [0, 1], "xyz": [0, 1, 2], "RxRy": [3, 4], "xyRxRy": [0, 1, 3, 4]}[env_modes["movement_mode"]]   # :416-442
    t.act_dim = len(idx)
    for k in range(6):
        t.act_index[k] = idx[k] if k < len(idx) else -1
    t.act_min, t.act_max = -0.25, 0.25
    t.control_mode, t.pos_max_steps = _control_mode(env_modes, arm_type)                          # :36: _max_blocking_pos_move_steps = 10
    mv, ma = 0.01, 5.0 * (np.pi / 180)                                                          # :141-151
    if t.control_mode == 1:                                                                      # :129-139: m / rad per step
        mv, ma = 0.001, 1 * (np.pi / 180)
    hi = [mv, mv, mv, ma, ma, 0.0]
    for k in range(6):
        t.act_lo[k], t.act_hi[k] = -hi[k], hi[k]
    wf_pos, a45 = [0.55, 0.0, 0.35], 45 * np.pi / 180                                            # :73-92
    lims = [(-0.1, 0.1)] * 3 + [(-a45, a45)] * 3
    for k in range(3):
        t.workframe_pos[k], t.workframe_rpy[k], t.init_rpy[k] = wf_pos[k], 0.0, 0.0
    for k in range(6):
        t.tcp_lims[k][0], t.tcp_lims[k][1] = lims[k]
    with open(os.path.join(scene.ASSETS, "objects", "pole.json")) as f:
        pole = json.load(f)
    t.obj_mass = pole["mass"]
    for k in range(3):
        t.obj_inertia[k], t.obj_com_off[k], t.obj_base_com[k] = pole["inertia_diag"][k], pole["com_off"][k], pole["base_com"][k]
        t.obj_init_rpy[k] = [0.0, 0.0, -np.pi / 2][k]                                            # :207
    t.obj_base_w, t.obj_base_h, t.obj_force = 0.1, 0.0025, 0.1                                  # :170-172, :339
    t.obj_term_deg, t.obj_term_pos = 35.0, 0.1                                                   # :58-59
    t.p2p_erp, t.p2p_max_impulse = 0.2, 500.0                                                    # [EXT] bullet / pybullet defaults
    embed_default = {"tactip": 0.0035, "digitac": 0.0015, "digit": 0.0015}[sensor]               # :63-68
    lo, hi_e = {"tactip": (0.003, 0.006), "digitac": (0.001, 0.0025), "digit": (0.0015, 0.0025)}[sensor]   # :307-312
    t.n_draws = 4
    t.draw_default[0], t.draw_default[1], t.draw_default[2], t.draw_default[3] = -0.1, embed_default, 0.0, 0.0
    _program(t, [_uni(-1.0, -0.1) if env_modes.get("rand_gravity", False) else _CONST,                  # :300-305
                 _uni(lo, hi_e) if env_modes.get("rand_embed_dist", False) else _CONST,                  # :307-313
                 (L.TG_DRAW_CHOICE_RAND, 0.0, 0.0), (L.TG_DRAW_CHOICE_RAND, 0.0, 0.0)])                  # :366-371
    dep, gray, mask = scene.load_refimg(sensor, typ, S)
    tris = scene.load_stimulus("pole")
    rest = scene.load_rest_pose("object_balance", arm_type, sensor, typ, control_links)
    s = cfg.sensor
    s.image_size, s.border_on = S, 1
    s.fov_deg, s.near_, s.far_ = sj["fov"], sj["near"], sj["far"]
    s.h_nodef_dep = dep.ctypes.data_as(C.POINTER(C.c_float))
    s.h_nodef_gray = gray.ctypes.data_as(C.POINTER(C.c_float))
    s.h_border_mask = mask.ctypes.data_as(C.POINTER(C.c_uint8))
    prims, prim_nv = scene.merge_coplanar(tris)
    s.n_prim = len(prims)
    s.h_prims = prims.ctypes.data_as(C.POINTER(C.c_double))
    s.h_prim_nv = prim_nv.ctypes.data_as(C.POINTER(C.c_int32))
    parts, part_cen = scene.convex_parts(prims, prim_nv)           # boxes: the scanline raster applies
    if parts is not None:
        part_cen = np.ascontiguousarray(part_cen, dtype=np.float64)
        s.h_prim_part = parts.ctypes.data_as(C.POINTER(C.c_int32))
        s.h_part_centroid = part_cen.ctypes.data_as(C.POINTER(C.c_double))
        s.n_parts = len(part_cen)
    cfg.h_rest_q = rest.ctypes.data_as(C.POINTER(C.c_double))
    draw = object_balance_draws(bool(env_modes.get("rand_gravity", False)), bool(env_modes.get("rand_embed_dist", False)), lo, hi_e, embed_default)
    return cfg, (dep, gray, mask, tris, rest, prims, prim_nv, parts, part_cen), draw


def surface_follow_draws(noise_mode="simplex", one_d=False):
    """BaseSurfaceEnv.reset_task draws (base_surface_env.py:539-547): update_surface's `np_random.randint(1e8)` (the
    OpenSimplex seed, :448; not drawn for noise_mode "none") then make_goal's `uniform(-pi, pi)` (:508) or, for the yz / yzRx
    movement modes, `choice([-1, 1])` (:513).  randint's rejection sampling consumes a variable number of raw outputs, so the
    legacy RandomState calls themselves are used."""

    def draw(rng, rounds):
        out = np.empty((rounds, 2))
        for r in range(rounds):
            out[r, 0] = rng.randint(1e8) if noise_mode == "simplex" else 0
            out[r, 1] = rng.choice([-1, 1]) if one_d else rng.uniform(-np.pi, np.pi)
        return out

    return draw


def surface_follow_goal_config(env_modes, image_size, max_steps, n_envs, lanes_per_warp=0):
    """surface_follow-v1: SurfaceFollowGoalEnv (surface_follow_goal/surface_follow_goal_env.py) - the policy steers x / y itself
    towards the goal instead of being driven there."""
    return surface_follow_config(env_modes, image_size, max_steps, n_envs, lanes_per_warp=lanes_per_warp, variant="goal")


def surface_follow_vert_config(env_modes, image_size, max_steps, n_envs, lanes_per_warp=0):
    """surface_follow-v2: SurfaceFollowVertEnv (surface_follow_vert/surface_follow_vert_env.py): on the horizontal surfaces
    (noise_mode 'simplex' - which leaves the surface flat for its 'xRz' movement mode, base_surface_env.py:443-452 - or 'none')
    and on its own upright one (noise_mode 'vertical_simplex': the `forward` sensor type facing a vertical heightfield)."""
    return surface_follow_config(env_modes, image_size, max_steps, n_envs, lanes_per_warp=lanes_per_warp, variant="vert")


def surface_follow_config(env_modes, image_size, max_steps, n_envs, lanes_per_warp=0, variant="auto"):
    """SurfaceFollowAutoEnv (variant "auto", surface_follow-v0) / SurfaceFollowGoalEnv (variant "goal", surface_follow-v1) on
    BaseSurfaceEnv.__init__ (rl_envs/exploration/surface_follow/base_surface_env.py:14-150) as a TgConfig.
    Returns (cfg, keepalive, draw_fn)."""
    arm_type, sensor = env_modes["arm_type"], env_modes["tactile_sensor_name"]
    noise_mode, movement_mode = env_modes.get("noise_mode", "simplex"), env_modes["movement_mode"]
    vertical = noise_mode == "vertical_simplex"
    if vertical:
        # upright heightfield + `forward` sensor type (base_surface_env.py:60-63, 83-107): TgTask.surf_vertical
        if variant != "vert" or movement_mode != "xRz":
            raise ValueError("Incorrect movement mode specified")                                # update_surface :459-463 knows "xRz" only
    elif noise_mode not in ("simplex", "none", "random"):
        raise ValueError("Incorrect noise mode specified")                                       # :461, :463
    if variant == "vert":
        if movement_mode != "xRz":
            raise NotImplementedError("surface_follow-v2 encodes actions for movement_mode 'xRz' only (surface_follow_vert_env.py:43-45)")
    elif movement_mode not in ("yz", "xyz", "yzRx", "xyzRxRy"):
        raise ValueError("Incorrect movement mode specified: %r" % movement_mode)
    one_d = movement_mode in ("yz", "yzRx", "xRz")
    typ, S = ("forward" if vertical else "standard"), int(image_size[0])                         # :60-63
    mj = scene.load_model_json(arm_type, sensor, typ)
    sj = scene.load_sensor_json(sensor, typ)
    cam = sj["types"][typ]
    arm, control_links = scene.reduce_model(mj, sensor, cam["cam_pos"], cam["cam_rpy"])
    cfg = L.TgConfig()
    cfg.n_envs, cfg.lanes_per_warp, cfg.arm = n_envs, lanes_per_warp, arm
    cfg.phys = scene.default_physics(substeps=int(np.floor((1.0 / 10.0) / (1.0 / 240.0))))      # :26-28 -> 24
    t = cfg.task
    t.task, t.max_steps = L.TG_TASK_SURFACE_FOLLOW, int(max_steps)
    t.control_mode, t.pos_max_steps = _control_mode(env_modes, arm_type)
    if variant == "vert":
        idx = [0, 5]                                                                             # surface_follow_vert_env.py:43-45
    elif variant == "goal":
        idx = {"yz": [1, 2], "xyz": [0, 1, 2], "yzRx": [1, 2, 3], "xyzRxRy": [0, 1, 2, 3, 4]}[movement_mode]   # surface_follow_goal_env.py:27-52
    else:
        idx = {"yz": [2], "xyz": [2], "yzRx": [2, 3], "xyzRxRy": [2, 3, 4]}[movement_mode]        # surface_follow_auto_env.py:45-55
    t.sparse_reward = _sparse_flag(env_modes)                                                    # surface_follow_auto_env.py:59-73
    # heights (update_surface :434-463): 2-d simplex for xyz / xyzRxRy, 1-d (along y) for yz / yzRx, flat for "none";
    # goal direction (make_goal :501-520): an angle, or choice([-1, 1]) along y
    # ("xRz" is in neither of update_surface's lists, so -v2's simplex surface keeps its zeros: flat; the seed is still drawn)
    t.surf_mode = 2 if noise_mode == "none" or movement_mode == "xRz" else (1 if one_d else 0)
    if vertical:
        t.surf_mode, t.surf_vertical = 3, 1                                                      # gen_heigtfield_simplex_1d_vertical :359-379
    if noise_mode == "random":
        # gen_heigtfield_noisey :302-318: 1,024 uniform draws per reset, whatever the movement mode - far too many to stream
        # from the host, so this mode runs on the device RNG (the world switches to it by itself)
        t.surf_mode = 4
    t.surf_dir_mode = 1 if one_d else 0
    t.act_dim = len(idx)
    for k in range(6):
        t.act_index[k] = idx[k] if k < len(idx) else -1
    t.act_min, t.act_max = -0.25, 0.25
    mv, ma = 0.01, 5.0 * (np.pi / 180)                                                          # :197-206
    if t.control_mode == 1:                                                                      # :172-181: m / rad per step
        mv, ma = 0.001, 1 * (np.pi / 180)
    hi = [mv, mv, 0.0, 0.0, 0.0, ma] if vertical else [mv, mv, mv, ma, ma, 0.0]                   # :183-194
    for k in range(6):
        t.act_lo[k], t.act_hi[k] = -hi[k], hi[k]
    wd = [0.33, 0.0, 0.0] if arm_type in ("mg400", "magician") else [0.65, 0.0, 0.0]              # :54-57
    hrange, extent = 0.025, 0.15                                                                 # :240-246
    wf_pos, wf_rpy = [wd[0], wd[1], hrange], [-np.pi, 0.0, np.pi / 2]                             # :111-114
    lims = [(-extent, extent), (-extent, extent), (-hrange, hrange), (-np.pi / 4, np.pi / 4), (-np.pi / 4, np.pi / 4), (0.0, 0.0)]
    surf_pos = [wd[0], wd[1], hrange]                                                            # :261
    if vertical:                                                                                 # :83-107, :248-259
        surf_pos = [wd[0], wd[1], 0.15 + hrange]
        wf_pos, wf_rpy = list(surf_pos), [-np.pi, 0.0, 0.0]
        lims = [(-hrange, hrange), (-extent, extent), (0.0, 0.0), (0.0, 0.0), (0.0, 0.0), (-np.pi / 4, np.pi / 4)]
    for k in range(3):
        t.workframe_pos[k], t.workframe_rpy[k], t.init_rpy[k] = wf_pos[k], wf_rpy[k], 0.0
        t.surf_pos[k] = surf_pos[k]
    for k in range(6):
        t.tcp_lims[k][0], t.tcp_lims[k][1] = lims[k]
    t.termination_dist = 0.01                                                                    # :78
    t.surf_grid, t.surf_range, t.surf_interp, t.surf_extent = 0.006, hrange, 0.05, extent
    t.surf_embed = {"tactip": 0.0025, "digitac": 0.0015, "digit": 0.0015}[sensor]                 # :67-75
    t.surf_drive = 0.25 * {"tactip": 1.0, "digitac": 0.9, "digit": 0.7}[sensor]                   # surface_follow_auto_env.py:35-43
    t.surf_w_norm = 0.0 if movement_mode in ("yz", "xyz") else 1.0                                # :88-89
    t.surf_w_goal, t.surf_w_surf = 0.0, 1.0                                                      # :79-80, :92
    if variant == "goal":                                                                        # surface_follow_goal_env.py:62-81
        t.surf_drive, t.surf_w_goal, t.surf_w_surf = 0.0, 1.0, 10.0
    if variant == "vert":                                                                        # surface_follow_vert_env.py:63-79
        t.surf_drive_y_only, t.surf_w_goal, t.surf_w_surf, t.surf_w_norm = 1, 0.0, 10.0, 3.0
    t.n_draws = 2
    t.draw_default[0], t.draw_default[1] = 0.0, 0.0
    _program(t, [(L.TG_DRAW_RANDINT, 0.0, 1e8) if noise_mode in ("simplex", "vertical_simplex") else _CONST,   # :448, :458
                 (L.TG_DRAW_CHOICE_PM1, 0.0, 0.0) if one_d else _uni(-np.pi, np.pi)])                            # :508, :513
    dep, gray, mask = scene.load_refimg(sensor, typ, S)
    rest = scene.load_rest_pose("surface_follow", arm_type, sensor, typ, control_links)
    s = cfg.sensor
    s.image_size, s.border_on = S, 1                                                             # turn_off_border=False :137
    s.fov_deg, s.near_, s.far_ = sj["fov"], sj["near"], sj["far"]
    s.h_nodef_dep = dep.ctypes.data_as(C.POINTER(C.c_float))
    s.h_nodef_gray = gray.ctypes.data_as(C.POINTER(C.c_float))
    s.h_border_mask = mask.ctypes.data_as(C.POINTER(C.c_uint8))
    s.n_prim = 0                                                                                 # the stimulus is the per-env heightfield
    cfg.h_rest_q = rest.ctypes.data_as(C.POINTER(C.c_double))
    return cfg, (dep, gray, mask, rest), surface_follow_draws("simplex" if vertical else ("none" if noise_mode == "random" else noise_mode), one_d)


def object_push_draws(rand_init_orn, rand_obj_mass, traj_type, default_mass):
    """ObjectPushEnv draws per reset in the reference's call order: reset_object's `uniform(-pi/32, pi/32)` (if
    rand_init_orn) and `uniform(0.4, 0.8)` (if rand_obj_mass) (object_push_env.py:204-229), then make_goal ->
    update_trajectory's `randint(1e8)` (simplex, :289) or `uniform(-pi/8, pi/8)` (straight, :310)."""

    def draw(rng, rounds):
        out = np.empty((rounds, 3))
        for r in range(rounds):
            out[r, 0] = rng.uniform(-np.pi / 32, np.pi / 32) if rand_init_orn else 0.0
            out[r, 1] = rng.uniform(0.4, 0.8) if rand_obj_mass else default_mass
            out[r, 2] = rng.randint(1e8) if traj_type == "simplex" else rng.uniform(-np.pi / 8, np.pi / 8)
        return out

    return draw


def object_push_config(env_modes, image_size, max_steps, n_envs, lanes_per_warp=0):
    """ObjectPushEnv.__init__ (rl_envs/nonprehensile_manipulation/object_push/object_push_env.py:24-125) + BaseObjectEnv
    as a TgConfig.  Returns (cfg, keepalive, draw_fn)."""
    arm_type, sensor = env_modes["arm_type"], env_modes["tactile_sensor_name"]
    movement_mode, traj_type = env_modes["movement_mode"], env_modes.get("traj_type", "simplex")
    if arm_type not in ("ur5", "mg400"):
        raise ValueError("Incorrect arm type specified {}".format(arm_type))
    if traj_type not in ("simplex", "straight"):
        raise ValueError("Incorrect traj_type specified: {}".format(traj_type))          # :272
    if movement_mode not in ("y", "yRz", "xyRz", "TyRz", "TxTyRz"):
        raise ValueError("Incorrect movement_mode specified: {}".format(movement_mode))
    typ = "right_angle"                                                                    # :58
    if arm_type == "mg400" and sensor == "tactip":
        typ = "mini_right_angle"                                                           # :84-86 (the reference's own PPO set-up)
    S = int(image_size[0])
    mj = scene.load_model_json(arm_type, sensor, typ)
    sj = scene.load_sensor_json(sensor, typ)
    cam = sj["types"][typ]
    arm, control_links = scene.reduce_model(mj, sensor, cam["cam_pos"], cam["cam_rpy"])
    cfg = L.TgConfig()
    cfg.n_envs, cfg.lanes_per_warp, cfg.arm = n_envs, lanes_per_warp, arm
    cfg.phys = scene.default_physics(substeps=int(np.floor((1.0 / 10.0) / (1.0 / 240.0))))      # :36-38 -> 24
    t = cfg.task
    t.task, t.max_steps = L.TG_TASK_OBJECT_PUSH, int(max_steps)
    idx = {"y": [1], "yRz": [1, 5], "xyRz": [0, 1, 5], "TyRz": [-1, -1], "TxTyRz": [-1, -1, -1]}[movement_mode]   # :369-454
    t.push_mode = {"y": L.TG_PUSH_WORK_DRIVE, "yRz": L.TG_PUSH_WORK_DRIVE, "xyRz": L.TG_PUSH_WORK, "TyRz": L.TG_PUSH_TCP_TYRZ,
                   "TxTyRz": L.TG_PUSH_TCP_TXTYRZ}[movement_mode]
    t.act_dim = len(idx)
    for k in range(6):
        t.act_index[k] = idx[k] if k < len(idx) else -1
    t.act_min, t.act_max = -0.25, 0.25
    t.control_mode, t.pos_max_steps = _control_mode(env_modes, arm_type)                          # :39: _max_blocking_pos_move_steps = 10
    mv, ma = 0.01, 5.0 * (np.pi / 180)                                                          # :150-160
    if t.control_mode == 1:                                                                      # :137-147: m / rad per step
        mv, ma = 0.001, 1 * (np.pi / 180)
    hi = [mv, mv, 0.0, 0.0, 0.0, ma]
    for k in range(6):
        t.act_lo[k], t.act_hi[k] = -hi[k], hi[k]
    obj_w = obj_h = 0.08                                                                         # :54-55
    a45 = 45 * np.pi / 180
    if arm_type == "mg400":                                                                      # :72-88
        lims = [(0.0, 0.3), (-0.1, 0.08), (0.0, 0.0), (0.0, 0.0), (0.0, 0.0), (-a45, a45)]
        wd = [0.30 if sensor == "tactip" else 0.25, -0.1, obj_h / 2]                             # :82-88
    else:                                                                                        # :89-101
        lims = [(0.0, 0.3), (-0.1, 0.1), (0.0, 0.0), (0.0, 0.0), (0.0, 0.0), (-a45, a45)]
        wd = [0.55, -0.20, obj_h / 2]
    wf_rpy = [-np.pi, 0.0, np.pi / 2]                                                            # :105
    for k in range(3):
        t.workframe_pos[k], t.workframe_rpy[k], t.init_rpy[k] = wd[k], wf_rpy[k], 0.0
        t.obj_init_rpy[k] = [-np.pi, 0.0, np.pi / 2][k]                                          # :196, :210
        t.obj_base_com[k] = 0.0
        t.push_init_pos[k] = [wd[0], wd[1] + obj_w / 2, obj_h / 2][k]                            # :193
    for k in range(6):
        t.tcp_lims[k][0], t.tcp_lims[k][1] = lims[k]
    cube = scene.load_object_json("cube")
    box = [0.08, 0.08, 0.08]
    stiff, damp, fric = {"tactip": (50, 100, 10.0), "digitac": (300, 100, 10.0), "digit": (50, 200, 10.0)}[sensor]   # :61-66
    cube_mu, table_mu = 0.065, 1.0                                                               # :218, table.urdf
    for k in range(3):
        t.push_half[k] = box[k] / 2
        t.push_inertia_per_mass[k] = cube["inertia_diag"][k] / cube["mass"]
    t.push_table_z = 0.0
    t.push_mu_table, t.push_mu_tip = cube_mu * table_mu, min(cube_mu * fric, 10.0)               # [EXT] product, clamped at 10
    t.push_tip_k, t.push_tip_d = 1.0 / (1.0 / stiff + 1.0 / 1e18), damp + 0.1                    # [EXT] combined with bullet's defaults
    t.push_erp, t.push_slop = 0.2, 1e-4
    t.push_lin_damping, t.push_ang_damping = 0.04, 0.04
    t.push_term_dist = 0.025                                                                     # :68
    t.push_traj_spacing, t.push_traj_perturb = 0.025, 0.1                                        # :234-235
    t.push_traj_offset = obj_w / 2 + 0.025                                                       # :290
    t.push_traj_straight = 1 if traj_type == "straight" else 0
    t.push_sparse_reward = 1 if env_modes.get("reward_mode", "dense") == "sparse" else 0
    t.n_draws = 3
    t.draw_default[0], t.draw_default[1], t.draw_default[2] = 0.0, cube["mass"], 0.0
    _program(t, [_uni(-np.pi / 32, np.pi / 32) if env_modes.get("rand_init_orn", False) else _CONST,    # :204-229
                 _uni(0.4, 0.8) if env_modes.get("rand_obj_mass", False) else _CONST,
                 (L.TG_DRAW_RANDINT, 0.0, 1e8) if traj_type == "simplex" else _uni(-np.pi / 8, np.pi / 8)])   # :289, :310
    dep, gray, mask = scene.load_refimg(sensor, typ, S)
    tris = scene.load_stimulus("cube")
    rest = scene.load_rest_pose("object_push", arm_type, sensor, typ, control_links)
    hull_link = scene.load_tip_hull(arm_type, sensor, typ)
    hb, hull = scene.link_points_in_body(arm, sensor + "_tip_link", hull_link)
    if hb != arm.tcp_body:
        raise ValueError("the tip link and the TCP must ride on the same body")
    s = cfg.sensor
    s.image_size, s.border_on = S, 1
    s.fov_deg, s.near_, s.far_ = sj["fov"], sj["near"], sj["far"]
    s.h_nodef_dep = dep.ctypes.data_as(C.POINTER(C.c_float))
    s.h_nodef_gray = gray.ctypes.data_as(C.POINTER(C.c_float))
    s.h_border_mask = mask.ctypes.data_as(C.POINTER(C.c_uint8))
    prims, prim_nv = scene.merge_coplanar(tris)
    s.n_prim = len(prims)
    s.h_prims = prims.ctypes.data_as(C.POINTER(C.c_double))
    s.h_prim_nv = prim_nv.ctypes.data_as(C.POINTER(C.c_int32))
    parts, part_cen = scene.convex_parts(prims, prim_nv)           # boxes: the scanline raster applies
    if parts is not None:
        part_cen = np.ascontiguousarray(part_cen, dtype=np.float64)
        s.h_prim_part = parts.ctypes.data_as(C.POINTER(C.c_int32))
        s.h_part_centroid = part_cen.ctypes.data_as(C.POINTER(C.c_double))
        s.n_parts = len(part_cen)
    cfg.h_rest_q = rest.ctypes.data_as(C.POINTER(C.c_double))
    cfg.h_tip_hull = hull.ctypes.data_as(C.POINTER(C.c_double))
    cfg.n_tip_hull = len(hull)
    draw = object_push_draws(bool(env_modes.get("rand_init_orn", False)), bool(env_modes.get("rand_obj_mass", False)), traj_type, cube["mass"])
    return cfg, (dep, gray, mask, tris, rest, prims, prim_nv, hull, parts, part_cen), draw


def object_roll_draws(rand_obj_size, rand_embed_dist, rand_init_obj_pos):
    """ObjectRollEnv draws per reset in the reference's call order: reset_task's `uniform(1, 2)` (if rand_obj_size) and
    `uniform(0.0015, 0.003)` (if rand_embed_dist) (object_roll_env.py:182-195), reset_object's two `uniform(-0.009, 0.009)`
    (if rand_init_obj_pos, :209-216), make_goal's `uniform(-pi, pi)` and `uniform(0 | 0.005, 0.015)` (:244-250)."""

    def draw(rng, rounds):
        out = np.empty((rounds, 6))
        for r in range(rounds):
            out[r, 0] = rng.uniform(1.0, 2.0) if rand_obj_size else 1.0
            out[r, 1] = rng.uniform(0.0015, 0.003) if rand_embed_dist else 0.0015
            out[r, 2] = rng.uniform(-0.009, 0.009) if rand_init_obj_pos else 0.0
            out[r, 3] = rng.uniform(-0.009, 0.009) if rand_init_obj_pos else 0.0
            out[r, 4] = rng.uniform(-np.pi, np.pi)
            out[r, 5] = rng.uniform(0.0, 0.015) if rand_init_obj_pos else rng.uniform(0.005, 0.015)
        return out

    return draw


def object_roll_config(env_modes, image_size, max_steps, n_envs, lanes_per_warp=0):
    """ObjectRollEnv.__init__ (rl_envs/nonprehensile_manipulation/object_roll/object_roll_env.py:23-104) + BaseObjectEnv as a
    TgConfig.  Returns (cfg, keepalive, draw_fn)."""
    arm_type, sensor = env_modes["arm_type"], env_modes["tactile_sensor_name"]
    if env_modes["movement_mode"] != "xy":
        raise ValueError("Incorrect movement_mode specified: {}".format(env_modes["movement_mode"]))      # :288-297 knows "xy" only
    if arm_type != "ur5":
        raise ValueError("object_roll has rest poses for the ur5 only among the built arms (rest_poses.py)")
    if sensor != "tactip":
        raise NotImplementedError("object_roll uses the flat TacTip (t_s_type 'flat', :58); %r has no flat variant" % sensor)
    typ, S = "flat", int(image_size[0])                                                          # :58
    mj = scene.load_model_json(arm_type, sensor, typ)
    sj = scene.load_sensor_json(sensor, typ)
    cam = sj["types"][typ]
    arm, control_links = scene.reduce_model(mj, sensor, cam["cam_pos"], cam["cam_rpy"])
    cfg = L.TgConfig()
    cfg.n_envs, cfg.lanes_per_warp, cfg.arm = n_envs, lanes_per_warp, arm
    cfg.phys = scene.default_physics(substeps=int(np.floor((1.0 / 10.0) / (1.0 / 240.0))))      # :34-36 -> 24
    t = cfg.task
    t.task, t.max_steps = L.TG_TASK_OBJECT_ROLL, int(max_steps)
    t.act_dim = 2
    for k in range(6):
        t.act_index[k] = [0, 1][k] if k < 2 else -1                                              # :288-297
    t.act_min, t.act_max = -0.25, 0.25
    t.control_mode, t.pos_max_steps = _control_mode(env_modes, arm_type)                          # :37: _max_blocking_pos_move_steps = 10
    mv = 0.001 if t.control_mode == 1 else 0.01                                                  # :112-139
    hi = [mv, mv, 0.0, 0.0, 0.0, 0.0]
    for k in range(6):
        t.act_lo[k], t.act_hi[k] = -hi[k], hi[k]
    radius, embed = 0.0025, 0.0015                                                               # :166, :67
    wf = [0.65, 0.0, 2 * radius - embed]                                                         # :72 (z is per episode on the device)
    lims = [(-0.05, 0.05), (-0.05, 0.05), (-0.01, 0.01), (0.0, 0.0), (0.0, 0.0), (0.0, 0.0)]     # :75-81
    for k in range(3):
        t.workframe_pos[k], t.workframe_rpy[k], t.init_rpy[k] = wf[k], [-np.pi, 0.0, np.pi / 2][k], 0.0
        t.push_init_pos[k] = [0.65, 0.0, radius][k]                                              # :167
        t.obj_init_rpy[k], t.obj_base_com[k] = 0.0, 0.0
        t.push_half[k], t.push_inertia_per_mass[k] = radius, 0.4 * radius * radius               # unused by the sphere narrow phase
    for k in range(6):
        t.tcp_lims[k][0], t.tcp_lims[k][1] = lims[k]
    t.push_shape = 1
    t.roll_radius = radius
    t.obj_mass = 0.05                                                                            # sphere.urdf
    cyl = mj.get("tip_collision")
    if not cyl or cyl["type"] != "cylinder":
        raise ValueError("the flat tip's collision primitive is missing from the compiled model")
    ax_link = scene.rpy_to_mat(cyl["rpy"]) @ np.array([0.0, 0.0, 1.0])
    pts_link = np.array([cyl["xyz"], np.array(cyl["xyz"]) + ax_link])
    hb, pts = scene.link_points_in_body(arm, sensor + "_tip_link", pts_link)
    if hb != arm.tcp_body:
        raise ValueError("the tip link and the TCP must ride on the same body")
    for k in range(3):
        t.roll_cyl_pos[k], t.roll_cyl_axis[k] = pts[0][k], (pts[1] - pts[0])[k]
    t.roll_cyl_half_len, t.roll_cyl_radius = cyl["length"] / 2, cyl["radius"]
    stiff, damp, fric = 10.0, 100, 10.0                                                          # t_s_dynamics :62
    sphere_mu, table_mu = 10.0, 1.0                                                              # :226, table.urdf
    t.push_table_z = 0.0
    t.push_mu_table, t.push_mu_tip = min(sphere_mu * table_mu, 10.0), min(sphere_mu * fric, 10.0)  # [EXT] product, clamped at 10
    t.push_tip_k, t.push_tip_d = 1.0 / (1.0 / stiff + 1.0 / 1e18), damp + 0.1
    t.push_erp, t.push_slop = 0.2, 1e-4
    t.push_lin_damping, t.push_ang_damping = 0.04, 0.04
    t.push_term_dist = 0.001                                                                     # :64
    t.push_sparse_reward = 1 if env_modes.get("reward_mode", "dense") == "sparse" else 0
    t.n_draws = 6
    for k, v in enumerate([1.0, embed, 0.0, 0.0, 0.0, 0.01]):
        t.draw_default[k] = v
    rp = bool(env_modes.get("rand_init_obj_pos", False))
    _program(t, [_uni(1.0, 2.0) if env_modes.get("rand_obj_size", False) else _CONST,                   # :182-195
                 _uni(0.0015, 0.003) if env_modes.get("rand_embed_dist", False) else _CONST,
                 _uni(-0.009, 0.009) if rp else _CONST, _uni(-0.009, 0.009) if rp else _CONST,           # :209-216
                 _uni(-np.pi, np.pi), _uni(0.0, 0.015) if rp else _uni(0.005, 0.015)])                   # :244-250
    dep, gray, mask = scene.load_refimg(sensor, typ, S)
    rest = scene.load_rest_pose("object_roll", arm_type, sensor, typ, control_links)
    s = cfg.sensor
    s.image_size, s.border_on = S, 1
    s.fov_deg, s.near_, s.far_ = sj["fov"], sj["near"], sj["far"]
    s.h_nodef_dep = dep.ctypes.data_as(C.POINTER(C.c_float))
    s.h_nodef_gray = gray.ctypes.data_as(C.POINTER(C.c_float))
    s.h_border_mask = mask.ctypes.data_as(C.POINTER(C.c_uint8))
    s.n_prim = 0                                                                                 # the stimulus is the per-env sphere
    cfg.h_rest_q = rest.ctypes.data_as(C.POINTER(C.c_double))
    draw = object_roll_draws(bool(env_modes.get("rand_obj_size", False)), bool(env_modes.get("rand_embed_dist", False)),
                             bool(env_modes.get("rand_init_obj_pos", False)))
    return cfg, (dep, gray, mask, rest), draw


class TactileWorld:
    """N envs of one task on one device."""

    def __init__(self, cfg, keepalive, device=0, draw_fn=None, rng=None):
        """rng: "host" - reset draws come from one numpy RandomState per env on the host, streamed through the device ring;
        "device" - every env carries that RandomState's MT19937 state on the device and draws there (same sequence, no host in
        the loop; csrc/tg_rng.cuh; bit-identical episodes, tests/test_gpu_rng.py).  Default: TG_RNG or "device"; surface_follow's
        noise_mode "random" always runs on "device"; explicit set_draws() (parity tests) switches to the ring."""
        import torch

        if not torch.cuda.is_available():
            raise L.TgError("no CUDA device visible: tactile_gym_b200 has no CPU fallback")
        self.torch = torch
        self.lib = L.load()
        self.cfg, self._keep = cfg, keepalive
        self.device = torch.device("cuda", device)
        self.n, self.S, self.act_dim = cfg.n_envs, cfg.sensor.image_size, cfg.task.act_dim
        self.nb = cfg.arm.nb
        h = C.c_void_p()
        L.check(self.lib.tg_create(C.byref(cfg), device, C.byref(h)))
        self.h = h
        with torch.cuda.device(self.device):
            # the observation tensors [N, S, S, 1] u8 are allocated on first use: observation_mode "oracle" never renders,
            # and at 16,384 envs x 256 x 256 they are 1 GiB each
            self._obs = self._term_obs = None
            self.reward = torch.zeros(self.n, dtype=torch.float32, device=self.device)
            self.done = torch.zeros(self.n, dtype=torch.uint8, device=self.device)
            self.feat = self.term_feat = None
            self.nfeat = {L.TG_TASK_OBJECT_PUSH: 12, L.TG_TASK_OBJECT_ROLL: 3}.get(cfg.task.task, 0)
            if cfg.task.task == L.TG_TASK_SURFACE_FOLLOW and (cfg.task.surf_w_goal != 0.0 or cfg.task.surf_drive_y_only):
                self.nfeat = 6                   # surface_follow-v1 / -v2: TCP + goal position (surface_follow_goal_env.py:83-97)
            if self.nfeat:
                # extended_feature (object_push_env.py:611-629, object_roll_env.py:402-408), filled by every step / reset;
                # the first `nfeat` columns are meaningful
                self.feat = torch.zeros((self.n, L.TG_PUSH_NFEAT), dtype=torch.float32, device=self.device)
                self.term_feat = torch.zeros_like(self.feat)
                L.check(self.lib.tg_bind_features(self.h, self.feat.data_ptr(), self.term_feat.data_ptr()))
        # observation_mode "oracle" (get_oracle_obs): length of the task's state vector; buffers bound on demand
        self.n_oracle = {L.TG_TASK_EDGE_FOLLOW: 10, L.TG_TASK_SURFACE_FOLLOW: 20, L.TG_TASK_OBJECT_BALANCE: 26,
                         L.TG_TASK_OBJECT_PUSH: 30, L.TG_TASK_OBJECT_ROLL: 34}[cfg.task.task]
        self.oracle_obs = self.term_oracle_obs = None
        self._draw = draw_fn if draw_fn is not None else edge_follow_draws(cfg.task)
        self._rngs = [seeding.np_random(None)[0] for _ in range(self.n)]
        self.rng_mode = rng or os.environ.get("TG_RNG", "device")
        if cfg.task.task == L.TG_TASK_SURFACE_FOLLOW and cfg.task.surf_mode == 4:
            self.rng_mode = "device"
        if self.rng_mode not in ("host", "device"):
            raise ValueError("rng must be 'host' or 'device', got %r" % (self.rng_mode,))
        self._pin_ring = self._pin_avail = self._pin_counts = None
        self._poll_ev = self._upload_ev = None
        self._managed, self._since_poll = True, 0
        self._start_draws()

    @property
    def obs(self):
        if self._obs is None:
            with self.torch.cuda.device(self.device):
                self._obs = self.torch.zeros((self.n, self.S, self.S, 1), dtype=self.torch.uint8, device=self.device)
        return self._obs

    @property
    def term_obs(self):
        if self._term_obs is None:
            with self.torch.cuda.device(self.device):
                self._term_obs = self.torch.zeros((self.n, self.S, self.S, 1), dtype=self.torch.uint8, device=self.device)
        return self._term_obs

    def bind_oracle_obs(self):
        """observation_mode "oracle": every following step / reset fills self.oracle_obs [N, TG_ORACLE_NOBS] (first
        n_oracle columns meaningful) and, for finished envs, self.term_oracle_obs."""
        if self.oracle_obs is None:
            torch = self.torch
            self.oracle_obs = torch.zeros((self.n, L.TG_ORACLE_NOBS), dtype=torch.float32, device=self.device)
            self.term_oracle_obs = torch.zeros_like(self.oracle_obs)
            L.check(self.lib.tg_bind_oracle_obs(self.h, self.oracle_obs.data_ptr(), self.term_oracle_obs.data_ptr()))
        return self.oracle_obs

    # ------------------------------------------------------------------ seeding / draws
    def seed(self, seeds):
        """seeds: one int (env i gets seed + i, the make_vec_env convention) or a list of per-env seeds."""
        if seeds is None or isinstance(seeds, (int, np.integer)):
            seeds = [None if seeds is None else int(seeds) + i for i in range(self.n)]
        out = []
        for i, s in enumerate(seeds):
            self._rngs[i], sd = seeding.np_random(s)
            out.append(sd)
        self._start_draws()
        return out

    def _start_draws(self):
        """A fresh draw sequence (construction, seed()): DRAW_ROUNDS draws per env go up synchronously and the pre-computed next
        episodes are recomputed.  The host copy is a pinned RING [N, DRAW_ROUNDS, n_draws]: the k-th reset of env i reads slot
        k % DRAW_ROUNDS, `_avail[i]` counts the draws produced so far."""
        torch = self.torch
        nd = self.cfg.task.n_draws
        if self.rng_mode == "device":
            # hand the generators themselves to the device: numpy's MT19937 key + position per env
            keys = np.empty((self.n, L.MT_N), dtype=np.uint32)
            pos = np.empty(self.n, dtype=np.int32)
            for i, r in enumerate(self._rngs):
                st = r.get_state()
                keys[i], pos[i] = st[1], st[2]
            L.check(self.lib.tg_set_rng_state(self.h, keys.ctypes.data, pos.ctypes.data))
            self._managed = False
            return
        if self._pin_ring is None:
            self._pin_ring = torch.zeros((self.n, DRAW_ROUNDS, nd), dtype=torch.float64).pin_memory()
            self._pin_avail = torch.zeros(self.n, dtype=torch.int32).pin_memory()
            self._pin_counts = torch.zeros(self.n + 1, dtype=torch.int32).pin_memory()
        if self._upload_ev is not None:
            self._upload_ev.synchronize()
        ring = self._pin_ring.numpy()
        for i, r in enumerate(self._rngs):
            ring[i] = self._draw(r, DRAW_ROUNDS)
        self._pin_avail.fill_(DRAW_ROUNDS)
        L.check(self.lib.tg_set_draws(self.h, self._pin_ring.data_ptr(), DRAW_ROUNDS))
        self._managed, self._poll_ev, self._since_poll = True, None, 0

    def _feed_draws(self, wait=False):
        """Called once per step / reset; never synchronises in steady state.  Every DRAW_CHECK_EVERY calls the per-env reset
        counters are copied out asynchronously; when that copy has landed, the slots of consumed draws are refilled from the
        envs' own RandomStates (episode order is kept: draw k of env i is always the k-th output of its stream) and the ring
        goes back up, asynchronously too.  Only if the device ran so far ahead of the host that the ring could run dry before
        the counters arrive does this wait for them."""
        if not self._managed:
            return
        torch = self.torch
        self._since_poll += 1
        if self._poll_ev is None:
            if self._since_poll < DRAW_CHECK_EVERY and not wait:
                return
            L.check(self.lib.tg_draws_poll(self.h, self._pin_counts.data_ptr(), self._stream()))
            self._poll_ev = torch.cuda.Event()
            self._poll_ev.record(torch.cuda.current_stream(self.device))
            if not wait:
                return
        if not self._poll_ev.query():
            if not wait and self._since_poll < DRAW_CHECK_EVERY + DRAW_ROUNDS // 8:
                return
            self._poll_ev.synchronize()
        self._poll_ev, self._since_poll = None, 0
        counts = self._pin_counts.numpy()
        if int(counts[self.n]) & 2:
            raise L.TgError("reset draws ran out on the device (an env reset with TgTask.draw_default): the host fell behind refilling the ring")
        consumed = counts[: self.n].astype(np.int64)
        avail = self._pin_avail.numpy()
        free = consumed + DRAW_ROUNDS - avail            # slots whose draw has been consumed
        if free.max() < DRAW_ROUNDS // 4:
            return                                       # plenty left; a refill costs a few us per env on the host
        if self._upload_ev is not None:
            self._upload_ev.synchronize()                # the previous upload read the pinned ring (long done)
        ring = self._pin_ring.numpy()
        for i in np.nonzero(free > 0)[0]:
            c, a = int(free[i]), int(avail[i])
            ring[i, (a + np.arange(c)) % DRAW_ROUNDS] = self._draw(self._rngs[i], c)
            avail[i] = a + c
        L.check(self.lib.tg_draws_upload(self.h, self._pin_ring.data_ptr(), self._pin_avail.data_ptr(), self._stream()))
        self._upload_ev = torch.cuda.Event()
        self._upload_ev.record(torch.cuda.current_stream(self.device))

    def set_draws(self, draws):
        """Explicit draws [N, rounds, n_draws] (parity tests): the r-th reset of env i uses draws[i, r]; nothing is refilled."""
        arr = np.ascontiguousarray(draws, dtype=np.float64)
        assert arr.shape[0] == self.n and arr.shape[2] == self.cfg.task.n_draws, arr.shape
        L.check(self.lib.tg_set_draws(self.h, arr.ctypes.data, arr.shape[1]))
        self._managed = False

    # ------------------------------------------------------------------ stepping
    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def reset(self, mask=None, render=True):
        """Reset the masked envs (None: all).  render=False (observation_mode "oracle") skips their first image."""
        mp = None
        if mask is not None:
            mask = mask.to(device=self.device, dtype=self.torch.uint8).contiguous()
            mp = mask.data_ptr()
        if not render:
            L.check(self.lib.tg_reset_only(self.h, mp, self._stream()))
        else:
            L.check(self.lib.tg_reset(self.h, mp, self.obs.data_ptr(), self._stream()))
        self._feed_draws()
        return self.obs if render else None

    def step(self, actions, want_terminal_obs=False, out=None):
        """actions: float32 cuda tensor [N, act_dim].  Returns (obs, reward, done) tensors (no sync).
        out = (obs u8 [N,S,S,1], reward f32 [N], done u8 [N], feat f32 [N,TG_PUSH_NFEAT] | None): write this step's results
        into the caller's tensors instead of the world's own (distributed.CollatedBatch hands out views of its packed
        all-gather buffer, so the kernels write straight into the send slot)."""
        a = actions
        if a.device != self.device or a.dtype != self.torch.float32 or not a.is_contiguous():
            a = a.to(device=self.device, dtype=self.torch.float32).contiguous()
        if a.shape != (self.n, self.act_dim):
            raise ValueError("actions must have shape (%d, %d), got %s" % (self.n, self.act_dim, tuple(a.shape)))
        obs, reward, done = (self.obs, self.reward, self.done) if out is None else out[:3]
        if out is not None and self.nfeat:
            if out[3] is None or tuple(out[3].shape) != (self.n, L.TG_PUSH_NFEAT):
                raise ValueError("this task writes an extended feature: out[3] must be float32 [N, %d]" % L.TG_PUSH_NFEAT)
            L.check(self.lib.tg_bind_features(self.h, out[3].data_ptr(), self.term_feat.data_ptr()))
        L.check(self.lib.tg_step(self.h, a.data_ptr(), obs.data_ptr(), reward.data_ptr(), done.data_ptr(),
                                 self.term_obs.data_ptr() if want_terminal_obs else None, self._stream()))
        if out is not None and self.nfeat:
            L.check(self.lib.tg_bind_features(self.h, self.feat.data_ptr(), self.term_feat.data_ptr()))
        self._feed_draws()
        return obs, reward, done

    def step_host(self, h_actions, h_obs, h_reward, h_done, h_feat=None, h_oracle=None, want_terminal_obs=True, chunks=0, term=None):
        """One env step with pinned HOST tensors in and out (tg_step_host): the observation is rendered and copied out in
        chunks so the PCIe transfer overlaps the raster.  h_obs None (observation_mode "oracle"): nothing is rendered.
        Valid after synchronising the current stream.  term = (h_term_obs [cap, S, S, 1] u8, h_term_idx [1 + cap] i32 (count first),
        h_term_feat [cap, 12] f32 or None), pinned: the finished envs' terminal observations arrive compacted with the same call."""
        hs = L.TgHostStep()
        hs.h_actions, hs.h_reward, hs.h_done = h_actions.data_ptr(), h_reward.data_ptr(), h_done.data_ptr()
        hs.d_reward, hs.d_done = self.reward.data_ptr(), self.done.data_ptr()
        if h_obs is not None:
            hs.h_obs, hs.d_obs = h_obs.data_ptr(), self.obs.data_ptr()
            hs.d_term_obs = self.term_obs.data_ptr() if want_terminal_obs else None
        if h_feat is not None:
            hs.d_feat, hs.h_feat = self.feat.data_ptr(), h_feat.data_ptr()
        if h_oracle is not None:
            self.bind_oracle_obs()
            hs.h_oracle = h_oracle.data_ptr()
        hs.chunks = chunks
        if term is not None and h_obs is not None and want_terminal_obs:
            t_obs, t_idx, t_feat = term
            hs.h_term_obs, hs.h_term_idx, hs.term_cap = t_obs.data_ptr(), t_idx.data_ptr(), int(t_obs.shape[0])
            if t_feat is not None and self.term_feat is not None:
                hs.h_term_feat = t_feat.data_ptr()
        L.check(self.lib.tg_step_host(self.h, C.byref(hs), self._stream()))
        self._feed_draws()

    def physics_only(self, actions):
        L.check(self.lib.tg_physics_only(self.h, actions.data_ptr(), self.reward.data_ptr(), self.done.data_ptr(), self._stream()))
        self._feed_draws()

    def raster_only(self):
        L.check(self.lib.tg_raster_only(self.h, self.obs.data_ptr(), self._stream()))
        return self.obs

    # ------------------------------------------------------------------ state
    def state_size(self):
        return self.lib.tg_state_size(self.h)

    def get_state(self):
        st = np.zeros((self.n, self.state_size()))
        L.check(self.lib.tg_get_state(self.h, st.ctypes.data, self._stream()))
        return st

    def set_state(self, st):
        st = np.ascontiguousarray(st, dtype=np.float64)
        L.check(self.lib.tg_set_state(self.h, st.ctypes.data, self._stream()))

    def get_camera(self):
        cam = np.zeros((self.n, 12))
        L.check(self.lib.tg_get_camera(self.h, cam.ctypes.data, self._stream()))
        return cam

    def pipeline_error(self):
        """sticky error flag of the reset pipeline / heightfield raster (bit 0 of tg_pipeline_error)"""
        return bool(self.lib.tg_pipeline_error(self.h, self._stream()) & 1)

    def draws_exhausted(self):
        """a reset found the draw ring empty and used the task's default draws (bit 1 of tg_pipeline_error).  Expected with
        explicit set_draws() once its rounds are used up (the pipeline pre-computes one episode ahead); an error otherwise."""
        return bool(self.lib.tg_pipeline_error(self.h, self._stream()) & 2)

    def scan_fallbacks(self):
        """envs of the last raster pass that the scanline raster handed to the general raster kernel (profiling counter)"""
        return int(self.lib.tg_scan_fallbacks(self.h, self._stream()))

    def scan_fallback_reasons(self):
        """per reason code: [kept, near-plane cut, eye inside a part, too many front faces, test hook, ...] of the last raster pass"""
        c = np.zeros(8, dtype=np.int32)
        L.check(self.lib.tg_scan_fallback_reasons(self.h, c.ctypes.data, self._stream()))
        return c

    def nan_resets(self):
        """env steps whose state came out non-finite (reported as done, reward 0; the env then starts its next episode)"""
        return int(self.lib.tg_nan_resets(self.h, self._stream()))

    def save_checkpoint(self):
        """Everything needed to resume this world bit for bit (tg_checkpoint_save: every device buffer incl. the pre-computed next
        episodes, heightfields and RNG states) plus the host side of the draw stream.  Returns a dict of numpy arrays / plain
        values (picklable, np.savez-able)."""
        nbytes = int(self.lib.tg_checkpoint_bytes(self.h))
        blob = np.empty(nbytes, dtype=np.uint8)
        L.check(self.lib.tg_checkpoint_save(self.h, blob.ctypes.data, nbytes, self._stream()))
        ck = {"blob": blob, "rng_mode": self.rng_mode, "managed": bool(self._managed), "reward": self.reward.cpu().numpy(),
              "done": self.done.cpu().numpy(), "rngs": [r.get_state() for r in self._rngs]}
        if self.feat is not None:
            ck["feat"] = self.feat.cpu().numpy()
        if self._managed and self._pin_ring is not None:
            # host-drawn ring: what is on the device is in the blob; the host copy and its counters ride along
            if self._upload_ev is not None:
                self._upload_ev.synchronize()
            ck["ring"], ck["avail"] = self._pin_ring.numpy().copy(), self._pin_avail.numpy().copy()
        return ck

    def load_checkpoint(self, ck):
        """Resume from save_checkpoint() of a world built from the same configuration."""
        torch = self.torch
        if ck["rng_mode"] != self.rng_mode:
            raise ValueError("checkpoint was taken with rng %r, this world runs %r" % (ck["rng_mode"], self.rng_mode))
        blob = np.ascontiguousarray(ck["blob"], dtype=np.uint8)
        L.check(self.lib.tg_checkpoint_load(self.h, blob.ctypes.data, blob.nbytes, self._stream()))
        self.reward.copy_(torch.from_numpy(np.asarray(ck["reward"])))
        self.done.copy_(torch.from_numpy(np.asarray(ck["done"])))
        if self.feat is not None and "feat" in ck:
            self.feat.copy_(torch.from_numpy(np.asarray(ck["feat"])))
        for r, st in zip(self._rngs, ck["rngs"]):
            r.set_state(st)
        self._managed = bool(ck["managed"])
        if self._managed and "ring" in ck:
            if self._pin_ring is None:
                nd = self.cfg.task.n_draws
                self._pin_ring = torch.zeros((self.n, DRAW_ROUNDS, nd), dtype=torch.float64).pin_memory()
                self._pin_avail = torch.zeros(self.n, dtype=torch.int32).pin_memory()
                self._pin_counts = torch.zeros(self.n + 1, dtype=torch.int32).pin_memory()
            self._pin_ring.numpy()[...] = ck["ring"]
            self._pin_avail.numpy()[...] = ck["avail"]
            self._poll_ev, self._since_poll = None, 0
        if self._obs is not None:
            self.raster_only()                      # the live observation is a function of the restored state

    def pipeline_stalls(self):
        """episode ends that had to finish their pre-computed next episode inline (exact, just slower)"""
        return int(self.lib.tg_pipeline_stalls(self.h, self._stream()))

    def launch_count(self):
        return int(self.lib.tg_launch_count(self.h))

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.tg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
