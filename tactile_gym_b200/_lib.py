"""ctypes binding of libtactile_gym_b200.so (include/tactile_gym_b200.h).

There is no CPU fallback: if the CUDA library is missing or no GPU is visible, creating a world raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TG_LIB_OVERRIDE") or os.path.join(HERE, "libtactile_gym_b200.so")   # override: diagnostic builds (tools/)

TG_MAXB, TG_MAXSUB, TG_MAXTRI, TG_MAXDRAW = 8, 16, 64, 8
TG_TOPO_CHAIN6, TG_TOPO_MG400 = 0, 1
TG_TASK_EDGE_FOLLOW, TG_TASK_OBJECT_BALANCE, TG_TASK_SURFACE_FOLLOW, TG_TASK_OBJECT_PUSH, TG_TASK_OBJECT_ROLL = 0, 1, 2, 3, 4
TG_PUSH_NTRAJ, TG_PUSH_NFEAT = 10, 12
TG_ORACLE_NOBS = 36
TG_PUSH_WORK, TG_PUSH_WORK_DRIVE, TG_PUSH_TCP_TYRZ, TG_PUSH_TCP_TXTYRZ = 0, 1, 2, 3
TG_DRAW_CONST, TG_DRAW_UNIFORM, TG_DRAW_RANDINT, TG_DRAW_CHOICE_PM1, TG_DRAW_CHOICE_RAND = 0, 1, 2, 3, 4
MT_N = 624

D3 = C.c_double * 3
D9 = C.c_double * 9


class TgArm(C.Structure):
    _fields_ = [
        ("topo", C.c_int32), ("nb", C.c_int32), ("nsub", C.c_int32), ("pad0", C.c_int32),
        ("jpos", D3 * TG_MAXB), ("jrot", D9 * TG_MAXB), ("axis", D3 * TG_MAXB),
        ("mass", C.c_double * TG_MAXB), ("com", D3 * TG_MAXB), ("inertia", (C.c_double * 6) * TG_MAXB),
        ("sub_start", C.c_int32 * (TG_MAXB + 1)), ("pad1", C.c_int32),
        ("sub_body", C.c_int32 * TG_MAXSUB), ("sub_mass", C.c_double * TG_MAXSUB), ("sub_com", D3 * TG_MAXSUB),
        ("sub_rot", D9 * TG_MAXSUB), ("sub_inertia", D3 * TG_MAXSUB),
        ("tcp_body", C.c_int32), ("cam_body", C.c_int32),
        ("tcp_pos", D3), ("tcp_rot", D9), ("cam_pos", D3), ("cam_rot", D9),
    ]


class TgPhysics(C.Structure):
    _fields_ = [
        ("gravity", D3), ("dt", C.c_double), ("solver_iters", C.c_int32), ("substeps", C.c_int32),
        ("lin_damping", C.c_double), ("ang_damping", C.c_double), ("joint_damping", C.c_double),
        ("max_force", C.c_double), ("pos_gain", C.c_double), ("vel_gain", C.c_double), ("solver_residual_threshold", C.c_double), ("blocking_force", C.c_double),
        ("gravity_comp", C.c_int32), ("pad0", C.c_int32),
    ]


class TgTask(C.Structure):
    _fields_ = [
        ("task", C.c_int32), ("act_dim", C.c_int32), ("max_steps", C.c_int32), ("n_draws", C.c_int32),
        ("act_index", C.c_int32 * 6), ("act_min", C.c_double), ("act_max", C.c_double),
        ("act_lo", C.c_double * 6), ("act_hi", C.c_double * 6),
        ("workframe_pos", D3), ("workframe_rpy", D3), ("tcp_lims", (C.c_double * 2) * 6),
        ("edge_pos", D3), ("edge_len", C.c_double), ("edge_height", C.c_double), ("termination_dist", C.c_double),
        ("embed_lo", C.c_double), ("embed_hi", C.c_double), ("init_rpy", D3), ("draw_default", C.c_double * TG_MAXDRAW),
        ("obj_mass", C.c_double), ("obj_inertia", D3), ("obj_com_off", D3), ("obj_base_com", D3), ("obj_init_rpy", D3),
        ("obj_base_w", C.c_double), ("obj_base_h", C.c_double), ("obj_force", C.c_double),
        ("obj_term_deg", C.c_double), ("obj_term_pos", C.c_double), ("p2p_erp", C.c_double), ("p2p_max_impulse", C.c_double),
        ("surf_pos", D3), ("surf_grid", C.c_double), ("surf_range", C.c_double), ("surf_interp", C.c_double),
        ("surf_extent", C.c_double), ("surf_embed", C.c_double), ("surf_drive", C.c_double), ("surf_w_norm", C.c_double),
        ("surf_w_goal", C.c_double), ("surf_w_surf", C.c_double),
        ("push_mode", C.c_int32), ("push_traj_straight", C.c_int32), ("push_sparse_reward", C.c_int32), ("push_shape", C.c_int32),
        ("sparse_reward", C.c_int32), ("surf_mode", C.c_int32), ("surf_dir_mode", C.c_int32), ("surf_drive_y_only", C.c_int32),
        ("control_mode", C.c_int32), ("pos_max_steps", C.c_int32),
        ("surf_vertical", C.c_int32), ("pad_vertical", C.c_int32),
        ("push_half", D3), ("push_table_z", C.c_double), ("push_mu_table", C.c_double), ("push_mu_tip", C.c_double),
        ("push_tip_k", C.c_double), ("push_tip_d", C.c_double), ("push_erp", C.c_double), ("push_slop", C.c_double),
        ("push_lin_damping", C.c_double), ("push_ang_damping", C.c_double), ("push_init_pos", D3), ("push_inertia_per_mass", D3),
        ("push_term_dist", C.c_double), ("push_traj_spacing", C.c_double), ("push_traj_perturb", C.c_double), ("push_traj_offset", C.c_double),
        ("roll_radius", C.c_double), ("roll_cyl_pos", D3), ("roll_cyl_axis", D3), ("roll_cyl_half_len", C.c_double), ("roll_cyl_radius", C.c_double),
        ("draw_kind", C.c_int32 * TG_MAXDRAW), ("draw_lo", C.c_double * TG_MAXDRAW), ("draw_hi", C.c_double * TG_MAXDRAW),
    ]


class TgSensor(C.Structure):
    _fields_ = [
        ("image_size", C.c_int32), ("border_on", C.c_int32),
        ("fov_deg", C.c_double), ("near_", C.c_double), ("far_", C.c_double),
        ("h_nodef_dep", C.POINTER(C.c_float)), ("h_nodef_gray", C.POINTER(C.c_float)), ("h_border_mask", C.POINTER(C.c_uint8)),
        ("n_prim", C.c_int32), ("pad0", C.c_int32), ("h_prims", C.POINTER(C.c_double)), ("h_prim_nv", C.POINTER(C.c_int32)),
        ("h_prim_part", C.POINTER(C.c_int32)), ("h_part_centroid", C.POINTER(C.c_double)), ("n_parts", C.c_int32), ("pad1", C.c_int32),
    ]


class TgConfig(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int32), ("lanes_per_warp", C.c_int32),
        ("arm", TgArm), ("phys", TgPhysics), ("task", TgTask), ("sensor", TgSensor),
        ("h_rest_q", C.POINTER(C.c_double)),
        ("h_tip_hull", C.POINTER(C.c_double)), ("n_tip_hull", C.c_int32), ("pad1", C.c_int32),
    ]


class TgHostStep(C.Structure):
    _fields_ = [
        ("h_actions", C.c_void_p), ("d_obs", C.c_void_p), ("d_reward", C.c_void_p), ("d_done", C.c_void_p), ("d_term_obs", C.c_void_p),
        ("d_feat", C.c_void_p), ("h_obs", C.c_void_p), ("h_reward", C.c_void_p), ("h_done", C.c_void_p), ("h_feat", C.c_void_p), ("h_oracle", C.c_void_p),
        ("chunks", C.c_int32),
        ("h_term_obs", C.c_void_p), ("h_term_idx", C.c_void_p), ("h_term_feat", C.c_void_p), ("term_cap", C.c_int32),
    ]


EXPORTS = [
    "tg_version", "tg_last_error", "tg_create", "tg_destroy", "tg_set_draws", "tg_set_rng_state", "tg_draws_poll", "tg_draws_upload", "tg_pipeline_error", "tg_pipeline_stalls", "tg_nan_resets", "tg_scan_fallbacks", "tg_scan_fallback_reasons", "tg_checkpoint_bytes", "tg_checkpoint_save", "tg_checkpoint_load", "tg_get_reset_counts", "tg_reset", "tg_step", "tg_step_host", "tg_bind_features", "tg_bind_oracle_obs",
    "tg_physics_only", "tg_raster_only", "tg_reset_only", "tg_state_size", "tg_get_state", "tg_set_state", "tg_get_camera",
    "tg_test_inverse_dynamics", "tg_test_mass_matrix", "tg_test_substep", "tg_test_substep_g8", "tg_launch_count",
]

_lib = None


class TgError(RuntimeError):
    pass


def load():
    """Load the CUDA library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise TgError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(tactile_gym_b200 has no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.tg_last_error.restype = C.c_char_p
    lib.tg_create.argtypes = [C.POINTER(TgConfig), C.c_int, C.POINTER(vp)]
    lib.tg_destroy.argtypes = [vp]
    lib.tg_set_draws.argtypes = [vp, vp, C.c_int]
    lib.tg_set_rng_state.argtypes = [vp, vp, vp]
    lib.tg_draws_poll.argtypes = [vp, vp, vp]
    lib.tg_draws_upload.argtypes = [vp, vp, vp, vp]
    lib.tg_pipeline_error.argtypes = [vp, vp]
    lib.tg_pipeline_stalls.argtypes = [vp, vp]
    lib.tg_nan_resets.argtypes = [vp, vp]
    lib.tg_scan_fallbacks.argtypes = [vp, vp]
    lib.tg_scan_fallback_reasons.argtypes = [vp, vp, vp]
    lib.tg_checkpoint_bytes.argtypes = [vp]
    lib.tg_checkpoint_bytes.restype = C.c_size_t
    lib.tg_checkpoint_save.argtypes = [vp, vp, C.c_size_t, vp]
    lib.tg_checkpoint_load.argtypes = [vp, vp, C.c_size_t, vp]
    lib.tg_get_reset_counts.argtypes = [vp, vp, vp]
    lib.tg_reset.argtypes = [vp, vp, vp, vp]
    lib.tg_step.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.tg_step_host.argtypes = [vp, C.POINTER(TgHostStep), vp]
    lib.tg_bind_features.argtypes = [vp, vp, vp]
    lib.tg_bind_oracle_obs.argtypes = [vp, vp, vp]
    lib.tg_physics_only.argtypes = [vp, vp, vp, vp, vp]
    lib.tg_raster_only.argtypes = [vp, vp, vp]
    lib.tg_reset_only.argtypes = [vp, vp, vp]
    lib.tg_state_size.argtypes = [vp]
    lib.tg_get_state.argtypes = [vp, vp, vp]
    lib.tg_set_state.argtypes = [vp, vp, vp]
    lib.tg_get_camera.argtypes = [vp, vp, vp]
    lib.tg_test_inverse_dynamics.argtypes = [vp, C.c_int, vp, vp, vp]
    lib.tg_test_mass_matrix.argtypes = [vp, C.c_int, vp, vp]
    lib.tg_test_substep.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
    lib.tg_test_substep_g8.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
    lib.tg_launch_count.argtypes = [vp]
    lib.tg_launch_count.restype = C.c_longlong
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise TgError("tactile_gym_b200 error %d: %s" % (rc, load().tg_last_error().decode()))
