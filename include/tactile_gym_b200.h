/*
 * tactile_gym_b200.h - C ABI of the batched tactile-RL environment engine (libtactile_gym_b200.so).
 *
 * The reference (ac-93/tactile_gym) has no FFI of its own: its seam to native code is the pybullet
 * C-extension, called ~100 times per env step through BulletClient.  This ABI sits one
 * level up and replaces, for N envs at once, exactly the calls on the per-step hot path:
 *
 *   tg_step      <- BaseTactileEnv.step                      rl_envs/base_tactile_env.py:166-185
 *                     encode/scale actions                   rl_envs/exploration/edge_follow/edge_follow_env.py:345-369,
 *                                                            rl_envs/base_tactile_env.py:141-164
 *                     Robot.apply_action                     robots/arms/robot.py:156-186
 *                       tcp_velocity_control                 robots/arms/base_robot_arm.py:281-332 (pb.calculateJacobian :300)
 *                       24 x Robot.step_sim                  robots/arms/robot.py:131-141
 *                         pb.calculateInverseDynamics        robots/arms/base_robot_arm.py:174-179
 *                         pb.stepSimulation                  robots/arms/robot.py:141
 *                     get_step_data / reward / termination   rl_envs/exploration/edge_follow/edge_follow_env.py:371-452
 *                     TactileSensor.get_imgs + t_s_camera    sensors/tactile_sensor.py:212-294 (pb.getCameraImage :239)
 *   tg_reset     <- EdgeFollowEnv.reset / Robot.reset        edge_follow_env.py:311-336, robots/arms/robot.py:114-125,188-260
 *                     pb.calculateInverseKinematics          robots/arms/base_robot_arm.py:201-209
 *   tg_set_draws <- self.np_random.uniform(...) draws        edge_follow_env.py:293-297,240 (host generates them with the
 *                                                            gym RandomState so seeds reproduce; the device consumes them)
 *
 * Conventions
 *   - plain C: pointers + sizes only, no torch types.  `d_*` arguments are DEVICE pointers owned by the
 *     caller (e.g. torch.cuda tensors' data_ptr()); `h_*` are host pointers.
 *   - every call returns 0 on success or a negative TG_E* code, never throws / exits; tg_last_error()
 *     gives the thread-local message.
 *   - all work is enqueued on the caller's stream (cudaStream_t passed as void*); no hidden syncs in
 *     tg_step / tg_reset.  A TgWorld is not thread safe.  One world per device; one rank per GPU.
 *   - there is no CPU fallback: without a CUDA device tg_create fails with TG_ENODEV.
 */
#ifndef TACTILE_GYM_B200_H
#define TACTILE_GYM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TG_VERSION 100 /* 0.1.0 */

#define TG_OK 0
#define TG_EINVAL (-1)
#define TG_ENODEV (-2)
#define TG_ECUDA (-3)
#define TG_ENOMEM (-4)
#define TG_EUNSUPPORTED (-5)

#define TG_MAXB 8     /* moving bodies == dofs of the arm */
#define TG_MAXSUB 16  /* original URDF links that carry mass (per-link damping terms) */
#define TG_MAXTRI 64  /* stimulus triangles per env */
#define TG_MAXDRAW 8  /* random draws consumed per reset */

/* arm topologies the kernels are specialised for */
#define TG_TOPO_CHAIN6 0 /* UR5: 6 revolute joints in a serial chain               */
#define TG_TOPO_MG400 1  /* MG400: 8 revolute joints, tree 0-1-2-3-4 + 0-5-6-7       */

/* tasks */
#define TG_TASK_EDGE_FOLLOW 0
#define TG_TASK_OBJECT_BALANCE 1
#define TG_TASK_SURFACE_FOLLOW 2
#define TG_TASK_OBJECT_PUSH 3
#define TG_TASK_OBJECT_ROLL 4
#define TG_SURF_N 64 /* heightfield rows = columns */
#define TG_PUSH_NTRAJ 10 /* object_push: goals along the trajectory (object_push_env.py:233) */
#define TG_ORACLE_NOBS 36 /* observation_mode "oracle": floats per env (10 edge, 20 surface, 26 balance, 30 push, 34 roll; rest 0) */
#define TG_PUSH_NFEAT 12 /* object_push: extended_feature length (object_push_env.py:611-629) */

/* object_push action encodings (object_push_env.py:369-454) */
#define TG_PUSH_WORK 0       /* xyRz: plain scatter of the policy action                                  */
#define TG_PUSH_WORK_DRIVE 1 /* y, yRz: x slot driven with max_action                                     */
#define TG_PUSH_TCP_TYRZ 2   /* TyRz: drive along the tip axis, action[0] across it, both in the TCP frame */
#define TG_PUSH_TCP_TXTYRZ 3 /* TxTyRz                                                                    */

/* Reduced arm model: fixed joints are merged into their moving parent at asset-compile time
 * (tactile_gym_b200/scene.py); `sub_*` keeps the original mass-carrying links for bullet's per-link
 * velocity damping.  All vectors are in the owning body's frame (= URDF frame of its first link). */
typedef struct {
    int32_t topo, nb, nsub, pad0;
    double jpos[TG_MAXB][3];   /* joint origin in the parent body's frame          */
    double jrot[TG_MAXB][9];   /* parent body frame -> this body frame at q = 0     */
    double axis[TG_MAXB][3];   /* unit joint axis in this body's frame              */
    double mass[TG_MAXB];
    double com[TG_MAXB][3];
    double inertia[TG_MAXB][6]; /* about the COM, body frame: xx xy xz yy yz zz       */
    int32_t sub_start[TG_MAXB + 1]; /* sub-links of body b are sub_start[b] .. sub_start[b+1]-1 */
    int32_t pad1;
    int32_t sub_body[TG_MAXSUB];
    double sub_mass[TG_MAXSUB];
    double sub_com[TG_MAXSUB][3];
    double sub_rot[TG_MAXSUB][9]; /* body frame -> that link's inertial frame           */
    double sub_inertia[TG_MAXSUB][3];
    int32_t tcp_body, cam_body;
    double tcp_pos[3], tcp_rot[9]; /* TCP link's INERTIAL frame in tcp_body's frame (getLinkState()[0:2]) */
    double cam_pos[3], cam_rot[9]; /* camera frame (body link inertial frame o cam_pos/cam_rpy)           */
} TgArm;

typedef struct {
    double gravity[3];
    double dt;              /* 1/240 */
    int32_t solver_iters;   /* 150   */
    int32_t substeps;       /* 24    */
    double lin_damping, ang_damping, joint_damping; /* 0.04 0.04 0.01 */
    double max_force, pos_gain, vel_gain;           /* 1000 1 1       */
    double solver_residual_threshold; /* 1e-7: pybullet's default solverResidualThreshold [EXT] */
    double blocking_force;  /* 1e5: pybullet's default when setJointMotorControlArray gets no `forces` */
    int32_t gravity_comp;   /* 1: Robot.step_sim's calculateInverseDynamics torque is applied */
    int32_t pad0;
} TgPhysics;

typedef struct {
    int32_t task, act_dim, max_steps, n_draws;
    int32_t act_index[6];      /* policy action k -> slot of the 6-vector (x y z Rx Ry Rz), -1 unused */
    double act_min, act_max;   /* +-0.25 */
    double act_lo[6], act_hi[6]; /* per-slot affine range (m/s, rad/s) */
    double workframe_pos[3], workframe_rpy[3];
    double tcp_lims[6][2];
    /* edge_follow */
    double edge_pos[3], edge_len, edge_height, termination_dist;
    double embed_lo, embed_hi; /* uniform range; lo == hi -> fixed */
    double init_rpy[3];
    double draw_default[TG_MAXDRAW];
    /* object_balance (object_balance_env.py): a free rigid body hanging on the TCP by a point-to-point constraint.
     * draws per reset: gravity_z, embed_dist, fx, fy (signed fractions of the half base width where 0.1 N pushes down) */
    double obj_mass, obj_inertia[3]; /* composite about the composite COM, body axes */
    double obj_com_off[3];           /* composite COM minus base-link COM, body frame */
    double obj_base_com[3];          /* base-link COM in the base LINK frame (the stimulus frame) */
    double obj_init_rpy[3];          /* init_obj_rpy = (0, 0, -pi/2) */
    double obj_base_w, obj_base_h;   /* 0.1, 0.0025 */
    double obj_force;                /* 0.1 N */
    double obj_term_deg, obj_term_pos; /* 35 deg, 0.1 m */
    double p2p_erp, p2p_max_impulse; /* 0.2, 500 [EXT] */
    /* surface_follow (base_surface_env.py, surface_follow_auto_env.py): a 64 x 64 OpenSimplex heightfield, new every
     * episode.  draws per reset: OpenSimplex seed (randint(1e8), as a double), goal direction angle */
    double surf_pos[3];              /* surface body position (also the x/y centre of the grid) */
    double surf_grid, surf_range;    /* 0.006 m, 0.025 m */
    double surf_interp, surf_extent; /* 0.05 noise zoom, 0.15 m goal distance / TCP limits */
    double surf_embed;               /* embed_dist: 0.0025 tactip, 0.0015 digit / digitac */
    double surf_drive;               /* constant drive along the goal direction: max_action x {1, 0.9, 0.7} */
    double surf_w_norm;              /* weight of the normal-alignment term (0 for yz / xyz movement) */
    double surf_w_goal, surf_w_surf; /* weights of the xy goal distance / surface distance: 0, 1 surface_follow-v0
                                        (surface_follow_auto_env.py:76-94); 1, 10 surface_follow-v1 (surface_follow_goal_env.py:62-81) */
    /* object_push (object_push_env.py, base_object_env.py): a free cube on the table pushed by the tip core; contact rows
     * tip hull <-> cube and cube <-> table.  draws per reset: init_obj_ang, obj_mass, OpenSimplex seed | trajectory angle */
    int32_t push_mode, push_traj_straight, push_sparse_reward;
    int32_t push_shape;              /* 0: cube vs tip hull (object_push); 1: sphere vs cylinder cap (object_roll) */
    /* reward_mode "sparse" of edge_follow (1 at the goal else 0, edge_follow_env.py:430-438), object_balance (-1 once fallen else
     * 0, object_balance_env.py:508-518) and surface_follow (the dense reward accumulated over the episode, reset included, paid
     * out at the goal, surface_follow_auto_env.py:59-73 / surface_follow_goal_env.py:53-67); object_push / object_roll keep
     * push_sparse_reward */
    int32_t sparse_reward;
    int32_t surf_mode;               /* heights: 0 simplex 2-d (xyz, xyzRxRy), 1 simplex 1-d along y (yz, yzRx; :339-357), 2 flat (noise_mode "none"), 3 simplex 1-d along the rows (vertical), 4 uniform noise on 2 x 2 blocks (noise_mode "random", :302-318; device RNG only) */
    int32_t surf_dir_mode;           /* goal direction: 0 (cos, sin) of the drawn angle; 1 (0, +-1) = the drawn choice([-1, 1]) (:512-514) */
    int32_t surf_drive_y_only;       /* 1: surface_follow-v2 drives along y only, x stays the policy's (surface_follow_vert_env.py:30-45) */
    /* control_mode (robots/arms/robot.py:156-186): 0 TCP_velocity_control (Jacobian inverse, velocity motors, `substeps` steps);
     * 1 TCP_position_control (base_robot_arm.py:228-279: the scaled action is a pose DELTA in the work frame, clipped to
     * tcp_lims, IK from the current joints, position motors, then blocking_move(max_steps = pos_max_steps, robot.py:188-260):
     * step until the pose error / joint speed test passes).  Built for the motor-only tasks (edge_follow, surface_follow), UR5. */
    int32_t control_mode, pos_max_steps;
    /* surface_follow-v2's own surface (noise_mode "vertical_simplex", base_surface_env.py:60-63,83-107,248-259): the heightfield
     * body stands upright, turned by euler (0, -pi/2, 0) about surf_pos (a local point (x, y, z) sits at surf_pos + (-z, y, x)),
     * the `forward` sensor type faces it; heights vary along the rows only (surf_mode 3, gen_heigtfield_simplex_1d_vertical
     * :359-379); surface distance and tip axis are taken along x (:708-754) */
    int32_t surf_vertical, pad_vertical;
    double push_half[3];             /* cube half extents (cube.urdf) */
    double push_table_z;             /* table top (base_tactile_env.py:135-139 + table.urdf) */
    double push_mu_table, push_mu_tip; /* products of the lateralFriction pairs (object_push_env.py:218, :61-66, table.urdf) */
    double push_tip_k, push_tip_d;   /* combined contactStiffness / contactDamping of the tip <-> cube pair */
    double push_erp, push_slop;      /* 0.2 [EXT]; contactBreakingThreshold 1e-4 (base_tactile_env.py:129) */
    double push_lin_damping, push_ang_damping; /* cube: btMultiBody defaults 0.04 [EXT] */
    double push_init_pos[3];         /* init_obj_pos (:193) */
    double push_inertia_per_mass[3]; /* box inertia / mass: changeDynamics(mass=) keeps the shape's inertia [EXT] */
    double push_term_dist;           /* 0.025 (:68) */
    double push_traj_spacing, push_traj_perturb, push_traj_offset; /* 0.025, 0.1, obj_width / 2 + spacing (:233-235, :290) */
    /* object_roll (object_roll_env.py): a marble (sphere.urdf, radius roll_radius x the episode's scaling factor) between the
     * table and the flat TacTip, whose core is a cylinder (ur5_with_flat_tactip.urdf:320-325).  It shares object_push's
     * contact solve (push_* materials, push_init_pos = marble x / y, obj_mass) with push_shape = 1.  The workframe height
     * changes per episode (2 r - embed_dist, :197-202); the goal is fixed in the TCP frame (:244-286).
     * draws per reset: scaling_factor, embed_dist, dx, dy, goal_ang, goal_dist */
    double roll_radius;              /* default_obj_radius 0.0025 (:166) */
    double roll_cyl_pos[3], roll_cyl_axis[3]; /* cylinder centre / unit axis in the frame of arm.tcp_body */
    double roll_cyl_half_len, roll_cyl_radius; /* 0.00325, 0.02 */
    /* The reset draws as a PROGRAM over the env's own random stream (device RNG, tg_set_rng_state; csrc/tg_rng.cuh): draw d is
     * produced by draw_kind[d] - 0 constant draw_default[d] (consumes nothing), 1 uniform(draw_lo, draw_hi), 2 randint(draw_hi),
     * 3 choice([-1, 1]), 4 choice([-1, 1]) * rand() - in index order, which is the reference's call order for every task.
     * surf_mode 4 (noise_mode "random", base_surface_env.py:290-309) additionally draws 1,024 uniform(0, 0.2 surf_range) heights
     * between draw 0 and draw 1, as the reference does (update_surface before make_goal, :539-547). */
    int32_t draw_kind[TG_MAXDRAW];
    double draw_lo[TG_MAXDRAW], draw_hi[TG_MAXDRAW];
} TgTask;

typedef struct {
    int32_t image_size;  /* S */
    int32_t border_on;
    double fov_deg, near_, far_;
    const float* h_nodef_dep;      /* [S*S] */
    const float* h_nodef_gray;     /* [S*S] */
    const uint8_t* h_border_mask;  /* [S*S] */
    int32_t n_prim, pad0;
    const double* h_prims;         /* [n_prim][4][3] stimulus primitives in the stimulus frame: convex planar polygons
                                      with 3 or 4 vertices (coplanar triangle pairs merged; unused 4th vertex ignored) */
    const int32_t* h_prim_nv;      /* [n_prim] vertex counts */
    /* optional: the primitives grouped into CONVEX parts (closed convex polyhedra: the edge box, the cube, the pole's plate and
     * post).  With it the scanline raster (csrc/tg_raster_scan.cuh) renders the stimulus; NULL / 0: the general kernel does. */
    const int32_t* h_prim_part;    /* [n_prim] part index of each primitive */
    const double* h_part_centroid; /* [n_parts][3] an interior point of each part, stimulus frame */
    int32_t n_parts, pad1;
} TgSensor;

typedef struct {
    int32_t n_envs;
    int32_t lanes_per_warp; /* physics kernels: 0 = auto; 1/2/4/8/16/32 = one thread per env with that many active lanes per
                               warp; -8 = one env per group of 8 lanes, one body per lane (csrc/tg_g8.cuh: edge_follow /
                               surface_follow under TCP_velocity_control; what auto picks for them up to 6,144 envs) */
    TgArm arm;
    TgPhysics phys;
    TgTask task;
    TgSensor sensor;
    const double* h_rest_q; /* [nb] */
    /* contact envs: vertices of the tip core's convex hull, in the frame of arm.tcp_body (the body that carries the
     * sensor); NULL / 0 otherwise */
    const double* h_tip_hull; /* [n_tip_hull][3] */
    int32_t n_tip_hull, pad1;
} TgConfig;

typedef struct TgWorld TgWorld;

int tg_version(void);
const char* tg_last_error(void);

int tg_create(const TgConfig* cfg, int device, TgWorld** out);
int tg_destroy(TgWorld* w);

/* Start a sequence of random draws for future resets: h_draws[n_envs][rounds][n_draws] (double).  The r-th reset of env i after
 * this call consumes h_draws[i][r] (r < rounds).  Synchronous host->device copy; pre-computed next episodes are recomputed. */
int tg_set_draws(TgWorld* w, const double* h_draws, int rounds);
/* Keeping the sequence fed without synchronising.  The device buffer is a RING: the r-th reset of env i reads slot r % rounds
 * and needs r < avail[i] (tg_set_draws: avail = rounds).
 *   tg_draws_poll:   enqueue on `stream` the device->host copy of resets consumed per env [N] followed by the error flags [1]
 *                    into h_counts (page-locked, N + 1 ints); the caller waits for it with its own event.
 *   tg_draws_upload: enqueue the host->device copy of the whole ring h_ring[N][rounds][n_draws] and of h_avail[N] (both
 *                    page-locked, untouched until the copy has run).  The caller may only have changed slots whose draws were
 *                    consumed (r < counts[i]); every other slot must hold what the device already has. */
/* The alternative to streaming draws from the host: every env carries the reference's own generator (numpy RandomState =
 * MT19937) on the device and the resets pull from it with numpy's call semantics (TgTask.draw_kind).  h_key[n_envs][624] +
 * h_pos[n_envs] are the generators' states (numpy's get_state() after gym's seeding; pos 624 = freshly seeded).  Starts a new
 * sequence like tg_set_draws (synchronous; pre-computed next episodes are recomputed) and stays in force until tg_set_draws. */
int tg_set_rng_state(TgWorld* w, const uint32_t* h_key, const int32_t* h_pos);
int tg_draws_poll(TgWorld* w, int32_t* h_counts, void* stream);
int tg_draws_upload(TgWorld* w, const double* h_ring, const int32_t* h_avail, void* stream);
/* Sticky error flags (0: none).  bit 0: reset pipeline / heightfield raster overflow; bit 1: a reset found no draw left in the
 * ring (it used TgTask.draw_default): the host fell behind with tg_draws_upload; bit 2: an env step ended in a non-finite
 * state (see tg_nan_resets).  Synchronises. */
int tg_pipeline_error(TgWorld* w, void* stream);
/* Per-env NaN / inf guard: env steps so far whose state came out non-finite.  Such a step reports done = 1,
 * reward = 0 and - under auto-reset - the env starts its next episode like after any other episode end.  Synchronises. */
int tg_nan_resets(TgWorld* w, void* stream);
/* Profiling counter: envs of the LAST raster pass that the scanline raster (convex stimuli) handed to the general raster kernel
 * (a face plane seen edge-on, a vertex behind the eye or outside [near, far]).  Synchronises. */
int tg_scan_fallbacks(TgWorld* w, void* stream);
/* ... and why: h_counts[8] = envs of the last raster pass per reason code (0 kept, 1 near-plane cut, 2 eye inside a part,
 * 3 more front faces than the table holds, 4 test hook).  Synchronises. */
int tg_scan_fallback_reasons(TgWorld* w, int32_t* h_counts, void* stream);

/* Full checkpoint / resume: every device buffer of the world - live state, pre-computed next episodes and
 * partial rebuilds, both heightfields, per-env RNG states, the draw ring, counters.  After tg_checkpoint_load into a world
 * created from the same TgConfig the following steps reproduce the original run bit for bit.  (tg_get_state / tg_set_state
 * carry the live episode only: test and debugging hooks.)  The caller's output buffers are its own to save. */
size_t tg_checkpoint_bytes(const TgWorld* w);
int tg_checkpoint_save(TgWorld* w, void* h_blob, size_t bytes, void* stream);
int tg_checkpoint_load(TgWorld* w, const void* h_blob, size_t bytes, void* stream);
/* Number of episode ends so far that found their pre-computed next episode unfinished and completed it inline
 * (exact either way; a performance counter: episodes shorter than the rebuild - about 40 step launches at max_steps >= 200,
 * where the quanta are smallest, one launch for very short episodes).  Synchronises. */
int tg_pipeline_stalls(TgWorld* w, void* stream);
/* resets consumed per env since the last tg_set_draws (device->host, synchronises the stream) */
int tg_get_reset_counts(TgWorld* w, int32_t* h_counts, void* stream);

/* Reset envs with d_mask[i] != 0 (NULL: all) and render their first observation into d_obs[i]. */
int tg_reset(TgWorld* w, const uint8_t* d_mask, uint8_t* d_obs, void* stream);

/* One env step for all envs.  d_actions [N][act_dim] f32, d_obs [N][S][S][1] u8, d_reward [N] f32,
 * d_done [N] u8.  Envs that finish are reset in place and d_obs holds their first observation of the new
 * episode (VecEnv semantics); if d_term_obs != NULL their terminal observation is kept in d_term_obs[i]. */
int tg_step(TgWorld* w, const float* d_actions, uint8_t* d_obs, float* d_reward, uint8_t* d_done,
            uint8_t* d_term_obs, void* stream);

/* One env step with HOST buffers, the call behind the numpy VecEnv (replaces SubprocVecEnv.step_async/step_wait's pipes,
 * stable_baselines3 via tactile_gym/sb3_helpers/rl_utils.py:17-35, around BaseTactileEnv.step base_tactile_env.py:166-185).
 * h_* must be page-locked host memory for the copies to overlap; d_* are the caller-owned device tensors tg_step takes
 * (they hold the same results afterwards).  The observation is rendered and copied out in `chunks` env ranges: the
 * device->host copy of one range runs (on an internal copy stream) while the next range and the terminal observations
 * are rendered.  Everything is ordered into `stream`: synchronising it makes the host buffers valid.  d_term_obs / d_feat /
 * h_feat may be NULL (d_feat is the buffer bound with tg_bind_features). */
#define TG_HOST_MAX_CHUNKS 16
typedef struct TgHostStep {
    const float* h_actions; /* [N][act_dim] */
    uint8_t* d_obs;         /* [N][S][S][1]; NULL together with h_obs: no image is rendered (observation_mode "oracle") */
    float* d_reward;        /* [N] */
    uint8_t* d_done;        /* [N] */
    uint8_t* d_term_obs;    /* [N][S][S][1] or NULL */
    const float* d_feat;    /* [N][TG_PUSH_NFEAT] or NULL */
    uint8_t* h_obs;
    float* h_reward;
    uint8_t* h_done;
    float* h_feat;          /* or NULL */
    float* h_oracle;        /* [N][TG_ORACLE_NOBS] or NULL: the buffer bound with tg_bind_oracle_obs, copied out */
    int chunks;             /* <= 0: 4 */
    /* Terminal observations of the envs that finished in this step, compacted on the device (ascending env index) and copied
     * out with everything else, so that the caller needs no second round trip: h_term_idx[0] = how many finished (all of them,
     * also beyond term_cap); h_term_idx[1 + j] / h_term_obs[j] / h_term_feat[j] = env index, terminal image [S][S][1] and
     * terminal feature row [TG_PUSH_NFEAT] of the j-th, for j < min(count, term_cap).  term_cap rows are copied every step
     * whatever the count (the copy is enqueued before the count is known): keep it near the number of episode ends a step
     * really sees.  term_cap 0 or h_term_idx NULL: off.  Needs d_term_obs.  h_term_feat may be NULL. */
    uint8_t* h_term_obs;    /* [term_cap][S][S][1] pinned, or NULL */
    int32_t* h_term_idx;    /* [1 + term_cap] */
    float* h_term_feat;     /* [term_cap][TG_PUSH_NFEAT] or NULL */
    int term_cap;
} TgHostStep;
int tg_step_host(TgWorld* w, const TgHostStep* hs, void* stream);

/* object_push / object_roll / surface_follow-v1, observation_mode "tactile_and_feature": bind caller-owned device buffers [N][TG_PUSH_NFEAT] f32
 * that every following tg_step / tg_reset / tg_physics_only fills with the extended_feature (object_push_env.py:611-629: TCP
 * pos(3) + rpy(3) in the work frame, goal pos(3) + rpy(3) in the work frame; object_roll_env.py:402-408: the goal position in
 * the TCP frame in the first 3 entries; surface_follow_goal_env.py:83-97: TCP pos(3) + goal pos(3) in the work frame).  d_term_feat (may be NULL) receives the features of the
 * state a finished env terminated in.  NULL d_feat unbinds. */
int tg_bind_features(TgWorld* w, float* d_feat, float* d_term_feat);

/* observation_mode "oracle" (get_oracle_obs: edge_follow_env.py:454-476, base_surface_env.py:789-819, object_balance_env.py:528-563,
 * object_push_env.py:571-609, object_roll_env.py:371-409; the "oracle" entry of get_observation, base_tactile_env.py:258-262):
 * bind caller-owned device buffers [N][TG_ORACLE_NOBS] f32 that every following tg_step / tg_step_host / tg_reset /
 * tg_physics_only fills with the state vector of each env (the task's first k entries are meaningful).  d_term_oracle (may be
 * NULL) receives the vector of the state a finished env terminated in.  NULL d_oracle unbinds. */
int tg_bind_oracle_obs(TgWorld* w, float* d_oracle, float* d_term_oracle);

/* kernel-level entry points (tests, ncu) */
int tg_physics_only(TgWorld* w, const float* d_actions, float* d_reward, uint8_t* d_done, void* stream);
int tg_raster_only(TgWorld* w, uint8_t* d_obs, void* stream);
int tg_reset_only(TgWorld* w, const uint8_t* d_mask, void* stream);

/* state access (parity tests, checkpointing).  Layout: per env doubles
 * [q(nb) qd(nb) tcp_pos(3) tcp_quat(4) embed edge_ang steps reset_substeps | obj pos(3) quat(4) vel(3) omg(3) scalar | goal]
 * scalar = the episode's gravity_z (object_balance), cube mass (object_push) or marble radius (object_roll);
 * goal = object_push's trajectory index */
int tg_state_size(const TgWorld* w); /* doubles per env */
int tg_get_state(TgWorld* w, double* h_state, void* stream);
int tg_set_state(TgWorld* w, const double* h_state, void* stream);
/* camera of every env as the raster sees it: [N][12] doubles eye, fwd, up, right */
int tg_get_camera(TgWorld* w, double* h_cam, void* stream);

/* test hooks on the dynamics kernels: inputs/outputs are HOST arrays, n rows */
int tg_test_inverse_dynamics(TgWorld* w, int n, const double* h_q, const double* h_qd, double* h_tau);
int tg_test_mass_matrix(TgWorld* w, int n, const double* h_q, double* h_M /* [n][nb][nb] */);
int tg_test_substep(TgWorld* w, int n, int nsteps, double* h_q, double* h_qd, const double* h_target_vel);
/* the same through the 8-lanes-per-env substep (csrc/tg_g8.cuh) */
int tg_test_substep_g8(TgWorld* w, int n, int nsteps, double* h_q, double* h_qd, const double* h_target_vel);

/* number of kernels launched by this library since creation (bench.py's gpu_launches) */
long long tg_launch_count(const TgWorld* w);

#ifdef __cplusplus
}
#endif
#endif
