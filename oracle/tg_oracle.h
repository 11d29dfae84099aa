/*
 * oracle/tg_oracle.h - CPU restatement (plain C, fp64) of the tactile_gym hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (tactile_gym_b200/) never does.
 *
 * PARITY STATUS
 *   raster + tactile post-process : pinned by the reference's .npy reference images
 *                                   (tests/test_oracle_raster.py).
 *   kinematics                    : pinned by the reference's rest_poses <-> workframe design
 *                                   identities (tests/test_oracle_kinematics.py).
 *   the reference's own Python    : pinned by vectors computed by RUNNING THE REFERENCE'S SOURCE (tools/make_reference_golden.py
 *   around the native calls         compiles its class bodies; tests/golden/reference_numpy.npz; tests/test_oracle_reference_golden.py):
 *                                   action encoding / scaling, work-frame transforms, tcp_velocity_control (UR5, MG400),
 *                                   tcp_position_control's IK target, blocking_move's retargeting / exit test, rewards and
 *                                   terminations of all tasks, surface lookups, push trajectories, sensor camera rig and
 *                                   t_s_camera, get_oracle_obs, the RNG call order of reset().
 *   dynamics (stepSimulation, IK) : PARITY UNPINNED.  The arithmetic lives in `pybullet`
 *                                   (requirements.txt:6, ">=3.1.0", not vendored, not installed).
 *                                   Restated from the published Bullet3 btMultiBody algorithm
 *                                   (Featherstone ABA in link-COM frames + projected Gauss-Seidel
 *                                   over multibody constraint rows); call sites anchored below.
 *
 * Every function cites the reference file:line (relative to /root/reference/tactile_gym/) it follows.
 */
#ifndef TG_ORACLE_H
#define TG_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define OR_MAXL 16 /* links  */
#define OR_MAXD 8  /* dofs   */

typedef struct {
    int nlinks, ndof;
    int parent[OR_MAXL];    /* bullet link index of parent, -1 = base */
    int jtype[OR_MAXL];     /* 0 fixed, 1 revolute */
    int dof_of_link[OR_MAXL]; /* -1 for fixed */
    int link_of_dof[OR_MAXD];
    double joint_xyz[OR_MAXL][3], joint_rpy[OR_MAXL][3], axis[OR_MAXL][3];
    double inertial_xyz[OR_MAXL][3], inertial_rpy[OR_MAXL][3];
    double mass[OR_MAXL], inertia[OR_MAXL][3];
    int tcp_link, body_link;
    /* physics parameters (base_tactile_env.py:125-130, base_robot_arm.py:24-25) */
    double gravity[3], dt;
    int solver_iters;
    double solver_residual_threshold; /* [EXT] pybullet default solverResidualThreshold = 1e-7 (not overridden at base_tactile_env.py:127-130) */
    double lin_damping, ang_damping, joint_damping;
    /* arm frames (base_robot_arm.py:39-46, set_TCP_lims :114-118) */
    double workframe_pos[3], workframe_rpy[3];
    double tcp_lims[6][2];
    double max_force, pos_gain, vel_gain;
    /* MG400 slaved joints (mg400.py:111-120): 1 = apply the parallelogram overwrite */
    int mg400_slave;
    /* camera (tactile_sensor.py:127-187) */
    double cam_pos[3], cam_rpy[3], fov_deg, focal_dist, near_, far_;
} OrModel;

/* joint motor state (the btMultiBodyJointMotor each joint owns) */
typedef struct {
    double q[OR_MAXD], qd[OR_MAXD];
    int motor_mode[OR_MAXD];   /* 0 = velocity, 1 = position */
    double target_pos[OR_MAXD], target_vel[OR_MAXD], kp[OR_MAXD], kd[OR_MAXD], max_force[OR_MAXD];
} OrState;

/* free rigid body (a pybullet floating-base multibody whose links are fixed to the base, e.g. the balance pole)
 * hanging on the arm by a point-to-point constraint (object_balance_env.py:261-283).  [EXT] conventions restated:
 * get/resetBasePositionAndOrientation act on the base link's inertial (COM) frame; velocities are world-frame;
 * createConstraint(JOINT_POINT2POINT) -> btMultiBodyPoint2Point: 3 rows along -x,-y,-z, Baumgarte term
 * erp (0.2) * gap / dt, |impulse| <= 500; applyExternalForce(WORLD_FRAME) lasts one stepSimulation. */
typedef struct {
    int enabled;
    double mass, inertia[3]; /* composite of base + fixed links about the composite COM, body axes */
    double com_off[3];       /* composite COM minus base-link COM, body frame */
    double pos[3], quat[4];  /* base-link COM pose (what getBasePositionAndOrientation returns) */
    double vel[3], omg[3];   /* world velocity of the base-link COM, world angular velocity */
    double ext_force[3], ext_pos[3]; /* pending applyExternalForce (world), consumed by the next step */
    int ext_pending;
    int p2p_enabled;
    double pivot_b[3];       /* constraint pivot in the base-link COM frame (childFramePosition) */
    double erp, max_impulse;
} OrObject;

/* Robot.step_sim() with the object in the world: gravity compensation + stepSimulation (motor rows + P2P rows) */
void or_step_sim_obj(const OrModel* m, OrState* s, OrObject* o);

/* ---- frame math (pybullet helpers used at base_robot_arm.py:47-118) ---- */
void or_quat_from_euler(const double rpy[3], double q[4]);
void or_euler_from_quat(const double q[4], double rpy[3]);
void or_mul_transforms(const double pa[3], const double qa[4], const double pb[3], const double qb[4], double po[3], double qo[4]);
void or_invert_transform(const double p[3], const double q[4], double po[3], double qo[4]);
void or_mat_from_quat(const double q[4], double R[9]);

/* ---- kinematics ---- */
/* getLinkState(...)[0:2] of every link: world pose of the link's INERTIAL frame (base_robot_arm.py:136-151) */
void or_link_states(const OrModel* m, const double* q, double pos[][3], double quat[][4]);
/* world pose of every URDF link frame (getLinkState(...)[4:6]); R row-major.  Visual meshes live in these frames. */
void or_link_frames(const OrModel* m, const double* q, double pos[][3], double R[][9]);
/* pb.calculateJacobian(link, localPosition=0) (base_robot_arm.py:300-307): rows 0-2 linear, 3-5 angular; [6][ndof] */
void or_jacobian(const OrModel* m, const double* q, int link, double J[6][OR_MAXD]);
/* pb.calculateInverseDynamics(q, qd, 0) (base_robot_arm.py:174-179) */
void or_inverse_dynamics(const OrModel* m, const double* q, const double* qd, const double* qdd, double* tau);
/* world-frame velocity of a link's inertial frame origin (getLinkState(...)[6:8]) */
void or_link_velocity(const OrModel* m, const double* q, const double* qd, int link, double lin[3], double ang[3]);

/* ---- dynamics: one pb.stepSimulation() with extra joint torques tau (robot.py:131-141) ---- */
void or_step_simulation(const OrModel* m, OrState* s, const double* tau_applied);
/* Robot.step_sim(): gravity compensation + stepSimulation (robot.py:131-141) */
void or_step_sim(const OrModel* m, OrState* s);
/* joint-space mass matrix via ABA unit responses (test hook) */
void or_mass_matrix_inverse(const OrModel* m, const double* q, double Minv[OR_MAXD][OR_MAXD]);
/* forward dynamics qdd = ABA(q, qd, tau) without damping (test hook: RNEA(ABA(tau)) == tau) */
void or_forward_dynamics(const OrModel* m, const double* q, const double* qd, const double* tau, int with_damping, double* qdd);

/* ---- control ---- */
/* BaseRobotArm.tcp_velocity_control (base_robot_arm.py:281-332, mg400.py:77-129) */
void or_tcp_velocity_control(const OrModel* m, OrState* s, const double vels_work[6]);
/* Robot.apply_action, TCP_velocity_control branch (robot.py:156-186) */
void or_apply_action(const OrModel* m, OrState* s, const double vels_work[6], int repeat);
/* BaseRobotArm.reset + tcp_direct_workframe_move + Robot.blocking_move (robot.py:114-125,188-260).
 * returns number of substeps used */
void or_blocking_retarget(int n, const double* q, const double* targ_j, double* cv, double* step_j);
int or_blocking_reached(const double tpos[3], const double targ_orn[4], const double tcp_pos[3], const double tcp_quat[4], const double* qd, int n);
int or_robot_reset(const OrModel* m, OrState* s, const double* rest_q, const double tcp_pos_work[3], const double tcp_rpy_work[3]);
/* pb.calculateInverseKinematics(..., maxNumIterations=100, residualThreshold=1e-8) (base_robot_arm.py:201-209) */
void or_tcp_position_target(const OrModel* m, const double* q, const double delta_work[6], double tpos[3], double targ_orn[4]);
int or_tcp_position_control(const OrModel* m, OrState* s, const double delta_work[6], int max_steps);
void or_inverse_kinematics(const OrModel* m, const double* q0, const double target_pos[3], const double target_quat[4], double* q_out);
void or_tcp_pose_workframe(const OrModel* m, const double* q, double pos[3], double rpy[3]);

/* ---- tactile raster (tactile_sensor.py:150-294) ---- */
/* camera pose from the body link: eye, and the 3 camera axes (forward, up, right) in world */
void or_camera_frame(const OrModel* m, const double* q, double eye[3], double fwd[3], double up[3], double right[3]);
/* window-space depth image of `ntri` world-space triangles composited over nodef_dep (float32), then the
 * t_s_camera post-process -> uint8 [S*S].  depth_out (float32 [S*S]) may be NULL. */
void or_tactile_image(const OrModel* m, const double* q, int S, const double* tris_world, int ntri,
                      const float* nodef_dep, const float* nodef_gray, const unsigned char* border_mask,
                      int border_on, unsigned char* img_out, float* depth_out);
/* depth image only (no nodef composite): background = 1.0 (far plane).  Used by the fixture KAT. */
void or_postprocess(int S, const float* cur, const float* nodef_dep, const float* nodef_gray, const unsigned char* border_mask,
                    int border_on, unsigned char* img_out);
void or_depth_image(const double eye[3], const double fwd[3], const double up[3], const double right[3],
                    double fov_deg, double near_, double far_, int S, const float* tris, int ntri, float* depth_out);

/* ---- surface_follow: OpenSimplex heightfield (rl_envs/exploration/surface_follow/base_surface_env.py:311-334,443-458) ----
 * PARITY UNPINNED: `opensimplex` (requirements.txt:4, unpinned, not vendored, not installed) - restated from the
 * published algorithm (K. Spencer's OpenSimplex, 2D, as shipped in the opensimplex 0.4 Python package): LCG-shuffled
 * permutation from the seed, stretch/squish constants (1/sqrt(3)-1)/2 and (sqrt(3)-1)/2, 8 gradients, norm 47. */
void or_opensimplex_init(long long seed, short perm[256]);
double or_opensimplex_noise2(const short perm[256], double x, double y);
/* gen_heigtfield_simplex_2d (:311-327): out[x*cols + y] = noise2(x*interp, y*interp) * range */
void or_surface_heights(long long seed, int rows, int cols, double interp, double range, double* out);


/* ---- object_push: contact rows (rl_envs/nonprehensile_manipulation/object_push/object_push_env.py) ----
 * PARITY UNPINNED, and more loosely restated than the motor-only path: bullet's narrow phase for this pair is
 * history dependent (GJK/EPA yields ONE point per frame, a persistent manifold caches up to four and drops the ones that
 * drift), so its exact contact set cannot be reproduced without bullet itself.  What is restated:
 *   geometry   memoryless manifolds rebuilt every stepSimulation from the poses at the start of the step:
 *                cube <-> table (table top z = 0, base_tactile_env.py:135-139 + table.urdf): the cube vertices with
 *                  z <= contactBreakingThreshold, normal +z, at most 4;
 *                tip core hull <-> cube (the convex hull of the tip link's collision mesh, t_s_core "fixed",
 *                  object_push_env.py:59): hull vertices inside the cube (per vertex: nearest face, signed distance),
 *                  reduced to <= 4 like a persistent manifold (the deepest, the two extremes along the first tangent axis
 *                  of the deepest point's face, the extreme along the second that is farther from the deepest);
 *                  cube-vertex-in-hull and edge-edge contacts are not generated;
 *   rows       [EXT] btMultiBodyConstraintSolver: per point one normal row (impulse >= 0; rhs = -rel_vel - dist/dt for
 *              dist > 0, -rel_vel - dist*erp/dt otherwise) and two friction rows (fixed basis btPlaneSpace1(normal); bullet's
 *              velocity-aligned first direction is deliberately not restated, see or_step_sim_push), implicit cone
 *              |f| <= mu * normal impulse (enableConeFriction=1,
 *              base_tactile_env.py:129); mu = product of the two lateralFriction values;
 *   materials  contactStiffness / contactDamping of the tip (sensors/tactile_sensor.py:314-332, values
 *              object_push_env.py:61-66) -> per-point erp = dt k / (dt k + d), cfm = 1 / (dt k + d) / dt [EXT];
 *              table contacts: erp 0.2, cfm 0 [EXT defaults];
 *   solver     motor rows (alternating order) then normal rows then friction pairs, <= 150 sweeps, residual exit as in
 *              or_step_simulation; no warm starting (memoryless);
 *   cube       floating base: gravity + [EXT] btMultiBody default damping 0.04 (1 + |v|) + gyroscopic term, explicit. */
#define OR_MAXC 8
typedef struct {
    double half[3];          /* cube half extents (cube.urdf: 0.08 box) */
    double table_z;          /* 0 */
    double mu_table, mu_tip; /* 0.065 * 1.0, 0.065 * 10 (object_push_env.py:216-225, :61-66, table.urdf) */
    double tip_k, tip_d;     /* combined contact stiffness / damping of the tip <-> cube pair */
    double erp;              /* 0.2 */
    double slop;             /* contactBreakingThreshold = 1e-4 (base_tactile_env.py:129) */
    double lin_damping, ang_damping; /* cube: btMultiBody defaults 0.04 / 0.04 [EXT] */
    int tip_link;            /* bullet link index owning the hull */
    int n_hull;
    const double* hull;      /* [n_hull][3], tip LINK frame */
    /* object_roll (rl_envs/nonprehensile_manipulation/object_roll/object_roll_env.py): the object is a SPHERE (sphere.urdf,
     * radius 0.0025 x globalScaling) between the table and the flat TacTip's core, a CYLINDER (ur5_with_flat_tactip.urdf:
     * length 0.0065, radius 0.02).  shape 1 selects: sphere <-> table = one point below the centre; sphere <-> cylinder =
     * one point on the cap facing the sphere while the centre projects inside the cap (rim / side contacts not generated) */
    int shape;               /* 0 cube vs hull (object_push), 1 sphere vs cylinder cap (object_roll) */
    double radius;           /* sphere radius this episode */
    double cyl_pos[3], cyl_axis[3], cyl_half_len, cyl_radius; /* cylinder centre / unit axis in the tip LINK frame */
    /* warm starting [EXT]: impulses (normal, friction 1, friction 2) the contact features ended the previous
     * stepSimulation with; part of the env state */
    double warmstart;        /* m_warmstartingFactor 0.85; 0 = off */
    int ws_n, ws_feature[OR_MAXC];
    double ws_impulse[OR_MAXC][3];
    /* diagnostics of the last substep */
    int n_contacts, n_iters;
    double normal_impulse[OR_MAXC];
    double contact_pos[OR_MAXC][3];
} OrPush;
/* Robot.step_sim() with the cube in the world: gravity compensation + stepSimulation (motor + contact rows) */
void or_step_sim_push(const OrModel* m, OrState* s, OrObject* cube, OrPush* p);
/* TCP_position_control's blocking move in the env's own world: o / p NULL = arm only, p NULL = arm + constrained object */
int or_tcp_position_control_world(const OrModel* m, OrState* s, OrObject* o, OrPush* p, const double delta_work[6], int max_steps);
/* object_roll tactile image: analytic sphere composited over nodef_dep, then the t_s_camera post-process */
void or_tactile_image_sphere(const OrModel* m, const double* q, int S, const double centre[3], double radius,
                             const float* nodef_dep, const float* nodef_gray, const unsigned char* border_mask,
                             int border_on, unsigned char* img_out);

#ifdef __cplusplus
}
#endif
#endif
