"""CPU known answers for the oracle's restatement of the modes beyond the BASELINE configs (no GPU): get_oracle_obs vectors,
sparse rewards, 1-d / flat surfaces, surface_follow-v2's action encoding, TCP_position_control.  The GPU parity tests compare
the CUDA path with these functions; here the functions themselves are checked against what the reference's code implies."""
import ctypes as C

import numpy as np


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_edge_oracle_obs_known_answers(oracle):
    """edge_follow_env.py:454-476: [tcp pos (3), tcp lin vel (3), goal pos (3), edge_ang], all in the work frame
    (rpy (-pi, 0, pi/2): x_work = world y, y_work = world x, z_work = -world z)"""
    e = oracle.EdgeFollowOracle(image_size=64)
    e.reset(draws=(0.003, 0.0))
    o = e.oracle_obs()
    assert o.shape == (10,)
    p, _ = oracle.tcp_pose_workframe(e.m, np.array(e.s.q[:6]))
    assert np.allclose(o[0:3], p, atol=1e-12) and abs(p[2] - 0.003) < 2e-4          # embedded 3 mm below the edge top
    assert np.allclose(o[6:9], [0.0, 0.175, 0.0], atol=1e-12)                        # goal at the +x end of the edge (edge_ang 0)
    assert o[9] == 0.0
    e.step(np.array([0.25, -0.125], np.float32))
    o = e.oracle_obs()
    assert np.allclose(o[3:6], [0.01, -0.005, 0.0], atol=2e-4)                      # the velocity motors hold the commanded twist
    e.reset(draws=(0.003, np.pi / 2))
    assert np.allclose(e.oracle_obs()[6:9], [0.175, 0.0, 0.0], atol=1e-12)


def test_object_oracle_obs_shapes_and_frames(oracle):
    b = oracle.ObjectBalanceOracle(image_size=64, seed=1); b.reset()
    o = b.oracle_obs()
    assert o.shape == (26,)
    assert abs(np.linalg.norm(o[3:7]) - 1) < 1e-9 and abs(np.linalg.norm(o[16:20]) - 1) < 1e-9      # two unit quaternions
    # object_balance's work frame is not flipped (rpy 0, object_balance_env.py:73-74): the pole's base COM sits
    # base_h / 2 - embed_dist above its origin (:214-219)
    assert abs(o[15] - (b.base_h / 2 - b.embed_dist)) < 1e-12 and abs(o[13]) < 1e-12 and abs(o[14]) < 1e-12
    p = oracle.ObjectPushOracle(image_size=64, seed=1); p.reset()
    o = p.oracle_obs()
    assert o.shape == (30,) and np.allclose(o[24:27], p.goal_pos_work) and np.allclose(o[27:30], p.goal_rpy_work)
    r = oracle.ObjectRollOracle(image_size=64, seed=1); r.reset()
    o = r.oracle_obs()
    assert o.shape == (34,) and np.allclose(o[26:29], r.goal_pos_tcp) and np.allclose(o[29:33], [0, 0, 0, 1]) and o[33] == r.radius


def test_sparse_rewards(oracle):
    # edge_follow_env.py:430-438: 1 inside termination_dist of the goal, else 0
    e = oracle.EdgeFollowOracle(image_size=64, reward_mode="sparse")
    e.reset(draws=(0.003, 0.0))
    assert e.step_data() == (0.0, False)
    pos, rpy = np.array([0.0, 0.17, 0.003]), np.zeros(3)                              # 5 mm from the goal (work frame y = world x)
    oracle.lib().or_robot_reset(C.byref(e.m), C.byref(e.s), _dptr(e.rest), _dptr(pos), _dptr(rpy))
    assert e.step_data() == (1.0, True)
    # object_balance_env.py:508-518: -1 once fallen, else 0
    b = oracle.ObjectBalanceOracle(image_size=64, seed=1); b.reward_mode = "sparse"; b.reset()
    assert b.step_data() == (0.0, False)
    q = oracle.quat_from_euler([0.7, 0.0, -np.pi / 2])                               # 40 degrees of roll
    for c in range(4):
        b.o.quat[c] = q[c]
    assert b.step_data() == (-1.0, True)


def test_surface_modes(oracle):
    # yz / yzRx: gen_heigtfield_simplex_1d (base_surface_env.py:339-357) and the (0, +-1) goal direction (:512-514)
    s = oracle.SurfaceFollowOracle(image_size=64, sensor="tactip", movement_mode="yzRx", reward_mode="sparse", render=False)
    s.reset(draws=(1234.0, -1.0))
    assert np.ptp(s.h, axis=0).max() == 0 and np.ptp(s.h) > 1e-3
    assert abs(s.h[5, 10] - oracle.opensimplex_noise2(1234, 1 * 0.05, 10 * 0.05) * 0.025) < 1e-15
    assert np.allclose(s.dirs, [0.0, -1.0, 0.0])
    assert abs(s.goal_pos[0] - (0.65 - 0.15)) < 1e-12 and abs(s.goal_pos[1]) < 1e-12    # y_work = world x
    # sparse: reset's own get_step_data already accumulates (:590-591, :638); nothing is paid away from the goal
    first = s.accum_rew
    assert first < 0 and s.reward == 0.0
    _, rew, done, _ = s.step(np.zeros(2, np.float32))
    assert rew == 0.0 and not done and s.accum_rew < first
    enc = s.encode_scale(np.array([0.1, -0.2], np.float32))                          # auto env: y driven, z / Rx the policy's
    assert abs(enc[1] + 0.01) < 1e-15 and enc[0] == 0 and abs(enc[2] - 0.004) < 1e-9 and enc[3] < 0 and enc[4] == 0
    # noise_mode "none": flat, no seed drawn
    f = oracle.SurfaceFollowOracle(image_size=64, sensor="digit", movement_mode="xyz", noise_mode="none", render=False, seed=3)
    a = oracle.gym_np_random(3)
    f.reset()
    assert np.all(f.h == 0) and abs(np.arctan2(f.dirs[1], f.dirs[0]) - a.uniform(-np.pi, np.pi)) < 1e-12
    # surface_follow-v2 (surface_follow_vert_env.py): flat for xRz, y driven, x the policy's, Rz without range, 10 / 3 weights
    v = oracle.SurfaceFollowOracle(image_size=64, sensor="digit", movement_mode="xRz", variant="vert", render=False)
    v.reset(draws=(99.0, 1.0))
    assert np.all(v.h == 0)
    enc = v.encode_scale(np.array([0.25, 0.25], np.float32))
    assert abs(enc[0] - 0.01) < 1e-15 and abs(enc[1] - 0.01 * 0.7) < 1e-15 and enc[5] == 0.0
    p, qt = v.tcp_world()
    # flat surface at z = 0.025; the reference point is the TCP moved embed_dist along the tip's own -z, which points up in the
    # world's z here (base_surface_env.py:727-757); tip level: no normal term
    assert abs(v.reward + 10.0 * abs((p[2] + v.embed_dist) - 0.025)) < 1e-6


def test_tcp_position_control(oracle):
    """base_robot_arm.py:228-279 + robot.py:188-260: 1 mm per full-scale step, the blocking move exits after a few substeps,
    targets are clipped to the TCP limits (check_TCP_pos_lims)"""
    e = oracle.EdgeFollowOracle(image_size=64, movement_mode="xyzRz", control_mode="TCP_position_control")
    e.reset(draws=(0.003, 0.3))
    p0 = e.oracle_obs()[:3].copy()
    for k in range(5):
        e.step(np.array([0.25, 0.0, 0.0, 0.0], np.float32))
        assert 1 <= e.last_move_substeps <= 10
    p1 = e.oracle_obs()[:3]
    assert abs((p1[0] - p0[0]) - 0.005) < 1e-4 and abs(p1[1] - p0[1]) < 1e-4 and abs(p1[2] - p0[2]) < 1e-4
    s = oracle.SurfaceFollowOracle(image_size=64, sensor="tactip", movement_mode="xyz", noise_mode="none", render=False,
                                   control_mode="TCP_position_control")
    s.reset(draws=(0.0, 0.0))
    for k in range(30):
        s.step(np.array([0.25], np.float32))                                          # z + 1 mm per step, limit + 25 mm
    z = s.oracle_obs()[2]
    assert 0.0245 < z < 0.0252


def test_vertical_surface_env(oracle):
    """surface_follow-v2 on its own (vertical) surface, oracle side: MG400 + forward TacTip facing an upright 1-d simplex
    heightfield 0.15 m above the table; the start pose embeds the tip 2.5 mm, the drive runs along y, the skin sees the surface"""
    e = oracle.SurfaceFollowOracle(image_size=128, arm="mg400", sensor="tactip", movement_mode="xRz", variant="vert",
                                   noise_mode="vertical_simplex", seed=5)
    o = e.reset()
    p, qt = e.tcp_world()
    ch = e.h[32, 32]
    assert abs(p[0] - (0.33 - ch + 0.0025)) < 3e-4 and abs(p[1]) < 3e-4 and abs(p[2] - 0.175) < 3e-4
    R = oracle.mat_from_quat(qt)
    assert np.allclose(R[:, 0], [1, 0, 0], atol=1e-3)                  # the forward sensor's axis points at the surface (+x)
    dep, gray, mask = e.ref
    assert (o[..., 0][mask == 0] > 0).sum() > 500                        # embedded: the surface shows on the skin
    y0 = p[1]
    for k in range(20):
        o, r, d, _ = e.step(np.array([0.0, 0.0], np.float32))
    p1, _ = e.tcp_world()
    # 1 mm per step along the drawn direction (work frame rpy (-pi, 0, 0): its y is the world's -y)
    assert abs(abs(p1[1] - y0) - 0.02) < 1e-3 and np.sign(p1[1] - y0) == -e.dirs[1]
    assert abs(p1[0] - p[0]) < 5e-4 and abs(p1[2] - p[2]) < 5e-4 and -1.0 < r < 0.0 and not d


def test_mg400_position_control(oracle):
    """MG400.tcp_position_control (mg400.py:131-190): the IK result's slaved joints follow j2_1 / j3_1; the 4-dof arm never meets
    blocking_move's orientation tolerance exactly, so every move runs its 10 substeps; 1 mm per full-scale step all the same"""
    e = oracle.EdgeFollowOracle(image_size=64, arm="mg400", sensor="digitac", movement_mode="xyzRz", control_mode="TCP_position_control", seed=1)
    e.reset()
    p0 = e.oracle_obs()[:3].copy()
    for k in range(5):
        e.step(np.array([0.25, 0.0, 0.0, 0.0], np.float32))
        assert e.last_move_substeps == 10
        q = np.array(e.s.q[:8])
        assert abs(q[5] - q[1]) + abs(q[6] + q[1]) + abs(q[7] - (q[1] + q[2])) < 5e-6
    d = e.oracle_obs()[:3] - p0
    assert abs(d[0] - 0.005) < 1e-4 and abs(d[1]) < 1e-4 and abs(d[2]) < 1e-4


def test_position_control_with_the_object_in_the_world(oracle):
    """TCP_position_control on the object tasks (or_tcp_position_control_world): the blocking move steps the env's whole world -
    the pole stays on its constraint while the tip is moved 1 mm per full-scale step, the cube is pushed, the marble is rolled
    along; every move takes 1 .. 10 substeps"""
    b = oracle.ObjectBalanceOracle(image_size=64, movement_mode="xy", control_mode="TCP_position_control", rand_gravity=False, rand_embed_dist=False, seed=1)
    b.reset()
    p0, _ = b.tcp_world() if hasattr(b, "tcp_world") else (None, None)
    tcp0 = b.oracle_obs()[:3].copy()
    for k in range(10):
        o, r, d, _ = b.step(np.array([0.25, 0.0], np.float32))
        assert 1 <= b.last_move_substeps <= 10 and r == 1.0 and not d
    tcp1 = b.oracle_obs()[:3]
    assert abs((tcp1[0] - tcp0[0]) - 0.010) < 3e-4 and abs(tcp1[1] - tcp0[1]) < 3e-4            # 10 x 1 mm along work-frame x
    piv = np.array(b.o.pos[:])                                                                    # the pole followed the tip: < 1 mm behind
    assert abs((piv[0] - b.init_obj_pos[0]) - 0.010) < 1.5e-3 and abs(piv[2] - b.init_obj_pos[2]) < 1e-3

    p = oracle.ObjectPushOracle(image_size=64, arm="ur5", sensor="tactip", movement_mode="xyRz", traj_type="straight", control_mode="TCP_position_control", seed=2)
    p.reset()
    c0 = np.array(p.o.pos[:])
    touched = False
    for k in range(40):
        o, r, d, _ = p.step(np.array([0.25, 0.0, 0.0], np.float32))
        assert 1 <= p.last_move_substeps <= 10 and not d
        touched = touched or p.p.n_contacts > 4
    c1 = np.array(p.o.pos[:])
    assert touched and 0.005 < np.linalg.norm(c1[:2] - c0[:2]) < 0.06 and abs(c1[2] - c0[2]) < 1e-4    # pushed about the 40 mm the tip moved (it coasts on mu = 0.065)

    rl = oracle.ObjectRollOracle(image_size=64, control_mode="TCP_position_control", seed=3)
    rl.reset(draws=np.array([1.0, 0.0028, 0.0, 0.0, 0.0, 0.014]))      # embedded deep enough for the tip core to hold the marble (as in test_gpu_roll.py)
    m0 = np.array(rl.o.pos[:])
    t0 = rl.tcp_world()[0].copy()
    for k in range(10):
        rl.step(np.array([0.25, 0.0], np.float32))
        assert 1 <= rl.last_move_substeps <= 10
    m1 = np.array(rl.o.pos[:]); t1 = rl.tcp_world()[0]
    dt_, dm = np.linalg.norm(t1[:2] - t0[:2]), np.linalg.norm(m1[:2] - m0[:2])
    # the tip moved its 10 x 1 mm; the marble went along (position control moves the plate in jerks - 3 substeps per move - so the
    # marble is not in steady rolling and need not sit at half the plate's travel as under velocity control)
    assert 0.008 < dt_ < 0.0105 and 0.002 < dm < 0.03 and np.dot(m1[:2] - m0[:2], t1[:2] - t0[:2]) > 0
