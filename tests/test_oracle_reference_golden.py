"""CPU: the oracle against golden vectors computed by the REFERENCE'S OWN SOURCE (tests/golden/reference_numpy.npz, written by
tools/make_reference_golden.py, which compiles the reference's class bodies from /root/reference and runs their pure-numpy
methods): action encoding + scaling of every env / movement mode / control mode (object_push's TCP-frame encodings included),
the work-frame transforms, the edge reward and termination geometry, the surface index lookup, distances and the three surface
envs' dense rewards, object_push's trajectory of goals (around a stand-in noise function), its rewards, goal advancing and
extended feature, object_balance's fall test, object_roll's TCP-frame goal, and get_oracle_obs of all five tasks.  This pins the
parts of the oracle that restate reference Python (as opposed to pybullet's C++) to the reference itself."""
import os

import numpy as np
import pytest

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_numpy.npz"))
TOL = 1e-12


def _ang_close(a, b, tol=1e-9):
    return np.allclose(np.sin(a), np.sin(b), atol=tol) and np.allclose(np.cos(a), np.cos(b), atol=tol)


def _check_actions(key, env):
    acts, want = GOLD["act_%s_in" % key], GOLD["act_%s_out" % key]
    got = np.array([env.encode_scale(a.astype(np.float64)) for a in acts])
    assert got.shape == want.shape and np.allclose(got, want, rtol=0, atol=1e-15), (key, np.abs(got - want).max())


@pytest.mark.parametrize("mode", ["xy", "xyz", "xyRz", "xyzRz"])
@pytest.mark.parametrize("cm", ["TCP_velocity_control", "TCP_position_control"])
def test_edge_actions(oracle, mode, cm):
    _check_actions("edge_%s_%s" % (mode, cm[4:7]), oracle.EdgeFollowOracle(image_size=64, movement_mode=mode, control_mode=cm))


@pytest.mark.parametrize("sensor", ["tactip", "digitac", "digit"])
def test_surface_actions(oracle, sensor):
    for mode in ("yz", "xyz", "yzRx", "xyzRxRy"):
        for cm in ("TCP_velocity_control", "TCP_position_control"):
            e = oracle.SurfaceFollowOracle(image_size=64, sensor=sensor, movement_mode=mode, control_mode=cm, render=False)
            e.dirs = GOLD["surface_dirs"]
            _check_actions("surfauto_%s_%s_%s" % (sensor, mode, cm[4:7]), e)
    v = oracle.SurfaceFollowOracle(image_size=64, sensor=sensor, movement_mode="xRz", variant="vert", render=False)
    v.dirs = np.array([0.0, -1.0, 0.0])
    _check_actions("surfvert_%s" % sensor, v)
    if sensor == "tactip":
        for mode in ("yz", "xyz", "yzRx", "xyzRxRy"):
            _check_actions("surfgoal_%s" % mode, oracle.SurfaceFollowOracle(image_size=64, sensor=sensor, movement_mode=mode, variant="goal", render=False))


def test_object_actions(oracle):
    for mode in ("xy", "xyz", "RxRy", "xyRxRy"):
        _check_actions("balance_%s" % mode, oracle.ObjectBalanceOracle(image_size=64, movement_mode=mode))
    _check_actions("roll_xy", oracle.ObjectRollOracle(image_size=64))


@pytest.mark.parametrize("key", ["flipped", "upright"])
def test_workframe_transforms(oracle, key):
    g = lambda n: GOLD["frame_%s_%s" % (key, n)]
    m = oracle.load_model("ur5", "tactip", "standard", g("wpos"), g("wrpy"), g("lims"))
    wq = oracle.quat_from_euler(g("wrpy"))
    for k in range(len(g("pos"))):
        p, r = oracle.world_to_work(m, g("pos")[k], oracle.quat_from_euler(g("rpy")[k]))        # worldframe_to_workframe
        assert np.allclose(p, g("w2k_pos")[k], atol=TOL) and _ang_close(r, g("w2k_rpy")[k])
        # workframe_to_worldframe (base_robot_arm.py:47-60) from the primitives the oracle's reset / position control use
        po, qo = oracle.mul_transforms(g("wpos"), wq, g("pos")[k] - g("wpos"), oracle.quat_from_euler(g("rpy")[k]))
        assert np.allclose(po, g("k2w_pos")[k], atol=TOL) and _ang_close(oracle.euler_from_quat(qo), g("k2w_rpy")[k])
        v = g("vec")[k]
        assert np.allclose(oracle.world_to_work_vec(m, v[:3]), g("vec_w2k")[k], atol=TOL)
        assert np.allclose(np.concatenate([oracle.world_to_work_vec(m, v[:3]), oracle.world_to_work_vec(m, v[3:])]), g("vel_w2k")[k], atol=TOL)
        R = oracle.mat_from_quat(wq)
        assert np.allclose(R @ v[:3], g("vec_k2w")[k], atol=TOL)


def test_edge_reward_geometry(oracle):
    rows, angs, tcps, steps = GOLD["edge_rows"], GOLD["edge_ang"], GOLD["edge_tcp"], GOLD["edge_steps"]
    e = oracle.EdgeFollowOracle(image_size=64, max_steps=250)
    e.reset(draws=(0.003, 0.0))
    c, s_ = np.cos, np.sin
    for k in range(len(rows)):
        ang = angs[k]
        # update_edge (edge_follow_env.py:237-283) as EdgeFollowOracle.reset restates it, without the arm move
        e.edge_ang = ang
        e.goal_pos = np.array([e.edge_pos[0] + e.edge_len * c(ang), e.edge_pos[1] + e.edge_len * s_(ang), e.edge_pos[2] + e.edge_height])
        e.edge_end_points = np.array([[e.edge_pos[0] - e.edge_len * c(ang), e.edge_pos[1] - e.edge_len * s_(ang), e.edge_pos[2] + e.edge_height],
                                      [e.edge_pos[0] + e.edge_len * c(ang), e.edge_pos[1] + e.edge_len * s_(ang), e.edge_pos[2] + e.edge_height]])
        assert np.allclose(e.goal_pos, rows[k, 8:11], atol=TOL)
        gw, _ = oracle.world_to_work(e.m, e.goal_pos, np.array([0.0, 0.0, 0.0, 1.0]))
        assert np.allclose(gw, rows[k, 5:8], atol=TOL)
        e.tcp_world = lambda k=k: (tcps[k], np.array([0.0, 0.0, 0.0, 1.0]))
        e.steps = int(steps[k])
        e.reward_mode = "dense"
        rew, done = e.step_data()
        assert abs(rew - rows[k, 2]) < TOL and done == bool(rows[k, 4])
        assert abs(-rew - (rows[k, 0] + 10.0 * rows[k, 1])) < TOL            # = W_goal * goal_dist + W_edge * edge_dist
        e.reward_mode = "sparse"
        assert e.step_data()[0] == rows[k, 3]
    assert rows[:, 3].max() == 1.0 and rows[:, 4].min() == 0.0 and rows[:, 4].max() == 1.0     # the vectors cover both outcomes


def test_surface_lookup_and_rewards(oracle, monkeypatch):
    h = GOLD["surf_h"]
    monkeypatch.setattr(oracle, "surface_heights", lambda *a, **k: h.copy())
    envs = []
    for variant, mode in (("auto", "xyzRxRy"), ("auto", "xyz"), ("goal", "xyzRxRy"), ("vert", "xRz")):
        e = oracle.SurfaceFollowOracle(image_size=64, sensor="tactip", movement_mode=mode, variant=variant, render=False, max_steps=200)
        e.reset(draws=(1.0, 0.3 if variant != "vert" else 1.0))
        envs.append(e)
    a = envs[0]
    assert np.array_equal(a.h, h)
    assert np.allclose(a.x_bins, GOLD["surf_x_bins"], atol=0) and np.allclose(a.y_bins, GOLD["surf_y_bins"], atol=0)
    for p, ij in zip(GOLD["surf_pts"], GOLD["surf_idx"]):
        assert a.xy_to_surface_idx(p[0], p[1]) == (int(ij[0]), int(ij[1])), p
    v = envs[3]                                   # -v2 keeps its surface flat for "xRz"; the reference vectors use h for all four
    v.h, v.surface_array, v.surface_normals = a.h, a.surface_array, a.surface_normals
    rows = GOLD["surf_rows"]
    for k, (p, r) in enumerate(zip(GOLD["surf_tcp_pos"], GOLD["surf_tcp_rpy"])):
        q = oracle.quat_from_euler(r)
        for c, e in enumerate(envs):
            e.goal_pos = GOLD["surf_goal"]
            e.tcp_world = lambda p=p, q=q: (p, q)
            e.steps = 10
            rew, done = e.step_data()
            z_dist, cos_dist, dense = rows[k, 3 * c: 3 * c + 3]
            assert abs(rew - dense) < 1e-12, (k, c, rew, dense)
            assert done == bool(rows[k, 14])
        # the weights, spelled out: auto = -(z + cos), auto/xyz = -z, goal = -(xy + 10 z + cos), vert = -(10 z + 3 cos)
        z, cs, xy = rows[k, 0], rows[k, 1], rows[k, 13]
        assert abs(rows[k, 2] + (z + cs)) < 1e-12 and abs(rows[k, 5] + z) < 1e-12
        assert abs(rows[k, 8] + (xy + 10 * z + cs)) < 1e-12 and abs(rows[k, 11] + (10 * z + 3 * cs)) < 1e-12


# ---------------------------------------------------------------- object tasks and the oracle observations
def _fake_noise(seed, x, y):       # the stand-in noise function of tools/make_reference_golden.py (section E)
    return 0.6 * np.sin(2.3 * x + 0.4) * np.cos(0.7 * y) + 0.2 * np.sin(5.1 * x)


def _inject_tcp(oracle, monkeypatch, env, tcp):
    """make the oracle env see a given world TCP state [pos, rpy, lin vel, ang vel] instead of its joints'"""
    pos, quat, lin, ang = tcp[0:3], oracle.quat_from_euler(tcp[3:6]), tcp[6:9], tcp[9:12]
    env.tcp_world = lambda: (pos, quat)

    def state(m, s):
        p, r = oracle.world_to_work(m, pos, quat)
        return p, r, oracle.quat_from_euler(r), oracle.world_to_work_vec(m, lin), oracle.world_to_work_vec(m, ang)

    monkeypatch.setattr(oracle, "tcp_state_workframe", state)


def _set_obj(env, o13):
    for c in range(3):
        env.o.pos[c] = o13[c]
    for c in range(4):
        env.o.quat[c] = o13[3 + c]
    if len(o13) >= 13:
        for c in range(3):
            env.o.vel[c] = o13[7 + c]; env.o.omg[c] = o13[10 + c]


def test_push_actions_trajectory_rewards_and_observations(oracle, monkeypatch):
    monkeypatch.setattr(oracle, "opensimplex_noise2", _fake_noise)
    assert np.allclose(GOLD["push_wpos"], oracle.ObjectPushOracle(image_size=64, arm="ur5", sensor="tactip").workframe_pos)
    rpy = GOLD["push_tcp_rpy_for_actions"]
    for mode in ("y", "yRz", "xyRz", "TyRz", "TxTyRz"):
        e = oracle.ObjectPushOracle(image_size=64, arm="ur5", sensor="tactip", movement_mode=mode)
        e.tcp_world = lambda: (GOLD["push_wpos"], oracle.quat_from_euler(rpy))
        _check_actions("push_%s" % mode, e)
    for traj_type, third in (("simplex", 4242.0), ("straight", 0.23)):
        g = lambda n: GOLD["push_%s_%s" % (traj_type, n)]
        e = oracle.ObjectPushOracle(image_size=64, arm="ur5", sensor="tactip", movement_mode="TyRz", traj_type=traj_type, max_steps=1000)
        e.reset(draws=np.array([0.0, 0.491, third]))
        assert np.allclose(e.traj_pos_work, g("traj_pos_work"), atol=TOL) and np.allclose(e.traj_rpy_work, g("traj_rpy_work"), atol=TOL)
        assert np.allclose(e.traj_pos_world, g("traj_pos_world"), atol=TOL)
        assert np.allclose(np.abs(np.sum(e.traj_orn_world * g("traj_orn_world"), axis=1)), 1.0, atol=1e-12)     # same rotation (q ~ -q)
        tcp = np.concatenate([GOLD["push_wpos"] + np.array([0.01, 0.02, 0.0]), rpy, [0.004, -0.003, 0.001], [0.01, 0.02, -0.2]])
        _inject_tcp(oracle, monkeypatch, e, tcp)
        e.targ = -1
        e.update_goal()
        e.steps = 5
        rows = g("rows")
        for k, row in enumerate(rows):
            _set_obj(e, g("obj")[k])
            e.reward_mode = "sparse"
            targ = e.targ
            assert e.step_data()[0] == row[1], (traj_type, k)
            e.targ = targ                                   # (step_data advances the goal: rewind for the dense evaluation)
            e.goal_pos_world, e.goal_orn_world = e.traj_pos_world[min(targ, 9)], e.traj_orn_world[min(targ, 9)]
            e.goal_pos_work, e.goal_rpy_work = e.traj_pos_work[min(targ, 9)], e.traj_rpy_work[min(targ, 9)]
            e.reward_mode = "dense"
            rew, done = e.step_data()
            assert abs(rew - row[0]) < 1e-12 and abs(rew - row[2]) < 1e-12 and done == bool(row[3]), (traj_type, k, rew, row[:4])
            assert e.targ == int(row[4]), (traj_type, k, e.targ, row[4])
            assert np.allclose(e.features(), row[5:17], atol=1e-12), (traj_type, k)
        assert rows[-1][3] == 1.0 and rows[-1][4] == 10            # walked the whole trajectory: done when the last goal is reached
        if traj_type == "simplex":
            e.targ = int(GOLD["push_oracle_goal_index"][0]) - 1
            e.update_goal()
            _set_obj(e, GOLD["push_oracle_obj"])
            assert np.allclose(e.oracle_obs(), GOLD["push_oracle_obs"], atol=1e-12)


def test_balance_fall_rewards_and_observation(oracle, monkeypatch):
    b = oracle.ObjectBalanceOracle(image_size=64, rand_gravity=False, rand_embed_dist=False)
    b.reset(draws=np.array([-0.1, 0.0035, 0.0, 0.0]))
    assert np.allclose(b.init_obj_pos, GOLD["balance_init_pos"], atol=1e-15)
    for o7, row in zip(GOLD["balance_obj"], GOLD["balance_rows"]):
        _set_obj(b, o7)
        b.steps = int(row[4])
        b.reward_mode = "dense"
        rew, done = b.step_data()
        assert rew == row[3] and done == bool(row[1])
        b.reward_mode = "sparse"
        assert b.step_data()[0] == row[2]
    assert 0 < GOLD["balance_rows"][:, 0].sum() < len(GOLD["balance_rows"])      # both fallen and standing poses are covered
    _inject_tcp(oracle, monkeypatch, b, GOLD["balance_oracle_tcp"])
    _set_obj(b, GOLD["balance_oracle_obj"])
    got, want = b.oracle_obs(), GOLD["balance_oracle_obs"]
    assert got.shape == want.shape == (26,) and np.allclose(got, want, atol=1e-12), np.abs(got - want).max()


def test_roll_goal_rewards_and_observation(oracle, monkeypatch):
    radius, embed = GOLD["roll_radius"]
    r = oracle.ObjectRollOracle(image_size=64, rand_obj_size=True, rand_embed_dist=True, rand_init_obj_pos=True, max_steps=250)
    r.reset(draws=np.array([radius / 0.0025, embed, 0.0, 0.0, 0.0, 0.01]))
    assert np.allclose(r.workframe_pos, GOLD["roll_wpos"], atol=1e-15)
    _inject_tcp(oracle, monkeypatch, r, np.concatenate([GOLD["roll_tcp"], [0.004, -0.003, 0.0], [0.0, 0.0, 0.0]]))
    for row in GOLD["roll_rows"]:
        r.goal_pos_tcp = row[0:3].copy()
        _set_obj(r, np.concatenate([row[3:6], [0.0, 0.0, 0.0, 1.0]]))
        r.steps = int(row[12])
        r.reward_mode = "dense"
        rew, done = r.step_data()
        assert np.allclose(r.goal_pos_world, row[6:9], atol=1e-12)
        assert abs(rew - row[9]) < 1e-12 and done == bool(row[11])
        r.reward_mode = "sparse"
        assert r.step_data()[0] == row[10]
        assert np.allclose(r.features(), row[13:16], atol=0)
    assert GOLD["roll_rows"][:, 10].max() == 1.0 and GOLD["roll_rows"][:, 10].min() == 0.0
    _set_obj(r, GOLD["roll_oracle_obj"])
    got, want = r.oracle_obs(), GOLD["roll_oracle_obs"]
    assert got.shape == want.shape == (34,) and np.allclose(got, want, atol=1e-12), np.abs(got - want).max()


def test_edge_and_surface_oracle_observations(oracle, monkeypatch):
    e = oracle.EdgeFollowOracle(image_size=64)
    e.reset(draws=(0.003, 1.1))
    _inject_tcp(oracle, monkeypatch, e, GOLD["edge_oracle_tcp"])
    got, want = e.oracle_obs(), GOLD["edge_oracle_obs"]
    assert got.shape == want.shape == (10,) and np.allclose(got, want, atol=1e-12), np.abs(got - want).max()
    h = GOLD["surf_h"]
    monkeypatch.setattr(oracle, "surface_heights", lambda *a, **k: h.copy())
    s = oracle.SurfaceFollowOracle(image_size=64, sensor="tactip", render=False)
    s.reset(draws=(1.0, 0.3))
    s.goal_pos = GOLD["surf_goal"]
    _inject_tcp(oracle, monkeypatch, s, GOLD["surf_oracle_tcp"])
    s.step_data()
    got, want = s.oracle_obs(), GOLD["surf_oracle_obs"]
    assert got.shape == want.shape == (20,) and np.allclose(got, want, atol=1e-12), np.abs(got - want).max()


SENSOR_CASES = [("tactip", "standard", 64, False), ("tactip", "standard", 128, False), ("tactip", "flat", 128, False),
                ("digit", "standard", 128, False), ("digitac", "right_angle", 128, True), ("tactip", "standard", 256, False)]


@pytest.mark.parametrize("name,typ,S,border_off", SENSOR_CASES)
def test_sensor_camera_rig_and_postprocess(oracle, name, typ, S, border_off):
    """TactileSensor.setup_camera_info / update_cam_frame / get_imgs / t_s_camera (sensors/tactile_sensor.py:127-294) run from
    the reference's source: the camera pose relative to the sensor body, and the depth -> uint8 arithmetic on a synthetic depth"""
    key = "sensor_%s_%s_%d" % (name, typ, S)
    cam = GOLD[key + "_cam"]
    eye0, target0, up0, (fov, focal, near, far) = cam[0:3], cam[3:6], cam[6:9], cam[9:13]
    m = oracle.load_model("ur5", name, typ, [0.65, 0.0, 0.035], [-np.pi, 0.0, np.pi / 2], np.zeros((6, 2)))
    assert (m.fov_deg, m.near_, m.far_) == (fov, near, far) and abs(m.focal_dist - focal) < 1e-15
    fwd0 = (target0 - eye0) / np.linalg.norm(target0 - eye0)
    assert abs(np.linalg.norm(target0 - eye0) - focal) < 1e-12 and abs(np.dot(fwd0, up0)) < 1e-12
    rng = np.random.RandomState(S)
    for k in range(3):
        q = rng.uniform(-1.0, 1.0, m.ndof)
        P, Q = oracle.link_states(m, q)           # getLinkState()[0:2]: the body link's INERTIAL frame (tactile_sensor.py:153-155)
        Pb, Rb = P[m.body_link], oracle.mat_from_quat(Q[m.body_link])
        e, f, u, r = oracle.camera_frame(m, q)
        assert np.allclose(e, Pb + Rb @ eye0, atol=1e-12) and np.allclose(f, Rb @ fwd0, atol=1e-12) and np.allclose(u, Rb @ up0, atol=1e-12)
        assert np.allclose(r, np.cross(f, u), atol=1e-12)
    # t_s_camera on the synthetic depth, with OUR copy of the fixture images: equal bytes wherever the segmentation mask did not
    # flag the sensor body; flagged pixels are zeroed by the reference and then take the border grey where the border is on
    ref = oracle.load_refimg(name, typ, S)
    got = oracle.postprocess(GOLD[key + "_cur"], ref, border_on=not border_off)
    want, body = GOLD[key + "_img"], GOLD[key + "_seg_body"]
    assert np.array_equal(got[~body], want[~body])
    dep, gray, mask = ref
    expect_body = np.where((mask == 1) & (not border_off), gray.astype(np.uint8), 0)
    assert np.array_equal(want[body], expect_body[body])
    free = (~body) & ((mask == 0) | border_off)
    assert (want[free] == 255).sum() > 10 and (want[free] == 0).sum() > 10 and len(np.unique(want[free])) > 50     # clip, dead band, ramp


@pytest.mark.parametrize("arm,sensor", [("ur5", "tactip"), ("mg400", "digitac")])
def test_tcp_velocity_control_targets(oracle, arm, sensor):
    """tcp_velocity_control (base_robot_arm.py:281-332; mg400.py:77-129) run from the reference source on stored kinematic inputs:
    limit handling, work -> world twist, Jacobian stacking, inverse / pseudo-inverse, MG400 slaving -> joint velocity targets"""
    import ctypes as C

    g = lambda n: GOLD["velctl_%s_%s" % (arm, n)]
    wpos = [0.33, 0.0, 0.035] if arm == "mg400" else [0.65, 0.0, 0.035]
    m = oracle.load_model(arm, sensor, "standard", wpos, [-np.pi, 0.0, np.pi / 2], g("lims"))
    n = m.ndof
    capped = 0
    for q, v, J, pose, want in zip(g("q"), g("v"), g("J"), g("pose"), g("target")):
        P, Q = oracle.link_states(m, q)                     # the stored inputs still are the oracle's kinematics
        assert np.allclose(np.concatenate([P[m.tcp_link], Q[m.tcp_link]]), pose, atol=1e-12)
        assert np.allclose(oracle.jacobian(m, q, m.tcp_link), J, atol=1e-12)
        s = oracle.OrState()
        for i in range(n):
            s.q[i] = q[i]
        vv = np.ascontiguousarray(v, dtype=np.float64)
        oracle.lib().or_tcp_velocity_control(C.byref(m), C.byref(s), vv.ctypes.data_as(C.POINTER(C.c_double)))
        got = np.array(s.target_vel[:n])
        assert np.allclose(got, want, rtol=1e-9, atol=1e-12), (np.abs(got - want).max(), got, want)
        assert all(s.motor_mode[i] == 0 and s.max_force[i] == 1000.0 and s.kd[i] == 1.0 for i in range(n))
        p, r = oracle.tcp_pose_workframe(m, q)
        cur = np.concatenate([p, r])
        capped += int(np.any(((cur < g("lims")[:, 0]) & (v < 0)) | ((cur > g("lims")[:, 1]) & (v > 0))))
    assert capped >= 2          # the vectors do exercise check_TCP_vel_lims


def test_tcp_position_control_target(oracle):
    """tcp_position_control (base_robot_arm.py:228-279) run from the reference source: the pose handed to the IK"""
    import ctypes as C

    m = oracle.load_model("ur5", "tactip", "standard", [0.65, 0.0, 0.035], [-np.pi, 0.0, np.pi / 2], GOLD["posctl_lims"])
    clipped = 0
    for q, d, want in zip(GOLD["posctl_q"], GOLD["posctl_delta"], GOLD["posctl_ik_target"]):
        tpos, torn = np.zeros(3), np.zeros(4)
        dp = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(C.POINTER(C.c_double))
        oracle.lib().or_tcp_position_target(C.byref(m), dp(q), dp(d), tpos.ctypes.data_as(C.POINTER(C.c_double)), torn.ctypes.data_as(C.POINTER(C.c_double)))
        assert np.allclose(tpos, want[:3], atol=1e-12)
        assert abs(abs(np.dot(torn, want[3:7])) - 1.0) < 1e-12           # the same rotation (q ~ -q)
        p, r = oracle.tcp_pose_workframe(m, q)
        t = np.concatenate([p, r]) + d
        clipped += int(np.any((t < GOLD["posctl_lims"][:, 0]) | (t > GOLD["posctl_lims"][:, 1])))
    assert clipped >= 4         # check_TCP_pos_lims is exercised


# ---------------------------------------------------------------- the RNG call order of reset()
def _draw_rows(draw_fn, seed, rounds=3):
    from tactile_gym_b200 import seeding

    return draw_fn(seeding.np_random(int(seed))[0], rounds)


def test_reset_draw_order_matches_the_reference(oracle):
    """Each env's reset() run from the reference source with a recording RandomState (tools/make_reference_golden.py, section L)
    against the engine's host-side draw functions and the oracle envs' own draws: same stream, same order, same ranges."""
    from tactile_gym_b200 import engine as E

    base = {"control_mode": "TCP_velocity_control", "observation_mode": "tactile", "reward_mode": "dense", "arm_type": "ur5"}
    g = lambda k: (GOLD["draws_" + k], int(GOLD["draws_%s_seed" % k][0]))
    # edge_follow: [embed_dist,] edge_ang
    for sensor in ("tactip", "digit", "digitac"):
        want, seed = g("edge_" + sensor)
        cfg, _ = E.edge_follow_config(dict(base, movement_mode="xy", noise_mode="rand_height", tactile_sensor_name=sensor), [64, 64], 200, 1)
        assert np.allclose(_draw_rows(E.edge_follow_draws(cfg.task), seed), want, rtol=0, atol=1e-15)
        o = oracle.EdgeFollowOracle(image_size=64, sensor=sensor, seed=seed)
        assert np.allclose([o.draw() for _ in range(3)], want, rtol=0, atol=1e-15)
    want, seed = g("edge_fixed")
    cfg, _ = E.edge_follow_config(dict(base, movement_mode="xy", noise_mode="fixed_height", tactile_sensor_name="tactip"), [64, 64], 200, 1)
    rows = _draw_rows(E.edge_follow_draws(cfg.task), seed)
    assert np.allclose(rows[:, 1:], want, rtol=0, atol=1e-15) and np.all(rows[:, 0] == 0.0035)
    # surface_follow: [OpenSimplex seed,] goal angle | +-1
    for key, fn, mode, noise in (("surface_xyzRxRy", E.surface_follow_config, "xyzRxRy", "simplex"), ("surface_yzRx", E.surface_follow_config, "yzRx", "simplex"),
                                 ("surface_xyz_none", E.surface_follow_config, "xyz", "none"), ("surface_vert_xRz", E.surface_follow_vert_config, "xRz", "simplex")):
        want, seed = g(key)
        _, _, draw = fn(dict(base, movement_mode=mode, noise_mode=noise, tactile_sensor_name="tactip"), [64, 64], 200, 1)
        rows = _draw_rows(draw, seed)
        assert np.allclose(rows if noise == "simplex" else rows[:, 1:], want, rtol=0, atol=1e-15), key
        variant = "vert" if "vert" in key else "auto"
        o = oracle.SurfaceFollowOracle(image_size=64, sensor="tactip", movement_mode=mode, variant=variant, noise_mode=noise, seed=seed, render=False)
        got = np.array([o.draw() for _ in range(3)])
        assert np.allclose(got if noise == "simplex" else got[:, 1:], want, rtol=0, atol=1e-15), key
    # object_balance: [gravity] [embed] choice rand choice rand -> gravity, embed, fx, fy
    for key, rg, re_ in (("balance_rand", True, True), ("balance_fixed", False, False), ("balance_gravity_only", True, False)):
        want, seed = g(key)
        _, _, draw = E.object_balance_config(dict(base, movement_mode="xy", object_mode="pole", rand_gravity=rg, rand_embed_dist=re_, tactile_sensor_name="tactip"),
                                             [64, 64], 250, 1)
        rows = _draw_rows(draw, seed)
        k = 0
        if rg:
            assert np.allclose(rows[:, 0], want[:, k], rtol=0, atol=1e-15); k += 1
        else:
            assert np.all(rows[:, 0] == -0.1)
        if re_:
            assert np.allclose(rows[:, 1], want[:, k], rtol=0, atol=1e-15); k += 1
        assert np.allclose(rows[:, 2], want[:, k] * want[:, k + 1], rtol=0, atol=1e-15) and np.allclose(rows[:, 3], want[:, k + 2] * want[:, k + 3], rtol=0, atol=1e-15)
        o = oracle.ObjectBalanceOracle(image_size=64, rand_gravity=rg, rand_embed_dist=re_, seed=seed)
        got = np.array([oracle.balance_draws(o.np_random, "tactip", rg, re_) for _ in range(3)])
        assert np.allclose(got[:, 2], rows[:, 2], rtol=0, atol=1e-15) and np.allclose(got[:, 0], rows[:, 0], rtol=0, atol=1e-15)
    # object_push: [init angle] [mass] OpenSimplex seed | trajectory angle
    for key, ro, rm, tt in (("push_rand_simplex", True, True, "simplex"), ("push_fixed_simplex", False, False, "simplex"), ("push_rand_straight", True, True, "straight")):
        want, seed = g(key)
        _, _, draw = E.object_push_config(dict(base, movement_mode="TyRz", rand_init_orn=ro, rand_obj_mass=rm, traj_type=tt, tactile_sensor_name="tactip"),
                                          [64, 64], 1000, 1)
        rows = _draw_rows(draw, seed)
        assert np.allclose(rows if ro else rows[:, 2:], want, rtol=0, atol=1e-15), key
        o = oracle.ObjectPushOracle(image_size=64, arm="ur5", sensor="tactip", rand_init_orn=ro, rand_obj_mass=rm, traj_type=tt, seed=seed)
        got = np.array([oracle.push_draws(o.np_random, ro, rm, tt) for _ in range(3)])
        assert np.allclose(got, rows, rtol=0, atol=1e-15), key
    # object_roll: [scale] [embed] [dx dy] goal angle, goal distance
    for key, a_, b_, c_ in (("roll_rand", True, True, True), ("roll_fixed", False, False, False), ("roll_pos_only", False, False, True)):
        want, seed = g(key)
        _, _, draw = E.object_roll_config(dict(base, movement_mode="xy", rand_obj_size=a_, rand_embed_dist=b_, rand_init_obj_pos=c_, tactile_sensor_name="tactip"),
                                          [64, 64], 250, 1)
        rows = _draw_rows(draw, seed)
        cols = ([0] if a_ else []) + ([1] if b_ else []) + ([2, 3] if c_ else []) + [4, 5]
        assert np.allclose(rows[:, cols], want, rtol=0, atol=1e-15), key
        o = oracle.ObjectRollOracle(image_size=64, rand_obj_size=a_, rand_embed_dist=b_, rand_init_obj_pos=c_, seed=seed)
        got = np.array([oracle.roll_draws(o.np_random, a_, b_, c_) for _ in range(3)])
        assert np.allclose(got, rows, rtol=0, atol=1e-15), key


def test_blocking_move_retargeting_and_exit(oracle):
    """Robot.blocking_move (robot.py:188-260) run from the reference source on an ideal position servo: the commanded joint
    target of every iteration (constant-velocity retargeting with the halving rule) and the iteration the move ends on (the
    test reads the pose and joint speeds from BEFORE the step) against the oracle's or_blocking_retarget / or_blocking_reached,
    the two functions its reset and its position control are built from"""
    import ctypes as C

    dp = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(C.POINTER(C.c_double))
    L = oracle.lib()
    L.or_blocking_reached.restype = C.c_int
    m = oracle.load_model("ur5", "tactip", "standard", [0.65, 0.0, 0.035], [-np.pi, 0.0, np.pi / 2], np.zeros((6, 2)))
    for case in (0, 1):
        hist, t = GOLD["blocking_%d_hist" % case], GOLD["blocking_%d_target" % case]
        targ_j, tpos, torn, cv0, max_steps = t[0:6], t[6:9], t[9:13], t[13], int(t[14])
        cv = C.c_double(cv0)
        assert len(hist) < max_steps                      # ended on the exit test, not on the step budget
        for k, row in enumerate(hist):
            q, qd, cmd = row[0:6].copy(), row[6:12].copy(), row[12:18]
            if cv0 > 0:
                step = np.zeros(6)
                L.or_blocking_retarget(C.c_int(6), dp(q), dp(targ_j), C.byref(cv), step.ctypes.data_as(C.POINTER(C.c_double)))
                assert np.allclose(step, cmd, rtol=0, atol=1e-15), (case, k)
            else:
                assert np.array_equal(cmd, targ_j)          # constant_vel=None: the target set by tcp_position_control stays
            P, Q = oracle.link_states(m, q)
            done = L.or_blocking_reached(dp(tpos), dp(torn), dp(P[m.tcp_link]), dp(Q[m.tcp_link]), dp(qd), C.c_int(6))
            assert bool(done) == (k == len(hist) - 1), (case, k)
        if cv0 > 0:
            assert cv.value < cv0                          # the halving rule fired on the way in


@pytest.mark.parametrize("arm,sign", [("mg400", 1), ("mg400", -1), ("ur5", 1), ("ur5", -1)])
def test_vertical_surface_geometry_and_rewards(oracle, monkeypatch, arm, sign):
    """surface_follow-v2's own surface (noise_mode "vertical_simplex", `forward` sensors) run from the reference source: the
    upright heightfield's surface_array / normals, goal, start pose and the reward terms - against the oracle's restatement
    (the CUDA path does not build this mode yet; this pins the checker it will be compared with)."""
    monkeypatch.setattr(oracle, "opensimplex_noise2", _fake_noise)
    key = "vert_%s_%s" % (arm, "p" if sign > 0 else "m")
    g = lambda n: GOLD[key + "_" + n]
    e = oracle.SurfaceFollowOracle(image_size=64, arm=arm, sensor="tactip", movement_mode="xRz", variant="vert", noise_mode="vertical_simplex",
                                   render=False, max_steps=200)
    assert e.typ == "forward"
    e.reset(draws=(77.0, float(sign)))
    assert np.allclose(np.stack([e.x_bins, e.y_bins]), g("bins"), atol=0) and np.allclose(e.surface_pos, g("surface_pos"), atol=0)
    assert np.allclose(e.h, g("h"), atol=1e-15) and np.ptp(e.h, axis=1).max() == 0 and np.ptp(e.h) > 1e-3       # varies along the rows only
    assert np.allclose(e.surface_array, g("array"), atol=1e-12)
    assert np.allclose(e.surface_normals, g("normals"), atol=1e-12)
    assert np.allclose(e.goal_pos, g("goal")[:3], atol=1e-12)
    gw, _ = oracle.world_to_work(e.m, e.goal_pos, np.array([0.0, 0.0, 0.0, 1.0]))
    assert np.allclose(gw, g("goal")[3:6], atol=1e-12)
    # update_init_pose (:556-563): the start pose handed to Robot.reset, in the work frame
    ch = e.h[32, 32]
    init_world = np.array([e.surface_pos[0] - (ch - e.embed_dist), e.surface_pos[1], e.surface_pos[2]])
    assert np.allclose(e.Rw.T @ (init_world - e.workframe_pos), g("init")[:3], atol=1e-12) and np.all(g("init")[3:] == 0)
    ended = 0
    for pose, row in zip(g("poses"), g("rows")):
        q = oracle.quat_from_euler(pose[3:6])
        e.tcp_world = lambda p=pose[:3], q=q: (p, q)
        e.steps = 7
        rew, done = e.step_data()
        assert (e.tip_i, e.tip_j) == (int(row[4]), int(row[5]))
        assert abs(rew - row[2]) < 1e-12 and done == bool(row[3])
        assert abs(rew + (10.0 * row[0] + 3.0 * row[1])) < 1e-12
        ended += int(done)
    assert ended == 1
    if arm == "ur5" and sign == -1:
        _check_actions("surfvert_vertical", e)        # x, y +-0.01 m/s, yaw +-5 deg/s (:183-194); e.dirs = (0, -1, 0) from the reset


@pytest.mark.skipif(not os.path.isdir("/root/reference/tactile_gym"), reason="needs the reference checkout (build container only)")
def test_committed_golden_is_what_the_reference_source_produces(tmp_path):
    """Where the reference is present, re-run it: the committed vectors must be exactly what tools/make_reference_golden.py and
    tools/make_reference_params.py produce from /root/reference today (no hand-edited or stale golden files)."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "ref.npz")
    subprocess.run([sys.executable, os.path.join(root, "tools", "make_reference_golden.py"), "/root/reference", out], check=True, capture_output=True)
    new = np.load(out)
    assert sorted(new.files) == sorted(GOLD.files)
    for k in new.files:
        assert new[k].shape == GOLD[k].shape and np.array_equal(new[k], GOLD[k]), k
    outj = str(tmp_path / "params.json")
    subprocess.run([sys.executable, os.path.join(root, "tools", "make_reference_params.py"), "/root/reference", outj], check=True, capture_output=True)
    assert json.load(open(outj)) == json.load(open(os.path.join(root, "tests", "golden", "reference_params.json")))


def test_balance_reset_scene(oracle):
    """object_balance's reset_task + reset_object (object_balance_env.py:295-381) run from the reference source: gravity, the
    constraint pivot, the pole's start position and the one-off 0.1 N push (where and which way), for three seeds"""
    for row in GOLD["balance_reset_rows"]:
        seed, log = int(row[0]), row[1:7]                       # draws: gravity, embed, choice, rand, choice, rand
        g, pivot, init_pos, force, fpos = row[7], row[8:11], row[11:14], row[14:17], row[17:20]
        b = oracle.ObjectBalanceOracle(image_size=64, seed=seed)
        b.reset()
        assert b.m.gravity[2] == g == log[0] and b.embed_dist == log[1]
        assert np.allclose(np.array(b.o.pivot_b[:]), pivot, atol=1e-15)
        assert np.allclose(b.init_obj_pos, init_pos, atol=1e-15) and np.allclose(np.array(b.o.pos[:]), init_pos, atol=1e-15)
        assert np.allclose(np.array(b.o.ext_force[:]), force, atol=0) and np.allclose(np.array(b.o.ext_pos[:]), fpos, atol=1e-15)
        assert b.o.ext_pending == 1


def test_push_and_roll_reset_scenes(oracle):
    """reset_object of object_push (object_push_env.py:196-229) and reset_task / update_workframe / reset_object / make_goal of
    object_roll (object_roll_env.py:176-256) run from the reference source: start poses, mass, radius, work frame, TCP-frame goal"""
    for row in GOLD["push_reset_rows"]:
        seed, ang, mass = int(row[0]), row[1], row[2]
        pos, orn, mass_set = row[3:6], row[6:10], row[10]
        p = oracle.ObjectPushOracle(image_size=64, arm="ur5", sensor="tactip", rand_init_orn=True, rand_obj_mass=True, seed=seed)
        p.reset()
        assert np.allclose(np.array(p.o.pos[:]), pos, atol=1e-15) and p.o.mass == mass == mass_set
        assert abs(abs(np.dot(np.array(p.o.quat[:]), orn)) - 1.0) < 1e-15
    for row in GOLD["roll_reset_rows"]:
        seed, (scale, embed, dx, dy, gang, gdist) = int(row[0]), row[1:7]
        radius, wpos, obj_pos, scaling, goal_tcp = row[7], row[8:11], row[11:14], row[14], row[15:18]
        r = oracle.ObjectRollOracle(image_size=64, rand_obj_size=True, rand_embed_dist=True, rand_init_obj_pos=True, seed=seed)
        r.reset()
        assert r.radius == radius and scaling == scale and r.embed_dist == embed
        assert np.allclose(r.workframe_pos, wpos, atol=1e-15) and np.allclose(np.array(r.o.pos[:]), obj_pos, atol=1e-15)
        assert np.allclose(r.goal_pos_tcp, goal_tcp, atol=1e-15)
        # the goal in the world for the TCP pose the generator injected (:258-286)
        r.tcp_world = lambda: (np.array([0.65, 0.0, 0.004]), oracle.quat_from_euler([-np.pi, 0.0, np.pi / 2]))
        r.update_goal()
        assert np.allclose(r.goal_pos_world, row[18:21], atol=1e-12)
