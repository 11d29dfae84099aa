"""CPU: the oracle against golden vectors computed by the REFERENCE'S OWN SOURCE (tests/golden/reference_numpy.npz, written by
tools/make_reference_golden.py, which compiles the reference's class bodies from /root/reference and runs their pure-numpy
methods): action encoding + scaling of every env / movement mode / control mode, the work-frame transforms, the edge reward and
termination geometry, the surface index lookup, distances and the three surface envs' dense rewards.  This pins the parts of
the oracle that restate reference Python (as opposed to pybullet's C++) to the reference itself."""
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
