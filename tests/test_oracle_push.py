"""CPU: the object_push restatement (oracle/tg_oracle.c:or_step_sim_push, oracle/oracle.py:ObjectPushOracle).

pybullet is not available, so these are physics known-answers the contact solve must satisfy whatever bullet's narrow phase
does in detail (dynamics parity is unpinned): static equilibrium on the table, Coulomb sliding, the
stiffness / damping contact acting as the spring it encodes, trajectory construction against numpy, kinematic design
identities of the scene (rest pose <-> workframe <-> cube face)."""
import ctypes as C

import numpy as np


def _env(oracle, **kw):
    e = oracle.ObjectPushOracle(image_size=64, **kw)
    return e


def _substep(oracle, e):
    oracle.lib().or_step_sim_push(C.byref(e.m), C.byref(e.s), C.byref(e.o), C.byref(e.p))


def test_scene_identities(oracle):
    """rest pose puts the TCP at the workframe origin, which lies on the cube's near face; the tip core's
    apex sits 2.8 mm behind the TCP, so the core does not touch the cube at reset and the skin (at the TCP) just does"""
    for arm, sensor in [("mg400", "digitac"), ("ur5", "digitac"), ("ur5", "tactip"), ("ur5", "digit")]:
        e = _env(oracle, arm=arm, sensor=sensor)
        e.reset(draws=[0.0, 0.491, 123.0])
        p, _ = e.tcp_world()
        assert np.abs(p - e.workframe_pos).max() < 3e-4, (arm, sensor, p - e.workframe_pos)
        face_y = e.o.pos[1] - 0.04
        assert abs(face_y - e.workframe_pos[1]) < 1e-12
        _substep(oracle, e)
        assert e.p.n_contacts == 4, (arm, sensor)                        # table contacts only
        feat = e.features()
        # (the first goal sits exactly termination_pos_dist away: whether reset already advances it is a rounding-level tie)
        assert np.allclose(feat[:6], 0, atol=1e-3) and feat[6] in (0.065, 0.09) and feat[8] == 0


def test_cube_rests_on_the_table(oracle):
    e = _env(oracle)
    e.reset(draws=[0.0, 0.491, 1.0])
    v = np.zeros(6)
    oracle.lib().or_tcp_velocity_control(C.byref(e.m), C.byref(e.s), oracle._dptr(v))
    for _ in range(240):
        _substep(oracle, e)
    assert abs(e.o.pos[2] - 0.04) < 2e-6 and np.abs(np.array(e.o.pos[:2]) - e.init_obj_pos[:2]).max() < 2e-5   # creep of the truncated solve: ~6 um per second
    imp = np.array(e.p.normal_impulse[:4])
    # the four corner impulses carry the weight: sum = m g dt to within the solver's residual exit
    # (|dv| <= sqrt(1e-7) per row -> m * 3.2e-4)
    assert abs(imp.sum() - 0.491 * 9.81 / 240) < 0.491 * 3.2e-4 and (imp > 0).all()
    assert np.abs(np.array(e.o.omg[:])).max() < 5e-3                      # truncated-solve jitter only


def test_coulomb_sliding_deceleration(oracle):
    """a cube sliding freely on the table loses mu * g * dt of speed per substep (mu = 0.065 * 1.0)"""
    e = _env(oracle)
    e.reset(draws=[0.0, 0.491, 1.0])
    v = np.zeros(6)
    oracle.lib().or_tcp_velocity_control(C.byref(e.m), C.byref(e.s), oracle._dptr(v))
    e.m.solver_residual_threshold = 0.0                                   # converge, so the known answer is sharp
    e.p.lin_damping = e.p.ang_damping = 0.0
    e.o.vel[0] = 0.2                                                      # along world x: away from the tip
    speeds = []
    for _ in range(40):
        _substep(oracle, e)
        speeds.append(e.o.vel[0])
    dec = -np.diff(speeds)
    assert np.allclose(dec, 0.065 * 9.81 / 240, rtol=1e-6), dec[:3]
    assert np.abs(np.array(e.o.omg[:])).max() < 1e-9 and abs(e.o.vel[1]) < 1e-12
    for _ in range(200):
        _substep(oracle, e)
    assert abs(e.o.vel[0]) < 1e-9                                         # sticks once stopped (friction inside the cone)


def test_contact_stiffness_is_a_spring(oracle):
    """steady pushing at constant speed: tip force = table friction = k * penetration, per the erp / cfm pair bullet
    derives from contactStiffness / contactDamping (sensors/tactile_sensor.py:314-332)"""
    e = _env(oracle)
    e.reset(draws=[0.0, 0.491, 1.0])
    for _ in range(30):
        e.step(np.array([0.0, 0.0], dtype=np.float32))
    assert e.p.n_contacts > 4
    tip = np.array(e.p.normal_impulse[4:])
    f_tip = tip.sum() * 240
    assert abs(f_tip - 0.065 * 0.491 * 9.81) < 0.02                       # = friction force on the cube
    # penetration of the deepest hull vertex: F / k per touching point (points at the rim carry ~nothing)
    P, _ = oracle.link_states(e.m, np.array(e.s.q[: e.m.ndof]))
    fr_pos, fr_R = oracle.link_frames(e.m, np.array(e.s.q[: e.m.ndof]))
    hull_w = e.hull @ fr_R[e.p.tip_link].T + fr_pos[e.p.tip_link]
    depth = hull_w[:, 1].max() - (e.o.pos[1] - 0.04)                      # pushing along world +y
    assert 0.6 * f_tip / 300 < depth < 1.4 * f_tip / 300, (depth, f_tip / 300)
    # and the cube moves with the tip at the commanded 0.01 m/s (max_action along the tip axis, TyRz)
    assert abs(e.o.vel[1] - 0.01) < 5e-4


def test_trajectory_matches_numpy(oracle):
    e = _env(oracle, traj_type="simplex")
    e.reset(draws=[0.0, 0.491, 4242.0])
    y = np.array([oracle.opensimplex_noise2(4242, i * 0.1, 1) * 0.1 for i in range(10)])
    assert np.allclose(e.traj_pos_work[:, 1], y - y[0], atol=1e-15)
    assert np.allclose(e.traj_pos_work[:, 0], 0.065 + 0.025 * np.arange(10), atol=1e-15)
    assert np.allclose(e.traj_rpy_work[:, 2], np.gradient(y - y[0], 0.025), atol=1e-15)
    # work x = world y, work y = world x for the workframe rpy (-pi, 0, pi/2)
    assert np.allclose(e.traj_pos_world[:, 1] - e.workframe_pos[1], e.traj_pos_work[:, 0], atol=1e-12)
    assert np.allclose(e.traj_pos_world[:, 0] - e.workframe_pos[0], e.traj_pos_work[:, 1], atol=1e-12)
    s = _env(oracle, traj_type="straight")
    s.reset(draws=[0.0, 0.491, 0.3])
    assert np.allclose(s.traj_pos_work[:, 1], 0.025 * np.arange(10) * np.sin(0.3), atol=1e-15)
    assert np.allclose(s.traj_rpy_work[:, 2], np.sin(0.3), atol=1e-12)


def test_goals_advance_and_episode_ends(oracle):
    """dragging the cube along the trajectory by hand: +1 goal per waypoint reached, done after the last; sparse reward"""
    e = _env(oracle, reward_mode="sparse", max_steps=1000)
    e.reset(draws=[0.0, 0.491, 99.0])
    e.targ = -1; e.update_goal()
    seen = []
    for i in range(10):
        for c in range(3):
            e.o.pos[c] = e.traj_pos_world[i][c]
        r, d = e.step_data()
        seen.append((r, d, e.targ))
    assert [s[0] for s in seen] == [1.0] * 10
    assert [s[1] for s in seen] == [False] * 9 + [True]
    assert [s[2] for s in seen] == list(range(1, 10)) + [10]
    e.steps = 1000; e.o.pos[0] += 1.0
    assert e.step_data() == (0.0, True)                                   # max_steps


def test_push_draws_follow_the_reference_call_order(oracle):
    from tactile_gym_b200 import seeding
    from tactile_gym_b200.engine import object_push_draws

    for ro, rm, tt in [(False, False, "simplex"), (True, True, "simplex"), (True, False, "straight"), (False, True, "straight")]:
        a = object_push_draws(ro, rm, tt, 0.491)(seeding.np_random(5)[0], 7)
        rng = seeding.np_random(5)[0]
        b = np.array([oracle.push_draws(rng, ro, rm, tt) for _ in range(7)])
        assert np.array_equal(a, b)


def test_push_config_tables():
    """host-side TgConfig of BASELINE config 4 (no GPU needed to build it)"""
    from tactile_gym_b200 import _lib as L
    from tactile_gym_b200.engine import object_push_config

    modes = {"movement_mode": "TyRz", "control_mode": "TCP_velocity_control", "rand_init_orn": False, "rand_obj_mass": False,
             "traj_type": "simplex", "observation_mode": "tactile_and_feature", "reward_mode": "dense", "arm_type": "mg400",
             "tactile_sensor_name": "digitac"}
    cfg, keep, draw = object_push_config(modes, [128, 128], 1000, 8192)
    t = cfg.task
    assert t.task == L.TG_TASK_OBJECT_PUSH and t.push_mode == L.TG_PUSH_TCP_TYRZ and t.act_dim == 2 and cfg.phys.substeps == 24
    assert abs(t.push_mu_table - 0.065) < 1e-15 and abs(t.push_mu_tip - 0.65) < 1e-15
    assert abs(t.push_tip_k - 300) < 1e-9 and abs(t.push_tip_d - 100.1) < 1e-12
    assert list(t.push_init_pos) == [0.25, -0.1 + 0.04, 0.04] and cfg.n_tip_hull == 610 and cfg.sensor.n_prim == 6
    hull = keep[7]                       # (dep, gray, mask, tris, rest, prims, prim_nv, hull, parts, part_cen)
    assert hull.shape == (610, 3)
    # the hull rides on the TCP's body, a few mm behind the TCP point
    tcp = np.array(cfg.arm.tcp_pos[:])
    assert 0.002 < np.linalg.norm(hull - tcp, axis=1).min() < 0.004
    import pytest

    with pytest.raises(ValueError):
        object_push_config(dict(modes, traj_type="zigzag"), [64, 64], 10, 1)
    # MG400 + TacTip = the mini_right_angle sensor, workframe 5 cm further out (object_push_env.py:82-86)
    cfg2, keep2, _ = object_push_config(dict(modes, tactile_sensor_name="tactip"), [128, 128], 10, 1)
    assert abs(cfg2.task.workframe_pos[0] - 0.30) < 1e-15 and cfg2.n_tip_hull == 1089 and cfg2.sensor.fov_deg == 60
    assert abs(cfg2.task.push_tip_k - 50) < 1e-9 and abs(cfg2.task.push_tip_d - 100.1) < 1e-12


def test_mg400_mini_tactip_push(oracle):
    """object_push's MG400 + TacTip pairing (object_push_env.py:82-86: the mini_right_angle sensor, workframe x = 0.30) - the
    reference's own PPO set-up (sb3_helpers/params/object_push_params.py).  The 4-dof arm cannot reach the commanded
    orientation to blocking_move's 1e-3 rad (0.0014 rad is left), so the reset runs its full 1000 substeps (robot.py:188-260
    has no other exit); then the tip pushes the cube along the trajectory and the goals advance."""
    # (image size 128 as in the reference's set-up: its 64 / 256 fixtures for this sensor type were captured with the
    # right_angle camera - nodef depths 0.51+ instead of 0.0003+ - and show nothing)
    e = oracle.ObjectPushOracle(image_size=128, arm="mg400", sensor="tactip", seed=3)
    e.reset()
    assert e.typ == "mini_right_angle" and abs(e.workframe_pos[0] - 0.30) < 1e-15
    assert e.last_reset_substeps == 1000
    p, r = oracle.tcp_pose_workframe(e.m, np.array(e.s.q[: e.m.ndof]))
    assert np.abs(p).max() < 2e-4 and 1e-3 < np.linalg.norm(r) < 2e-3
    y0 = e.o.pos[1]
    touched = 0
    for k in range(32):
        o, rew, done, _ = e.step(np.array([0.0, 0.0], np.float32))
        touched = max(touched, int((o["tactile"][..., 0][e.ref[2] == 0] > 0).sum()))
    assert e.o.pos[1] - y0 > 0.02 and abs(e.o.pos[2] - 0.04) < 1e-3      # pushed 2+ cm along the work frame's x, still flat on the table
    assert touched > 100 and e.targ >= 1 and not done
