"""CPU: the object_roll restatement (oracle/oracle.py:ObjectRollOracle on or_step_sim_push, shape 1) - physics and geometry
known-answers (pybullet is not available)."""
import ctypes as C

import numpy as np


def test_scene_identities(oracle):
    """the workframe sits embed_dist below the marble's top (object_roll_env.py:197-202), the TCP reaches it, and the flat
    tip's core (a cylinder whose cap is 1.75 mm behind the TCP) only touches the marble when embed_dist > 1.75 mm"""
    for scale, embed, touching in [(1.0, 0.0015, False), (1.0, 0.0028, True), (2.0, 0.0028, True)]:
        e = oracle.ObjectRollOracle(image_size=64)
        e.reset(draws=[scale, embed, 0.0, 0.0, 0.3, 0.01])
        p, _ = e.tcp_world()
        assert np.abs(p - e.workframe_pos).max() < 3e-4
        assert abs(e.workframe_pos[2] - (2 * 0.0025 * scale - embed)) < 1e-15
        oracle.lib().or_step_sim_push(C.byref(e.m), C.byref(e.s), C.byref(e.o), C.byref(e.p))
        assert e.p.n_contacts == (2 if touching else 1), (scale, embed, e.p.n_contacts)
        assert np.allclose(e.features(), [0.01 * np.cos(0.3), 0.01 * np.sin(0.3), 0.0])


def test_marble_rolls_at_half_the_plate_speed(oracle):
    e = oracle.ObjectRollOracle(image_size=64)
    e.reset(draws=[1.0, 0.0028, 0.0, 0.0, 0.0, 0.014])
    a = np.array([0.25, 0.0], dtype=np.float32)
    for _ in range(6):
        e.step(a)
    p0, t0 = np.array(e.o.pos[:]), e.tcp_world()[0].copy()
    for _ in range(10):
        e.step(a)
    p1, t1 = np.array(e.o.pos[:]), e.tcp_world()[0]
    assert abs(np.linalg.norm(t1 - t0) - 0.01) < 2e-4
    assert abs(np.linalg.norm(p1[:2] - p0[:2]) / np.linalg.norm(t1[:2] - t0[:2]) - 0.5) < 0.03
    # rolling without slipping on the table: v = omega x r
    v, w = np.array(e.o.vel[:]), np.array(e.o.omg[:])
    assert np.allclose(v[:2], np.cross(w, [0, 0, 0.0025])[:2] * -1.0, atol=2e-4) or np.allclose(v[:2], np.cross(w, [0, 0, 0.0025])[:2], atol=2e-4)
    assert abs(p1[2] - 0.0025) < 5e-5                                              # stays on the table


def test_goal_follows_the_tcp_and_terminates(oracle):
    e = oracle.ObjectRollOracle(image_size=64, reward_mode="sparse")
    e.reset(draws=[1.0, 0.0015, 0.0, 0.0, 1.0, 0.01])
    tp, _ = e.tcp_world()
    e.update_goal()
    # work x = world y for the workframe rpy (-pi, 0, pi/2); the TCP frame is aligned with the work frame at reset
    assert abs(np.linalg.norm(e.goal_pos_world[:2] - tp[:2]) - 0.01) < 1e-9
    r, d = e.step_data()
    assert (r, d) == (0.0, False)
    e.o.pos[0], e.o.pos[1] = e.goal_pos_world[0] + 5e-4, e.goal_pos_world[1]
    assert e.step_data() == (1.0, True)


def test_sphere_image(oracle):
    """analytic sphere render: a disc centred under the tip, radially symmetric, growing with the marble, peak value =
    the t_s_camera quantisation of the depth difference at the top of the marble"""
    e = oracle.ObjectRollOracle(image_size=128)
    e.reset(draws=[1.0, 0.0028, 0.0, 0.0, 0.0, 0.01])
    img = e.observation()["tactile"][..., 0].astype(int)
    skin = e.ref[2] == 0
    ys, xs = np.nonzero((img > 0) & skin)
    assert 100 < len(ys) < 1500
    cy, cx = ys.mean(), xs.mean()
    assert abs(cy - 63.5) < 2.5 and abs(cx - 63.5) < 2.5
    rr = np.hypot(ys - cy, xs - cx)
    assert rr.max() - rr.min() < rr.max() + 1 and np.corrcoef(rr, img[ys, xs])[0, 1] < -0.9   # brightest in the middle
    e2 = oracle.ObjectRollOracle(image_size=128)
    e2.reset(draws=[2.0, 0.0028, 0.0, 0.0, 0.0, 0.01])
    img2 = e2.observation()["tactile"][..., 0].astype(int)
    assert ((img2 > 0) & skin).sum() > 1.5 * len(ys)
    # the top of the marble is embed_dist past the skin plane at the TCP: the same depth whatever the marble's size
    assert abs(img2[skin].max() - img[skin].max()) <= 3


def test_roll_draws_and_config(oracle):
    from tactile_gym_b200 import _lib as L, seeding
    from tactile_gym_b200.engine import object_roll_config, object_roll_draws

    for flags in [(False, False, False), (True, True, True), (True, False, True)]:
        a = object_roll_draws(*flags)(seeding.np_random(3)[0], 6)
        rng = seeding.np_random(3)[0]
        b = np.array([oracle.roll_draws(rng, *flags) for _ in range(6)])
        assert np.array_equal(a, b)
    modes = {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "rand_init_obj_pos": True, "rand_obj_size": True,
             "rand_embed_dist": True, "observation_mode": "tactile_and_feature", "reward_mode": "dense", "arm_type": "ur5",
             "tactile_sensor_name": "tactip"}
    cfg, keep, draw = object_roll_config(modes, [128, 128], 250, 16)
    t = cfg.task
    assert t.task == L.TG_TASK_OBJECT_ROLL and t.push_shape == 1 and t.n_draws == 6 and cfg.sensor.n_prim == 0
    assert t.push_mu_table == 10.0 and t.push_mu_tip == 10.0 and abs(t.push_tip_k - 10.0) < 1e-9
    ax = np.array(t.roll_cyl_axis[:])
    assert abs(np.linalg.norm(ax) - 1) < 1e-12
    # the cap of the core is 1.75 mm behind the TCP along the cylinder axis
    d = np.dot(np.array(cfg.arm.tcp_pos[:]) - np.array(t.roll_cyl_pos[:]), ax)
    assert abs(abs(d) - t.roll_cyl_half_len - 0.00175) < 2e-5
    import pytest

    with pytest.raises(NotImplementedError):
        object_roll_config(dict(modes, tactile_sensor_name="digit"), [64, 64], 10, 1)
