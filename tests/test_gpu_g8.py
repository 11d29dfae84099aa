"""The 8-lanes-per-env physics substep (csrc/tg_g8.cuh: one body per lane, shuffle scans over the kinematic tree, per-lane columns
of M^-1, lane-parallel projected Gauss-Seidel) against the CPU oracle and against the one-thread-per-env kernel it replaces for
edge_follow / surface_follow.  Tolerances as in tests/test_gpu_parity.py: joints 1e-10, velocities 1e-9 per env step from an
identical state; the two CUDA formulations differ only in the order of a few sums (1e-13)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

EDGE = {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "noise_mode": "rand_height", "observation_mode": "tactile",
        "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "tactip"}


@pytest.mark.parametrize("arm,sensor", [("ur5", "tactip"), ("mg400", "digitac")])
def test_g8_substep_matches_oracle_and_one_thread_kernel(oracle, arm, sensor):
    import tactile_gym_b200 as tg
    from tactile_gym_b200 import _lib as L

    modes = dict(EDGE, arm_type=arm, tactile_sensor_name=sensor)
    env = tg.make_vec("edge_follow-v0", 4, env_kwargs={"env_modes": modes, "image_size": [64, 64], "max_steps": 200})
    w = env.world
    nb = w.nb
    wf = [0.33, 0, 0.035] if arm == "mg400" else [0.65, 0, 0.035]
    m = oracle.load_model(arm, sensor, "standard", wf, [-np.pi, 0, np.pi / 2], np.zeros((6, 2)))
    rest = oracle.rest_pose("edge_follow", arm, sensor, "standard", m)
    rng = np.random.RandomState(0)
    n = 70                                  # not a multiple of 4 / 16: ragged last warp and block
    q = rest + rng.uniform(-0.3, 0.3, (n, nb)); qd = rng.uniform(-1, 1, (n, nb)) * 0.2
    tv = rng.uniform(-0.05, 0.05, (n, nb))
    if arm == "mg400":                      # keep the parallelogram's slaved joints consistent (mg400.py:111-120)
        for a in (q, qd, tv):
            a[:, 5] = a[:, 1]; a[:, 6] = -a[:, 1]; a[:, 7] = a[:, 1] + a[:, 2]
    q8, qd8 = q.copy(), qd.copy()
    L.check(w.lib.tg_test_substep_g8(w.h, n, 24, q8.ctypes.data, qd8.ctypes.data, tv.ctypes.data))
    q1, qd1 = q.copy(), qd.copy()
    L.check(w.lib.tg_test_substep(w.h, n, 24, q1.ctypes.data, qd1.ctypes.data, tv.ctypes.data))
    assert np.abs(q8 - q1).max() < 1e-12 and np.abs(qd8 - qd1).max() < 1e-10, (np.abs(q8 - q1).max(), np.abs(qd8 - qd1).max())
    for i in range(n):
        s = oracle.OrState()
        for k in range(nb):
            s.q[k] = q[i, k]; s.qd[k] = qd[i, k]; s.motor_mode[k] = 0; s.target_vel[k] = tv[i, k]; s.kd[k] = 1.0; s.max_force[k] = 1000.0
        for _ in range(24):
            oracle.lib().or_step_sim(C.byref(m), C.byref(s))
        assert np.allclose(q8[i], np.array(s.q[:nb]), atol=1e-10), (i, np.abs(q8[i] - np.array(s.q[:nb])).max())
        assert np.allclose(qd8[i], np.array(s.qd[:nb]), atol=1e-9), (i, np.abs(qd8[i] - np.array(s.qd[:nb])).max())
    env.close()


@pytest.mark.parametrize("env_id,modes,act_dim", [
    ("edge_follow-v0", EDGE, 2),
    ("edge_follow-v0", dict(EDGE, arm_type="mg400", tactile_sensor_name="digitac"), 2),
    ("surface_follow-v0", {"movement_mode": "xyzRxRy", "control_mode": "TCP_velocity_control", "noise_mode": "simplex", "observation_mode": "tactile",
                           "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "digit"}, 3),
    # the pole on its point-to-point rows (g8_substep_obj): the rows' dot products are shuffle sums there, sequential sums in the
    # one-thread kernel - the same numbers to ~1e-12 over an episode
    ("object_balance-v0", {"movement_mode": "xyRxRy", "control_mode": "TCP_velocity_control", "object_mode": "pole", "rand_gravity": True,
                           "rand_embed_dist": True, "observation_mode": "tactile", "reward_mode": "dense", "arm_type": "ur5",
                           "tactile_sensor_name": "tactip"}, 4),
])
def test_g8_step_kernel_equals_one_thread_step_kernel(env_id, modes, act_dim):
    """whole env steps, episode ends and auto-resets included: the two kernels give the same states (1e-11), rewards, dones and
    images (the raster reads cameras that differ by ~1e-15: at most a stray LSB)"""
    import tactile_gym_b200 as tg

    n, S, steps = 37, 64, 12
    out = {}
    for lanes in (-8, 8):
        env = tg.make_vec(env_id, n, seed=11, env_kwargs={"env_modes": modes, "image_size": [S, S], "max_steps": 5}, lanes_per_warp=lanes)
        env.reset()
        rng = np.random.RandomState(4)
        rec = []
        for k in range(steps):
            o, r, d, infos = env.step(rng.uniform(-0.25, 0.25, (n, act_dim)).astype(np.float32))
            rec.append((o["tactile"].copy(), r.copy(), d.copy(), env.world.get_state()))
        assert not env.world.pipeline_error()
        out[lanes] = rec
        env.close()
    for k in range(steps):
        o8, r8, d8, s8 = out[-8][k]; o1, r1, d1, s1 = out[8][k]
        assert np.array_equal(d8, d1) and np.abs(r8 - r1).max() < 1e-6, k
        assert np.abs(s8 - s1).max() < 1e-9, (k, np.abs(s8 - s1).max())
        diff = np.abs(o8.astype(int) - o1.astype(int))
        assert diff.max() <= 1 and (diff != 0).mean() < 1e-4, (k, diff.max())
    assert any(out[8][k][2].any() for k in range(steps))
