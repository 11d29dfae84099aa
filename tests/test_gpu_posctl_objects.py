"""GPU parity of control_mode="TCP_position_control" on the object tasks (robots/arms/robot.py:156-186,
base_robot_arm.py:228-279): the blocking move's steps are steps of the env's whole world - arm + constrained pole
(object_balance), arm + cube / marble with contacts (object_push, object_roll).  CUDA through the C ABI against the CPU oracle
(oracle/tg_oracle.c:or_tcp_position_control_world), each env step compared from an identical state, tolerances as in the
velocity-control tests of the same tasks (test_gpu_parity.py, test_gpu_push.py, test_gpu_roll.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _img_close(a, b):
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    return d.max(), (d != 0).mean()


def _sync(ref, row, nb):
    for k in range(nb):
        ref.s.q[k] = row[k]; ref.s.qd[k] = row[nb + k]
    o = row[2 * nb + 11:]
    for c in range(3):
        ref.o.pos[c] = o[c]; ref.o.vel[c] = o[7 + c]; ref.o.omg[c] = o[10 + c]
    for c in range(4):
        ref.o.quat[c] = o[3 + c]
    ref.steps = int(row[2 * nb + 9])


def _check_object(st_row, r, nb, tol, otol=None, quat_tol=None, twist_tol=3e-5, tag=None):
    """Joints to `tol` (1e-9 after the first step, as under velocity control).  The object is compared more loosely than under
    velocity control, for a reason that is a property of the algorithm, not of either implementation: the joint targets are the
    output of the 100-iteration damped IK, whose two implementations agree to ~1e-10 rad instead of the 1e-16 of the velocity
    targets, and that is enough for the PGS sweep - which stops on Bullet's residual threshold (1e-7 on the squared velocity
    change, i.e. velocities converged to ~3e-4 m/s) - to stop one sweep earlier or later on one side.  Measured on the B200:
    object positions still agree to 1e-9 for the cube and the pole (default otol = 20 tol), their twists to ~5e-6."""
    ob_ = st_row[2 * nb + 11:]
    otol = 20 * tol if otol is None else max(otol, 20 * tol)
    quat_tol = 10 * otol if quat_tol is None else max(quat_tol, 10 * otol)
    assert np.allclose(st_row[:nb], np.array(r.s.q[:nb]), atol=tol), tag
    assert np.allclose(ob_[:3], np.array(r.o.pos[:]), atol=otol), (tag, ob_[:3] - np.array(r.o.pos[:]))
    assert np.allclose(ob_[3:7], np.array(r.o.quat[:]), atol=quat_tol), tag
    assert np.allclose(ob_[7:13], np.array(list(r.o.vel[:]) + list(r.o.omg[:])), atol=max(twist_tol, 100 * otol)), tag


@pytest.mark.parametrize("S,movement", [(128, "xyRxRy"), (64, "xy")])
def test_object_balance_position_control(oracle, S, movement):
    import tactile_gym_b200 as tg

    modes = {"movement_mode": movement, "control_mode": "TCP_position_control", "object_mode": "pole", "rand_gravity": True,
             "rand_embed_dist": True, "observation_mode": "tactile", "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "tactip"}
    n, nb = 6, 6
    env = tg.make_vec("object_balance-v0", n, env_kwargs={"env_modes": modes, "image_size": [S, S], "max_steps": 250})
    rng = np.random.RandomState(S + 1)
    draws = np.stack([rng.uniform(-1.0, -0.1, (n, 2)), rng.uniform(0.003, 0.006, (n, 2)),
                      rng.choice([-1, 1], (n, 2)) * rng.rand(n, 2), rng.choice([-1, 1], (n, 2)) * rng.rand(n, 2)], axis=2)
    env.world.set_draws(draws)
    env.reset()
    st = env.world.get_state()
    refs = []
    for i in range(n):
        r = oracle.ObjectBalanceOracle(image_size=S, movement_mode=movement, control_mode="TCP_position_control")
        r.reset(draws=draws[i, 0])
        assert np.allclose(st[i, :6], np.array(r.s.q[:6]), atol=2e-6)
        _sync(r, st[i], nb)
        refs.append(r)
    alive = np.ones(n, dtype=bool)
    moved = np.zeros(n)
    tcp0 = st[:, 2 * nb:2 * nb + 3].copy()
    for k in range(30):
        act = rng.uniform(-0.25, 0.25, (n, env.world.act_dim)).astype(np.float32)
        o2, rew, done, infos = env.step(act)
        st = env.world.get_state()
        for i, r in enumerate(refs):
            if not alive[i]:
                continue
            o, rr, dd, _ = r.step(act[i])
            assert 1 <= r.last_move_substeps <= 10
            assert rr == rew[i] and bool(dd) == bool(done[i]), (k, i)
            if dd:
                alive[i] = False
                continue
            _check_object(st[i], r, nb, 5e-6 if k == 0 else 1e-9, tag=(k, i))
            moved[i] = np.linalg.norm(st[i, 2 * nb:2 * nb + 3] - tcp0[i])
            _sync(r, st[i], nb)
            mx, frac = _img_close(r.observation(), o2["tactile"][i])
            assert mx <= 1 and frac < 1e-3, (k, i, mx, frac)
    assert alive.sum() >= 3 and (moved[alive] > 2e-4).all()            # pose deltas of <= 1 mm per step really moved the tip
    assert not env.world.pipeline_error()
    env.close()


@pytest.mark.parametrize("arm,sensor,S,movement", [("mg400", "digitac", 128, "TyRz"), ("ur5", "tactip", 64, "xyRz")])
def test_object_push_position_control(oracle, arm, sensor, S, movement):
    import tactile_gym_b200 as tg

    modes = {"movement_mode": movement, "control_mode": "TCP_position_control", "rand_init_orn": False, "rand_obj_mass": False,
             "traj_type": "straight", "observation_mode": "tactile_and_feature", "reward_mode": "dense", "arm_type": arm,
             "tactile_sensor_name": sensor}
    n, nb = 6, (8 if arm == "mg400" else 6)
    env = tg.make_vec("object_push-v0", n, env_kwargs={"env_modes": modes, "image_size": [S, S], "max_steps": 1000})
    rng = np.random.RandomState(S + 7)
    draws = np.stack([np.zeros((n, 2)), np.full((n, 2), 0.491), rng.uniform(-np.pi / 8, np.pi / 8, (n, 2))], axis=2)
    env.world.set_draws(draws)
    env.reset()
    st = env.world.get_state()
    refs = []
    for i in range(n):
        r = oracle.ObjectPushOracle(image_size=S, arm=arm, sensor=sensor, movement_mode=movement, traj_type="straight",
                                    control_mode="TCP_position_control")
        r.reset(draws=draws[i, 0])
        assert np.allclose(st[i, :nb], np.array(r.s.q[:nb]), atol=2e-6)
        o = st[i, 2 * nb + 11:]
        d0 = np.linalg.norm(np.array(r.o.pos[:]) - r.traj_pos_world[0])
        if abs(d0 - r.termination_pos_dist) < 1e-12:                   # the rounding-level tie of test_gpu_push.py
            r.targ = int(o[14]) - 1
            r.update_goal()
        _sync(r, st[i], nb)
        refs.append(r)
    touched = np.zeros(n, dtype=bool)
    moved = np.zeros(n)
    for k in range(40):
        act = rng.uniform(-0.25, 0.25, (n, env.world.act_dim)).astype(np.float32)
        if movement == "xyRz":
            act[:, 0] = 0.25                                           # keep pushing forwards
        o2, rew, done, infos = env.step(act)
        st = env.world.get_state()
        for i, r in enumerate(refs):
            o, rr, dd, _ = r.step(act[i])
            assert 1 <= r.last_move_substeps <= 10
            touched[i] |= r.p.n_contacts > 4
            _check_object(st[i], r, nb, 5e-6 if k == 0 else 1e-9, tag=(k, i))
            assert abs(rr - rew[i]) < (1e-4 if k == 0 else 1e-6) and bool(dd) == bool(done[i]) and not dd, (k, i, rr, rew[i])
            ob_ = st[i, 2 * nb + 11:]
            assert int(ob_[14]) == r.targ, (k, i)
            moved[i] = np.linalg.norm(ob_[:2] - np.array(env.world.cfg.task.push_init_pos[:2]))
            _sync(r, st[i], nb)
            ref_obs = r.observation()
            mx, frac = _img_close(ref_obs["tactile"], o2["tactile"][i])
            assert mx <= 1 and frac < 2e-3, (k, i, mx, frac)
            assert np.allclose(ref_obs["extended_feature"], o2["extended_feature"][i], atol=1e-6), (k, i)
    assert touched.all() and (moved > 1e-3).all()                      # 1 mm per step for 40 steps: every cube was pushed
    assert not env.world.pipeline_error()
    env.close()


def test_object_roll_position_control(oracle):
    import tactile_gym_b200 as tg

    S, n, nb = 64, 6, 6
    modes = {"movement_mode": "xy", "control_mode": "TCP_position_control", "rand_init_obj_pos": True, "rand_obj_size": True,
             "rand_embed_dist": True, "observation_mode": "tactile_and_feature", "reward_mode": "dense", "arm_type": "ur5",
             "tactile_sensor_name": "tactip"}
    env = tg.make_vec("object_roll-v0", n, env_kwargs={"env_modes": modes, "image_size": [S, S], "max_steps": 250})
    rng = np.random.RandomState(11)
    draws = np.stack([rng.uniform(1.0, 2.0, (n, 2)), rng.uniform(0.0019, 0.003, (n, 2)), rng.uniform(-0.009, 0.009, (n, 2)),
                      rng.uniform(-0.009, 0.009, (n, 2)), rng.uniform(-np.pi, np.pi, (n, 2)), rng.uniform(0.0, 0.015, (n, 2))], axis=2)
    env.world.set_draws(draws)
    env.reset()
    st = env.world.get_state()
    refs = []
    for i in range(n):
        r = oracle.ObjectRollOracle(image_size=S, rand_obj_size=True, rand_embed_dist=True, rand_init_obj_pos=True, control_mode="TCP_position_control")
        r.reset(draws=draws[i, 0])
        assert np.allclose(st[i, :nb], np.array(r.s.q[:nb]), atol=2e-6)
        _sync(r, st[i], nb)
        refs.append(r)
    alive = np.ones(n, dtype=bool)
    rolled = np.zeros(n)
    for k in range(30):
        act = rng.uniform(-0.25, 0.25, (n, 2)).astype(np.float32)
        act[:, 0] = 0.25 * np.sign(draws[:, 0, 2] + 1e-9) * -1.0
        o2, rew, done, infos = env.step(act)
        st = env.world.get_state()
        for i, r in enumerate(refs):
            if not alive[i]:
                continue
            o, rr, dd, _ = r.step(act[i])
            assert 1 <= r.last_move_substeps <= 10
            # the 2.5 - 5 mm marble between two plates turns the sweep-count noise (see _check_object) into micrometres: it rolls
            # at half the plate's speed, whose solve is only converged to ~3e-4 m/s, over up to 10 substeps
            assert abs(rr - rew[i]) < 2e-5 and bool(dd) == bool(done[i]), (k, i, rr, rew[i])
            if dd:
                alive[i] = False
                continue
            _check_object(st[i], r, nb, 5e-6 if k == 0 else 1e-9, otol=2e-5, quat_tol=1e-2, twist_tol=0.2, tag=(k, i))
            ob_ = st[i, 2 * nb + 11:]
            rolled[i] = np.linalg.norm(ob_[:2] - np.array([0.65 + draws[i, 0, 2], draws[i, 0, 3]]))
            _sync(r, st[i], nb)
            ref_obs = r.observation()
            mx, frac = _img_close(ref_obs["tactile"], o2["tactile"][i])
            assert mx <= 1 and frac < 2e-3, (k, i, mx, frac)
            assert np.allclose(ref_obs["extended_feature"], o2["extended_feature"][i], atol=1e-7), (k, i)
    assert alive.sum() >= 3 and (rolled[alive] > 1e-4).all()                      # every marble was really rolled
    assert not env.world.pipeline_error()
    env.close()
