"""GPU parity of object_roll-v0: marble between the table and the flat TacTip - the CUDA path through the
C ABI against the CPU oracle, each step compared from an identical state (tolerances as in test_gpu_push.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROLL_MODES = {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "rand_init_obj_pos": True, "rand_obj_size": True,
              "rand_embed_dist": True, "observation_mode": "tactile_and_feature", "reward_mode": "dense", "arm_type": "ur5",
              "tactile_sensor_name": "tactip"}


def _img_close(a, b):
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    return d.max(), (d != 0).mean()


def _sync(ref, row, nb=6):
    for k in range(nb):
        ref.s.q[k] = row[k]; ref.s.qd[k] = row[nb + k]
    o = row[2 * nb + 11:]
    for c in range(3):
        ref.o.pos[c] = o[c]; ref.o.vel[c] = o[7 + c]; ref.o.omg[c] = o[10 + c]
    for c in range(4):
        ref.o.quat[c] = o[3 + c]
    ref.steps = int(row[2 * nb + 9])


@pytest.mark.parametrize("S,reward", [(128, "dense"), (64, "sparse")])
def test_object_roll_matches_oracle(oracle, S, reward):
    import tactile_gym_b200 as tg

    modes = dict(ROLL_MODES, reward_mode=reward)
    n, nb = 6, 6
    env = tg.make_vec("object_roll-v0", n, env_kwargs={"env_modes": modes, "image_size": [S, S], "max_steps": 250})
    rng = np.random.RandomState(S)
    draws = np.stack([rng.uniform(1.0, 2.0, (n, 2)), rng.uniform(0.0019, 0.003, (n, 2)), rng.uniform(-0.009, 0.009, (n, 2)),
                      rng.uniform(-0.009, 0.009, (n, 2)), rng.uniform(-np.pi, np.pi, (n, 2)), rng.uniform(0.0, 0.015, (n, 2))], axis=2)
    env.world.set_draws(draws)
    ob = env.reset()
    st = env.world.get_state()
    refs = []
    for i in range(n):
        r = oracle.ObjectRollOracle(image_size=S, rand_obj_size=True, rand_embed_dist=True, rand_init_obj_pos=True, reward_mode=reward)
        r.reset(draws=draws[i, 0])
        refs.append(r)
        assert np.allclose(st[i, :nb], np.array(r.s.q[:nb]), atol=2e-6)
        o = st[i, 2 * nb + 11:]
        assert np.allclose(o[:3], np.array(r.o.pos[:]), atol=1e-12) and np.allclose(o[3:7], np.array(r.o.quat[:]), atol=1e-12)
        _sync(r, st[i])
        ref_obs = r.observation()
        mx, frac = _img_close(ref_obs["tactile"], ob["tactile"][i])
        assert mx <= 1 and frac < 2e-3, (i, mx, frac)
        assert ob["extended_feature"].shape == (n, 3)
        assert np.allclose(ref_obs["extended_feature"], ob["extended_feature"][i], atol=1e-7)
        assert (ob["tactile"][i][..., 0][r.ref[2] == 0] > 0).sum() > 20           # the marble shows in the image
    rolled = np.zeros(n)
    touched = np.zeros(n, dtype=bool)
    alive = np.ones(n, dtype=bool)
    for k in range(30):
        act = rng.uniform(-0.25, 0.25, (n, 2)).astype(np.float32)
        act[:, 0] = 0.25 * np.sign(draws[:, 0, 2] + 1e-9) * -1.0                     # roll back towards the centre
        o2, rew, done, infos = env.step(act)
        st = env.world.get_state()
        for i, r in enumerate(refs):
            if not alive[i]:
                continue
            o, rr, dd, _ = r.step(act[i])
            touched[i] |= r.p.n_contacts == 2
            ob_ = st[i, 2 * nb + 11:]
            tol = 5e-6 if k == 0 else 1e-9
            assert abs(rr - rew[i]) < (1e-5 if k == 0 else 1e-6) and bool(dd) == bool(done[i]), (k, i, rr, rew[i])
            if dd:                     # the marble reached the goal: the device env auto-reset, the comparison of this env ends
                alive[i] = False
                mx, frac = _img_close(o["tactile"], infos[i]["terminal_observation"]["tactile"])
                assert mx <= 1 and frac < 2e-3, (k, i, mx, frac)
                continue
            assert np.allclose(st[i, :nb], np.array(r.s.q[:nb]), atol=tol), (k, i)
            assert np.allclose(ob_[:3], np.array(r.o.pos[:]), atol=tol), (k, i, ob_[:3] - np.array(r.o.pos[:]))
            assert np.allclose(ob_[3:7], np.array(r.o.quat[:]), atol=tol * 1e3), (k, i)   # |omega| ~ 1 rad/s: angles amplify by 1 / r
            assert np.allclose(ob_[7:13], np.array(list(r.o.vel[:]) + list(r.o.omg[:])), atol=max(tol, 1e-8) * 1e3), (k, i)
            rolled[i] = np.linalg.norm(ob_[:2] - np.array([0.65 + draws[i, 0, 2], draws[i, 0, 3]]))
            _sync(r, st[i])
            ref_obs = r.observation()
            mx, frac = _img_close(ref_obs["tactile"], o2["tactile"][i])
            assert mx <= 1 and frac < 2e-3, (k, i, mx, frac)
            assert np.allclose(ref_obs["extended_feature"], o2["extended_feature"][i], atol=1e-7), (k, i)
    assert touched.all() and (rolled > 1e-3).all() and alive.sum() >= 3            # every marble was really rolled (> 1 mm)
    assert not env.world.pipeline_error()
    env.close()


def test_object_roll_rolls_at_half_the_tip_speed(oracle):
    """rolling without slipping between two plates: the marble's centre moves at half the speed of the plate on top"""
    import tactile_gym_b200 as tg

    n = 4
    modes = dict(ROLL_MODES, rand_init_obj_pos=False, rand_obj_size=False)
    env = tg.make_vec("object_roll-v0", n, env_kwargs={"env_modes": modes, "image_size": [64, 64], "max_steps": 250})
    draws = np.tile(np.array([1.0, 0.0028, 0.0, 0.0, 0.0, 0.014]), (n, 2, 1))
    env.world.set_draws(draws)
    env.reset()
    act = np.tile(np.array([[0.25, 0.0]], dtype=np.float32), (n, 1))
    for _ in range(6):
        env.step(act)
    a = env.world.get_state()
    for _ in range(10):
        env.step(act)
    b = env.world.get_state()
    tcp_d = np.linalg.norm(b[:, 12:14] - a[:, 12:14], axis=1)                       # tcp_pos x, y (state layout: 2 nb = 12)
    obj_d = np.linalg.norm(b[:, 23:25] - a[:, 23:25], axis=1)
    assert np.allclose(tcp_d, 0.01, rtol=0.02)                                      # 0.01 m/s for 1 s
    assert np.allclose(obj_d / tcp_d, 0.5, atol=0.03)
    env.close()


def test_object_roll_gym_env_surface():
    import tactile_gym_b200 as tg

    env = tg.make("object_roll-v0", env_modes=ROLL_MODES, image_size=[64, 64], max_steps=20)
    assert env.observation_space.spaces["extended_feature"].shape == (3,) and env.action_space.shape == (2,)
    o = env.reset()
    assert o["tactile"].shape == (64, 64, 1) and o["extended_feature"].shape == (3,)
    o, r, d, info = env.step(np.array([0.1, 0.0], dtype=np.float32))
    assert r <= 0 and info == {}
    env.close()
