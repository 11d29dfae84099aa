"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs.

Tolerances (north_star allows image L-inf 2/255, pose and reward 1e-3):
  * one env step from an IDENTICAL state: joint positions 1e-10, velocities 1e-9, TCP 1e-9, reward 1e-6
    (fp64 on both sides, different formulations: bullet-style ABA over 11 links vs CRBA over 6 bodies);
  * reset: 2e-6 rad.  Bullet's IK takes its orientation error from 2*acos(w) of a quaternion product, which
    amplifies rounding noise to ~1e-7 rad near convergence (acos(1 - eps) ~ sqrt(2 eps)); both sides carry
    that noise, so tighter agreement is not meaningful;
  * free-running episodes: 1e-5 (the reset noise persists, it does not grow);
  * tactile image: <= 1 LSB everywhere and identical in >= 99.8 % of the pixels.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _world(edge_modes, n, S=128, max_steps=200, lanes=0, rng=None):
    import tactile_gym_b200 as tg

    return tg.make_vec("edge_follow-v0", n, env_kwargs={"env_modes": edge_modes, "image_size": [S, S], "max_steps": max_steps},
                       lanes_per_warp=lanes, rng=rng)


def _draws(rng, n, rounds=4):
    return np.stack([rng.uniform(0.0015, 0.0065, (n, rounds)), rng.uniform(-np.pi, np.pi, (n, rounds))], axis=2)


def _img_close(a, b):
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    return d.max(), (d != 0).mean()


def _sync_oracle(ref, st_row):
    """copy one env's GPU state into an oracle env"""
    for k in range(6):
        ref.s.q[k] = st_row[k]
        ref.s.qd[k] = st_row[6 + k]
    ref.steps = int(st_row[21])


def test_dynamics_hooks_match_oracle(oracle, edge_modes):
    env = _world(edge_modes, 4)
    w = env.world
    m = oracle.load_model("ur5", "tactip", "standard", [0.65, 0, 0.035], [-np.pi, 0, np.pi / 2], np.zeros((6, 2)))
    rest = oracle.rest_pose("edge_follow", "ur5", "tactip", "standard", m)
    rng = np.random.RandomState(0)
    n = 64
    q = rest + rng.uniform(-0.5, 0.5, (n, 6)); qd = rng.uniform(-1, 1, (n, 6))
    from tactile_gym_b200 import _lib as L

    tau = np.zeros((n, 6))
    L.check(w.lib.tg_test_inverse_dynamics(w.h, n, q.ctypes.data, qd.ctypes.data, tau.ctypes.data))
    M = np.zeros((n, 6, 6))
    L.check(w.lib.tg_test_mass_matrix(w.h, n, q.ctypes.data, M.ctypes.data))
    for i in range(n):
        assert np.allclose(tau[i], oracle.inverse_dynamics(m, q[i], qd[i]), atol=1e-9)
        assert np.allclose(M[i] @ oracle.mass_matrix_inverse(m, q[i]), np.eye(6), atol=1e-8)
    # 24 substeps with velocity motors, including an initial velocity far from the target (many PGS sweeps)
    tv = rng.uniform(-0.05, 0.05, (n, 6))
    q2, qd2 = q.copy(), qd.copy() * 0.2
    q_in, qd_in = q2.copy(), qd2.copy()
    L.check(w.lib.tg_test_substep(w.h, n, 24, q2.ctypes.data, qd2.ctypes.data, tv.ctypes.data))
    for i in range(n):
        s = oracle.OrState()
        for k in range(6):
            s.q[k] = q_in[i, k]; s.qd[k] = qd_in[i, k]; s.motor_mode[k] = 0; s.target_vel[k] = tv[i, k]; s.kd[k] = 1.0; s.max_force[k] = 1000.0
        for _ in range(24):
            oracle.lib().or_step_sim(C.byref(m), C.byref(s))
        assert np.allclose(q2[i], np.array(s.q[:6]), atol=1e-10)
        assert np.allclose(qd2[i], np.array(s.qd[:6]), atol=1e-9)
    env.close()


@pytest.mark.parametrize("S", [64, 128, 256])
def test_reset_and_steps_match_oracle(oracle, edge_modes, S):
    n, steps = 12, 6
    env = _world(edge_modes, n, S=S)
    rng = np.random.RandomState(S)
    draws = _draws(rng, n)
    env.world.set_draws(draws)
    obs = env.reset()["tactile"]
    st = env.world.get_state()
    refs = []
    for i in range(n):
        r = oracle.EdgeFollowOracle(image_size=S)
        o = r.reset(draws=tuple(draws[i, 0]))
        refs.append(r)
        assert np.allclose(st[i, :6], np.array(r.s.q[:6]), atol=2e-6)
        assert abs(int(st[i, 22]) - r.last_reset_substeps) <= 1
        # the image is compared at the SAME joint state (the 1e-7 rad reset noise can move a pixel on a steep
        # side face across a quantisation step): render the oracle at the GPU's state
        _sync_oracle(r, st[i])
        mx, frac = _img_close(r.observation(), obs[i])
        assert mx <= 1 and frac < 1e-3, (i, mx, frac)
        assert (obs[i][..., 0][r.ref[2] == 0] > 0).sum() > 20   # something is actually pressed into the skin
    for k in range(steps):
        act = rng.uniform(-0.3, 0.3, (n, 2)).astype(np.float32)   # beyond +-0.25: exercises the clip
        for i, r in enumerate(refs):
            _sync_oracle(r, st[i])
        o2, rew, done, infos = env.step(act)
        st = env.world.get_state()
        for i, r in enumerate(refs):
            o, rr, dd, _ = r.step(act[i])
            assert np.allclose(st[i, :6], np.array(r.s.q[:6]), atol=1e-10)
            assert np.allclose(st[i, 6:12], np.array(r.s.qd[:6]), atol=1e-9)
            p, qq = r.tcp_world()
            assert np.allclose(st[i, 12:15], p, atol=1e-9)
            assert abs(rr - rew[i]) < 1e-6 and bool(dd) == bool(done[i])
            mx, frac = _img_close(o, o2["tactile"][i])
            assert mx <= 1 and frac < 2e-3, (k, i, mx, frac)
    env.close()


def test_free_running_episode_stays_close(oracle, edge_modes):
    n, S = 6, 64
    env = _world(edge_modes, n, S=S)
    rng = np.random.RandomState(11)
    draws = _draws(rng, n)
    env.world.set_draws(draws)
    env.reset()
    refs = [oracle.EdgeFollowOracle(image_size=S) for _ in range(n)]
    for i, r in enumerate(refs):
        r.reset(draws=tuple(draws[i, 0]))
    exact = []
    for k in range(60):
        act = rng.uniform(-0.25, 0.25, (n, 2)).astype(np.float32)
        o2, rew, done, _ = env.step(act)
        for i, r in enumerate(refs):
            o, rr, dd, _ = r.step(act[i])
            assert abs(rr - rew[i]) < 1e-4 and bool(dd) == bool(done[i])
            mx, frac = _img_close(o, o2["tactile"][i])
            assert mx <= 1
            exact.append(1 - frac)
    st = env.world.get_state()
    for i, r in enumerate(refs):
        assert np.allclose(st[i, :6], np.array(r.s.q[:6]), atol=1e-5)
        assert np.allclose(st[i, 12:15], r.tcp_world()[0], atol=1e-5)
    assert np.mean(exact) > 0.998
    assert not env.world.pipeline_error()
    env.close()


def test_tcp_limits_zero_the_velocity(oracle, edge_modes):
    """check_TCP_vel_lims (base_robot_arm.py:357-380): drive into the +x limit (0.175 m) and stay there."""
    env = _world(edge_modes, 2, S=64, max_steps=10000)
    env.world.set_draws(np.array([[[0.0035, 0.3]], [[0.0035, -2.0]]]))
    env.reset()
    refs = [oracle.EdgeFollowOracle(image_size=64, max_steps=10000) for _ in range(2)]
    refs[0].reset(draws=(0.0035, 0.3)); refs[1].reset(draws=(0.0035, -2.0))
    act = np.array([[0.25, 0.0], [0.25, 0.25]], dtype=np.float32)
    for k in range(190):
        _, _, done, _ = env.step(act)
        for i in range(2):
            refs[i].step(act[i])
    st = env.world.get_state()
    for i in range(2):
        assert np.allclose(st[i, :6], np.array(refs[i].s.q[:6]), atol=1e-5)
        p, _ = oracle.tcp_pose_workframe(refs[i].m, np.array(refs[i].s.q[:6]))
        assert 0.175 < p[0] < 0.1775
    env.close()


def test_autoreset_and_terminal_observation(oracle, edge_modes, monkeypatch):
    """3-step episodes: every episode ends while its next episode is still being rebuilt in chunks by the standby
    blocks, so this drives the slot hand-over (claim / wait / complete inline) as well as the VecEnv semantics.  (The quanta are
    pinned small here: by default tg_create sizes them from max_steps so that such short episodes rebuild in one go.)"""
    monkeypatch.setenv("TG_IK_CHUNK", "8")
    monkeypatch.setenv("TG_RESET_CHUNK", "6")
    n, S = 6, 64
    env = _world(edge_modes, n, S=S, max_steps=3)
    rng = np.random.RandomState(5)
    draws = _draws(rng, n, rounds=4)
    env.world.set_draws(draws)
    env.reset()
    refs = []
    for i in range(n):
        r = oracle.EdgeFollowOracle(image_size=S, max_steps=3)
        r.reset(draws=tuple(draws[i, 0]))
        refs.append(r)
    for ep in range(3):
        for k in range(3):
            act = rng.uniform(-0.25, 0.25, (n, 2)).astype(np.float32)
            obs, rew, done, infos = env.step(act)
            last = [r.step(act[i]) for i, r in enumerate(refs)]
            assert done.all() == (k == 2) and done.any() == (k == 2)
        for i, r in enumerate(refs):
            assert _img_close(infos[i]["terminal_observation"]["tactile"], last[i][0])[0] <= 1
            assert infos[i]["episode"]["l"] == 3
            o = r.reset(draws=tuple(draws[i, ep + 1]))      # next round of draws
            assert _img_close(o, obs["tactile"][i])[0] <= 1
        st = env.world.get_state()
        assert (st[:, 21] == 0).all()                       # step counters restarted
        for i, r in enumerate(refs):
            assert np.allclose(st[i, :6], np.array(r.s.q[:6]), atol=2e-6)
            assert st[i, 22] == r.last_reset_substeps
    assert not env.world.pipeline_error()
    assert env.world.pipeline_stalls() > 0                  # the episodes were shorter than a standby rebuild
    env.close()


def test_standby_rebuild_hides_behind_steps(oracle, edge_modes):
    """Episodes longer than a rebuild: every finished env finds its next episode READY (no stall), and the episode it
    starts is the oracle's for the same draws."""
    n, S = 40, 64
    env = _world(edge_modes, n, S=S, max_steps=40)
    rng = np.random.RandomState(11)
    draws = _draws(rng, n, rounds=4)
    env.world.set_draws(draws)
    env.reset()
    for ep in range(3):
        for k in range(40):
            obs, rew, done, infos = env.step(rng.uniform(-0.25, 0.25, (n, 2)).astype(np.float32))
        assert done.all()
        st = env.world.get_state()
        for i in (0, 7, 39):
            r = oracle.EdgeFollowOracle(image_size=S, max_steps=40)
            o = r.reset(draws=tuple(draws[i, ep + 1]))
            assert np.allclose(st[i, :6], np.array(r.s.q[:6]), atol=2e-6)
            assert _img_close(o, obs["tactile"][i])[0] <= 1
    assert env.world.pipeline_stalls() == 0
    env.close()


def test_lane_packing_is_invisible(edge_modes):
    """1 ... 32 active lanes per warp is a scheduling choice only: identical results."""
    res = []
    for lanes in (1, 4, 8, 32):
        env = _world(edge_modes, 40, S=64, lanes=lanes)
        rng = np.random.RandomState(1)
        env.world.set_draws(_draws(rng, 40))
        env.reset()
        for _ in range(3):
            env.step(rng.uniform(-0.25, 0.25, (40, 2)).astype(np.float32))
        res.append((env.world.get_state(), env.world.obs.cpu().numpy()))
        env.close()
    for st, ob in res[1:]:
        assert np.array_equal(st, res[0][0]) and np.array_equal(ob, res[0][1])


def test_seeded_vec_env_reproduces_gym_rng_stream(oracle, edge_modes):
    """env i seeded with seed + i, draws in the reference's order: embed_dist then edge_ang (edge_follow_env.py:293,240)."""
    n = 5
    env = _world(edge_modes, n, S=64)
    env.seed(100)
    env.reset()
    st = env.world.get_state()
    for i in range(n):
        rng = oracle.gym_np_random(100 + i)
        embed = rng.uniform(0.0015, 0.0065); ang = rng.uniform(-np.pi, np.pi)
        assert st[i, 19] == embed and st[i, 20] == ang
    env.close()


def test_gym_env_surface(edge_modes):
    import tactile_gym_b200 as tg

    env = tg.make("edge_follow-v0", max_steps=5, image_size=[64, 64], env_modes=edge_modes)
    assert env.action_space.shape == (2,) and env.observation_space["tactile"].shape == (64, 64, 1)
    assert env.seed(3) == [3]
    o = env.reset()
    assert o["tactile"].dtype == np.uint8 and o["tactile"].shape == (64, 64, 1)
    for k in range(5):
        o, r, d, info = env.step(env.action_space.sample())
        assert isinstance(r, float) and isinstance(d, bool) and info == {}
    assert d is True
    env.close()


def test_full_size_properties(edge_modes):
    """BASELINE config 2 size (4096 x 128^2): size-independent properties instead of an oracle run."""
    import torch

    n = 4096
    env = _world(edge_modes, n, S=128)
    env.seed(1)
    obs = env.reset()["tactile"]
    from tactile_gym_b200 import scene

    dep, gray, mask = scene.load_refimg("tactip", "standard", 128)
    # border pixels are the baked grey image for every env; the edge is pressed into every skin
    assert (obs[:, mask == 1, 0] == gray[mask == 1].astype(np.uint8)[None]).all()
    assert (obs[:, mask == 0, 0] > 0).any(axis=1).all()
    # idempotence: rendering the same state twice gives the same bytes
    a = env.world.raster_only().clone()
    b = env.world.raster_only()
    assert torch.equal(a, b)
    # zero action keeps every env where it is
    st0 = env.world.get_state()
    env.step(np.zeros((n, 2), dtype=np.float32))
    st1 = env.world.get_state()
    assert np.abs(st1[:, 12:15] - st0[:, 12:15]).max() < 5e-6
    # opposite actions bring the TCP back (to first order) and move it by |v| * 0.1 s
    act = np.tile(np.array([[0.2, -0.1]], dtype=np.float32), (n, 1))
    env.step(act); st2 = env.world.get_state()
    env.step(-act); st3 = env.world.get_state()
    assert np.abs(st3[:, 12:15] - st1[:, 12:15]).max() < 5e-6
    assert np.allclose(np.linalg.norm(st2[:, 12:14] - st1[:, 12:14], axis=1), np.hypot(0.008, 0.004) * 0.1, atol=1e-5)
    env.close()


def test_draw_refill_keeps_the_rng_sequence(oracle, edge_modes):
    """Many short episodes: draws are consumed in episode order across host refills of the device-side ring and
    across the standby pipeline (which always holds one pre-computed episode)."""
    n, max_steps, steps = 4, 2, 150
    env = _world(edge_modes, n, S=64, max_steps=max_steps, rng="host")     # the host-fed ring is what this test is about
    env.seed(7)
    env.reset()
    act = np.zeros((n, 2), dtype=np.float32)
    for k in range(steps):
        _, _, done, _ = env.step(act)
        assert done.all() == ((k + 1) % max_steps == 0)
    st = env.world.get_state()
    episode = steps // max_steps          # index of the episode every env is in now
    for i in range(n):
        rng = oracle.gym_np_random(7 + i)
        for _ in range(episode + 1):
            embed = rng.uniform(0.0015, 0.0065); ang = rng.uniform(-np.pi, np.pi)
        assert st[i, 19] == embed and st[i, 20] == ang
    assert not env.world.pipeline_error()
    env.close()


@pytest.mark.parametrize("arm,sensor", [("ur5", "digit"), ("ur5", "digitac"), ("mg400", "tactip"), ("mg400", "digitac")])
def test_other_arms_and_sensors_match_oracle(oracle, edge_modes, arm, sensor):
    """edge_follow-v0 on the other arm (MG400: 8-joint tree, pseudo-inverse velocity control with slaved joints,
    mg400.py:77-129) and the other sensors (DIGIT / DigiTac: 40 deg camera, no border)."""
    import tactile_gym_b200 as tg

    modes = dict(edge_modes, arm_type=arm, tactile_sensor_name=sensor)
    n, S, steps = 8, 128, 5
    env = tg.make_vec("edge_follow-v0", n, env_kwargs={"env_modes": modes, "image_size": [S, S], "max_steps": 200})
    rng = np.random.RandomState(3)
    lo, hi = {"tactip": (0.0015, 0.0065), "digit": (0.0011, 0.0028), "digitac": (0.0015, 0.0045)}[sensor]
    draws = np.stack([rng.uniform(lo, hi, (n, 2)), rng.uniform(-np.pi, np.pi, (n, 2))], axis=2)
    env.world.set_draws(draws)
    obs = env.reset()["tactile"]
    st = env.world.get_state()
    nb = env.world.nb
    refs = []
    for i in range(n):
        r = oracle.EdgeFollowOracle(image_size=S, arm=arm, sensor=sensor)
        r.reset(draws=tuple(draws[i, 0]))
        refs.append(r)
        assert np.allclose(st[i, :nb], np.array(r.s.q[:nb]), atol=5e-6)
        for k in range(nb):
            r.s.q[k] = st[i, k]; r.s.qd[k] = st[i, nb + k]
        mx, frac = _img_close(r.observation(), obs[i])
        assert mx <= 1 and frac < 1e-3, (i, mx, frac)
        assert (obs[i] > 0).sum() > 20
    for k in range(steps):
        act = rng.uniform(-0.25, 0.25, (n, 2)).astype(np.float32)
        for i, r in enumerate(refs):
            for j in range(nb):
                r.s.q[j] = st[i, j]; r.s.qd[j] = st[i, nb + j]
        o2, rew, done, _ = env.step(act)
        st = env.world.get_state()
        for i, r in enumerate(refs):
            o, rr, dd, _ = r.step(act[i])
            assert np.allclose(st[i, :nb], np.array(r.s.q[:nb]), atol=1e-9)
            assert np.allclose(st[i, nb:2 * nb], np.array(r.s.qd[:nb]), atol=1e-8)
            assert np.allclose(st[i, 2 * nb:2 * nb + 3], r.tcp_world()[0], atol=1e-8)
            assert abs(rr - rew[i]) < 1e-6 and bool(dd) == bool(done[i])
            mx, frac = _img_close(o, o2["tactile"][i])
            assert mx <= 1 and frac < 1e-3, (k, i, mx, frac)
    env.close()


BALANCE_MODES = {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "object_mode": "pole", "rand_gravity": True,
                 "rand_embed_dist": True, "observation_mode": "tactile", "reward_mode": "dense", "arm_type": "ur5",
                 "tactile_sensor_name": "tactip"}


def _sync_balance(ref, row, nb=6):
    for k in range(nb):
        ref.s.q[k] = row[k]; ref.s.qd[k] = row[nb + k]
    o = row[2 * nb + 11:]
    for c in range(3):
        ref.o.pos[c] = o[c]; ref.o.vel[c] = o[7 + c]; ref.o.omg[c] = o[10 + c]
    for c in range(4):
        ref.o.quat[c] = o[3 + c]
    ref.steps = int(row[2 * nb + 9])


@pytest.mark.parametrize("S,movement", [(128, "xy"), (256, "xyRxRy")])
def test_object_balance_matches_oracle(oracle, S, movement):
    """object_balance-v0 (BASELINE config 5 at S = 256): free pole on a point-to-point constraint, per-episode gravity,
    one-step random push, fall termination; each step compared from an identical state."""
    import tactile_gym_b200 as tg

    modes = dict(BALANCE_MODES, movement_mode=movement)
    n = 6
    env = tg.make_vec("object_balance-v0", n, env_kwargs={"env_modes": modes, "image_size": [S, S], "max_steps": 250})
    rng = np.random.RandomState(S)
    draws = np.stack([rng.uniform(-1.0, -0.1, (n, 2)), rng.uniform(0.003, 0.006, (n, 2)),
                      rng.choice([-1, 1], (n, 2)) * rng.rand(n, 2), rng.choice([-1, 1], (n, 2)) * rng.rand(n, 2)], axis=2)
    env.world.set_draws(draws)
    obs = env.reset()["tactile"]
    st = env.world.get_state()
    refs = []
    for i in range(n):
        r = oracle.ObjectBalanceOracle(image_size=S, movement_mode=movement)
        r.reset(draws=draws[i, 0])
        refs.append(r)
        assert np.allclose(st[i, :6], np.array(r.s.q[:6]), atol=2e-6)
        assert np.allclose(st[i, 23:26], np.array(r.o.pos[:]), atol=1e-12) and np.allclose(st[i, 26:30], np.array(r.o.quat[:]), atol=1e-12)
        assert st[i, 36] == draws[i, 0, 0]
        _sync_balance(r, st[i])
        mx, frac = _img_close(r.observation(), obs[i])
        assert mx <= 1 and frac < 1e-3, (i, mx, frac)
        assert (obs[i][..., 0][r.ref[2] == 0] > 0).sum() > 100     # the base plate presses into the skin
    ever_done = np.zeros(n, dtype=bool)
    act_dim = env.world.act_dim
    for k in range(45):
        act = rng.uniform(-0.25, 0.25, (n, act_dim)).astype(np.float32) * (0.0 if k < 25 else 1.0)
        o2, rew, done, infos = env.step(act)
        st = env.world.get_state()
        for i, r in enumerate(refs):
            if ever_done[i]:
                continue
            o, rr, dd, _ = r.step(act[i])
            assert rr == rew[i] == 1.0 and bool(dd) == bool(done[i]), (k, i)
            if dd:
                ever_done[i] = True       # the env auto-reset; the terminal observation is checked instead
                mx, frac = _img_close(o, infos[i]["terminal_observation"]["tactile"])
                assert mx <= 1 and frac < 2e-3, (k, i, mx, frac)
                continue
            tol = 5e-6 if k == 0 else 1e-9  # step 0 starts from the (noisy) reset state on both sides
            assert np.allclose(st[i, :6], np.array(r.s.q[:6]), atol=tol)
            assert np.allclose(st[i, 23:26], np.array(r.o.pos[:]), atol=tol)
            assert np.allclose(st[i, 26:30], np.array(r.o.quat[:]), atol=tol * 10)
            assert np.allclose(st[i, 30:36], np.array(list(r.o.vel[:]) + list(r.o.omg[:])), atol=max(tol, 1e-8) * 100)
            _sync_balance(r, st[i])
            mx, frac = _img_close(r.observation(), o2["tactile"][i])
            assert mx <= 1 and frac < 1e-3, (k, i, mx, frac)
    assert ever_done.any()                # with no control for 25 steps some poles fall past 35 degrees
    assert not env.world.pipeline_error()
    env.close()


SURFACE_MODES = {"movement_mode": "xyzRxRy", "control_mode": "TCP_velocity_control", "noise_mode": "simplex",
                 "observation_mode": "tactile", "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "digit"}


@pytest.mark.parametrize("sensor,S,movement", [("digit", 128, "xyzRxRy"), ("tactip", 128, "xyz"), ("tactip", 64, "xyzRxRy"), ("digitac", 256, "xyzRxRy")])
def test_surface_follow_matches_oracle(oracle, sensor, S, movement):
    """surface_follow-v0 (BASELINE config 3 = digit, 128): per-env OpenSimplex heightfield generated on the device,
    heightfield raster with per-tile primitive lists, constant drive towards the goal, surface-distance + normal reward;
    each step compared from an identical state.  (digitac 256 is one of the reference's stale fixtures: the garbage is
    reproduced on both sides)"""
    import tactile_gym_b200 as tg

    modes = dict(SURFACE_MODES, movement_mode=movement, tactile_sensor_name=sensor)
    n = 5
    env = tg.make_vec("surface_follow-v0", n, env_kwargs={"env_modes": modes, "image_size": [S, S], "max_steps": 200})
    rng = np.random.RandomState(S + len(sensor))
    draws = np.stack([rng.randint(0, 10 ** 8, (n, 2)).astype(np.float64), rng.uniform(-np.pi, np.pi, (n, 2))], axis=2)
    env.world.set_draws(draws)
    obs = env.reset()["tactile"]
    st = env.world.get_state()
    refs = []
    for i in range(n):
        r = oracle.SurfaceFollowOracle(image_size=S, sensor=sensor, movement_mode=movement)
        r.reset(draws=(draws[i, 0, 0], draws[i, 0, 1]))
        refs.append(r)
        assert np.allclose(st[i, :6], np.array(r.s.q[:6]), atol=2e-6), i
        assert st[i, 22] == r.last_reset_substeps
        _sync_oracle(r, st[i])
        mx, frac = _img_close(r.observation(), obs[i])
        assert mx <= 1 and frac < 1e-3, (i, mx, frac)
    act_dim = env.world.act_dim
    touched = 0
    for k in range(30):
        act = rng.uniform(-0.25, 0.25, (n, act_dim)).astype(np.float32)
        act[:, 0] = 0.25 if k < 12 else act[:, 0]          # push down (workframe z = world -z) so the skin meets the surface
        o2, rew, done, infos = env.step(act)
        st = env.world.get_state()
        for i, r in enumerate(refs):
            o, rr, dd, _ = r.step(act[i])
            tol = 5e-6 if k == 0 else 1e-9
            assert np.allclose(st[i, :6], np.array(r.s.q[:6]), atol=tol), (k, i)
            assert not dd and not done[i]
            _sync_oracle(r, st[i])
            rr, _ = r.step_data()
            assert abs(rr - rew[i]) < 1e-6 * max(1.0, abs(rr)), (k, i, rr, rew[i])
            img = r.observation()
            mx, frac = _img_close(img, o2["tactile"][i])
            assert mx <= 1 and frac < 1e-3, (k, i, mx, frac)
            touched += int((img[..., 0][r.ref[2] == 0] > 0).sum() > 50)
    assert touched > 10                                       # the comparison is not vacuous: the surface shows in the images
    assert not env.world.pipeline_error()
    env.close()


def test_surface_follow_episode_turnover(oracle):
    """short episodes: the next heightfield is built in the other buffer while the live one is rendered, then swapped"""
    import tactile_gym_b200 as tg

    n, S, L = 4, 64, 150
    env = tg.make_vec("surface_follow-v0", n, env_kwargs={"env_modes": SURFACE_MODES, "image_size": [S, S], "max_steps": L})
    rng = np.random.RandomState(9)
    draws = np.stack([rng.randint(0, 10 ** 8, (n, 3)).astype(np.float64), rng.uniform(-np.pi, np.pi, (n, 3))], axis=2)
    env.world.set_draws(draws)
    env.reset()
    act = np.zeros((n, 3), dtype=np.float32); act[:, 0] = 0.1
    for k in range(L):
        obs, rew, done, infos = env.step(act)
    assert done.all()
    st = env.world.get_state()
    for i in range(n):
        r = oracle.SurfaceFollowOracle(image_size=S, sensor="digit")
        o = r.reset(draws=(draws[i, 1, 0], draws[i, 1, 1]))
        assert np.allclose(st[i, :6], np.array(r.s.q[:6]), atol=2e-6)
        _sync_oracle(r, st[i])
        assert _img_close(r.observation(), obs["tactile"][i])[0] <= 1
        assert infos[i]["terminal_observation"]["tactile"].shape == (S, S, 1)
    assert env.world.pipeline_stalls() == 0
    env.close()


@pytest.mark.parametrize("sensor,S,movement", [("tactip", 64, "xyzRxRy"), ("digit", 128, "xyz")])
def test_surface_follow_goal_matches_oracle(oracle, sensor, S, movement):
    """surface_follow-v1 (SurfaceFollowGoalEnv): the policy steers x / y itself, reward = goal distance + 10 x surface distance
    + normal alignment, extended_feature = TCP and goal position in the work frame; each step from an identical state."""
    import tactile_gym_b200 as tg

    modes = dict(SURFACE_MODES, movement_mode=movement, tactile_sensor_name=sensor, observation_mode="tactile_and_feature")
    n = 4
    env = tg.make_vec("surface_follow-v1", n, env_kwargs={"env_modes": modes, "image_size": [S, S], "max_steps": 200})
    rng = np.random.RandomState(S)
    draws = np.stack([rng.randint(0, 10 ** 8, (n, 2)).astype(np.float64), rng.uniform(-np.pi, np.pi, (n, 2))], axis=2)
    env.world.set_draws(draws)
    ob = env.reset()
    st = env.world.get_state()
    refs = []
    for i in range(n):
        r = oracle.SurfaceFollowOracle(image_size=S, sensor=sensor, movement_mode=movement, variant="goal")
        r.reset(draws=(draws[i, 0, 0], draws[i, 0, 1]))
        refs.append(r)
        _sync_oracle(r, st[i])
        assert ob["extended_feature"].shape == (n, 6)
        assert np.allclose(r.features(), ob["extended_feature"][i], atol=1e-6)
        assert abs(np.linalg.norm(ob["extended_feature"][i][3:5]) - 0.15) < 1e-6      # the goal is 0.15 m from the work origin
    act_dim = env.world.act_dim
    assert act_dim == (5 if movement == "xyzRxRy" else 3)
    for k in range(12):
        act = rng.uniform(-0.25, 0.25, (n, act_dim)).astype(np.float32)
        act[:, 2] = 0.25                                       # push down so the skin meets the surface
        o2, rew, done, infos = env.step(act)
        st = env.world.get_state()
        for i, r in enumerate(refs):
            o, rr, dd, _ = r.step(act[i])
            tol = 5e-6 if k == 0 else 1e-9
            assert np.allclose(st[i, :6], np.array(r.s.q[:6]), atol=tol), (k, i)
            assert not dd and not done[i]
            _sync_oracle(r, st[i])
            rr, _ = r.step_data()
            assert abs(rr - rew[i]) < 1e-6 * max(1.0, abs(rr)), (k, i, rr, rew[i])
            assert np.allclose(r.features(), o2["extended_feature"][i], atol=1e-6), (k, i)
            mx, frac = _img_close(r.observation(), o2["tactile"][i])
            assert mx <= 1 and frac < 1e-3, (k, i, mx, frac)
    assert not env.world.pipeline_error()
    env.close()
