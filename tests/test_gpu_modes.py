"""The env_modes beyond the BASELINE configs, CUDA path (through the C ABI) against the CPU oracle:
reward_mode "sparse" of edge_follow (edge_follow_env.py:430-438), object_balance (object_balance_env.py:508-518) and
surface_follow (the accumulated dense reward paid out at the goal, surface_follow_auto_env.py:59-73); surface_follow's 1-d
surfaces for the yz / yzRx movement modes (base_surface_env.py:339-357, 512-514) and noise_mode "none" (:436-437).
Observations are taken in "oracle" mode where the image is not the point (the raster has its own parity tests)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BASE = {"control_mode": "TCP_velocity_control", "observation_mode": "oracle", "reward_mode": "sparse", "arm_type": "ur5", "tactile_sensor_name": "tactip"}


def _sync(ref, row, nb=6, with_obj=False):
    for k in range(nb):
        ref.s.q[k] = row[k]; ref.s.qd[k] = row[nb + k]
    ref.steps = int(row[2 * nb + 9])
    if with_obj:
        o = row[2 * nb + 11:]
        for c in range(3):
            ref.o.pos[c] = o[c]; ref.o.vel[c] = o[7 + c]; ref.o.omg[c] = o[10 + c]
        for c in range(4):
            ref.o.quat[c] = o[3 + c]


def test_edge_follow_sparse_reward(oracle):
    """drive along the edge towards the goal: reward 0 until the TCP is within termination_dist, then 1 and done"""
    import tactile_gym_b200 as tg

    n = 4
    modes = dict(BASE, movement_mode="xy", noise_mode="rand_height")
    env = tg.make_vec("edge_follow-v0", n, env_kwargs={"env_modes": modes, "image_size": [64, 64], "max_steps": 400})
    rng = np.random.RandomState(2)
    draws = np.stack([rng.uniform(0.0015, 0.0065, (n, 2)), rng.uniform(-np.pi, np.pi, (n, 2))], axis=2)
    env.world.set_draws(draws)
    obs = env.reset()["oracle"]
    refs = [oracle.EdgeFollowOracle(image_size=64, max_steps=400, reward_mode="sparse") for _ in range(n)]
    for i, r in enumerate(refs):
        r.reset(draws=tuple(draws[i, 0]))
    paid = np.zeros(n, bool)
    for k in range(260):
        # oracle observation: tcp pos (3), lin vel (3), goal pos (3) in the work frame -> head for the goal at full speed
        d = obs[:, 6:8] - obs[:, 0:2]
        act = (0.25 * d / np.maximum(np.abs(d).max(axis=1, keepdims=True), 1e-9)).astype(np.float32)
        st = env.world.get_state()
        want = []
        for i, r in enumerate(refs):
            _sync(r, st[i])
            want.append(r.step(act[i])[1:3])
        o, rew, done, infos = env.step(act)
        obs = o["oracle"]
        for i in range(n):
            if paid[i]:
                continue
            assert rew[i] == want[i][0] and bool(done[i]) == want[i][1], (k, i, rew[i], want[i])
            if done[i]:
                assert rew[i] == 1.0
                paid[i] = True
        if paid.all():
            break
    assert paid.all() and k > 100          # 0.175 m at 1 mm per step
    env.close()


def test_object_balance_sparse_reward(oracle):
    """tilt the plate hard: 0 while the pole stands, -1 and done when it has fallen"""
    import tactile_gym_b200 as tg

    n = 4
    modes = dict(BASE, movement_mode="xyRxRy", object_mode="pole", rand_gravity=False, rand_embed_dist=False)
    env = tg.make_vec("object_balance-v0", n, env_kwargs={"env_modes": modes, "image_size": [64, 64], "max_steps": 250})
    rng = np.random.RandomState(3)
    draws = np.stack([np.full((n, 2), -1.0), np.full((n, 2), 0.0035), rng.choice([-1, 1], (n, 2)) * rng.rand(n, 2),
                      rng.choice([-1, 1], (n, 2)) * rng.rand(n, 2)], axis=2)
    env.world.set_draws(draws)
    env.reset()
    refs = []
    for i in range(n):
        r = oracle.ObjectBalanceOracle(image_size=64, movement_mode="xyRxRy", rand_gravity=False, rand_embed_dist=False)
        r.reward_mode = "sparse"
        r.reset(draws=draws[i, 0])
        refs.append(r)
    fell = np.zeros(n, bool)
    for k in range(250):
        act = np.tile(np.array([0.25, 0.25, 0.25, -0.25], np.float32), (n, 1))
        st = env.world.get_state()
        want = []
        for i, r in enumerate(refs):
            _sync(r, st[i], with_obj=True)
            rr, dd = r.step(act[i])[1:3] if not fell[i] else (0.0, False)
            want.append((rr, dd))
        o, rew, done, infos = env.step(act)
        for i in range(n):
            if fell[i]:
                continue
            assert rew[i] == want[i][0] and bool(done[i]) == want[i][1], (k, i, rew[i], want[i])
            if done[i] and rew[i] == -1.0:
                fell[i] = True
        if fell.all():
            break
    assert fell.all()
    env.close()


@pytest.mark.parametrize("env_id,movement,noise", [("surface_follow-v0", "xyzRxRy", "simplex"), ("surface_follow-v0", "yzRx", "simplex"),
                                                   ("surface_follow-v1", "yz", "simplex"), ("surface_follow-v0", "xyz", "none"),
                                                   ("surface_follow-v2", "xRz", "simplex")])
def test_surface_follow_modes_and_sparse_reward(oracle, env_id, movement, noise):
    """1-d / flat surfaces and the (0, +-1) goal direction show in the reset pose, the goal and the per-step reward terms; the
    sparse reward is 0 on the way and the accumulated dense reward (reset's own get_step_data included) at the goal"""
    import tactile_gym_b200 as tg

    n = 4
    variant = {"v0": "auto", "v1": "goal", "v2": "vert"}[env_id[-2:]]
    modes = dict(BASE, movement_mode=movement, noise_mode=noise)
    env = tg.make_vec(env_id, n, env_kwargs={"env_modes": modes, "image_size": [64, 64], "max_steps": 400})
    rng = np.random.RandomState(len(movement) + len(noise))
    one_d = movement in ("yz", "yzRx", "xRz")
    second = rng.choice([-1.0, 1.0], (n, 2)) if one_d else rng.uniform(-np.pi, np.pi, (n, 2))
    first = rng.randint(0, 10 ** 8, (n, 2)).astype(np.float64) if noise == "simplex" else np.zeros((n, 2))
    draws = np.stack([first, second], axis=2)
    env.world.set_draws(draws)
    obs = env.reset()["oracle"]
    st = env.world.get_state()
    refs = []
    for i in range(n):
        r = oracle.SurfaceFollowOracle(image_size=64, sensor="tactip", max_steps=400, movement_mode=movement, variant=variant,
                                       noise_mode=noise, reward_mode="sparse", render=False)
        r.reset(draws=(draws[i, 0, 0], draws[i, 0, 1]))
        refs.append(r)
        assert np.allclose(st[i, :6], np.array(r.s.q[:6]), atol=2e-6), i          # start pose = centre height of the new surface
        assert st[i, 22] == r.last_reset_substeps
        if noise == "none" or movement == "xRz":
            assert np.all(r.h == 0)
        elif one_d:
            assert np.ptp(r.h, axis=0).max() == 0 and np.ptp(r.h) > 1e-3          # constant along x, varying along y
        _sync(r, st[i]); r.accum_rew = 0.0; r.step_data()
        assert np.allclose(obs[i], r.oracle_obs(), atol=2e-5), (i, np.abs(obs[i] - r.oracle_obs()).max())
    act_dim = env.world.act_dim
    paid = np.zeros(n, bool)
    zi = {"auto": 0, "goal": 1 if one_d else 2, "vert": None}[variant]      # which policy action is z (-v2 has none)
    for k in range(330):
        # steer: z towards the goal's height (oracle obs: tcp pos [0:3], goal pos [13:16], work frame); v1 also steers x / y
        act = np.zeros((n, act_dim), np.float32)
        dz = obs[:, 15] - obs[:, 2]
        if zi is not None:
            act[:, zi] = np.clip(250.0 * dz * 0.1, -0.25, 0.25)
        else:
            act[:, 0] = np.clip(25.0 * (obs[:, 13] - obs[:, 0]), -0.25, 0.25)      # -v2: keep x on the goal's, Rz has no range
        if variant == "goal":
            d = obs[:, 13:15] - obs[:, 0:2]
            a = 0.25 * d / np.maximum(np.abs(d).max(axis=1, keepdims=True), 1e-9)
            if one_d:
                act[:, 0] = a[:, 1]
            else:
                act[:, 0], act[:, 1] = a[:, 0], a[:, 1]
        st = env.world.get_state()
        want = []
        for i, r in enumerate(refs):
            _sync(r, st[i])
            want.append(r.step(act[i])[1:3])
        o, rew, done, infos = env.step(act)
        obs = o["oracle"]
        for i in range(n):
            if paid[i]:
                continue
            assert bool(done[i]) == want[i][1], (k, i)
            assert abs(rew[i] - want[i][0]) < 1e-5 * max(1.0, abs(want[i][0])), (k, i, rew[i], want[i][0])
            if done[i]:
                assert rew[i] < -1e-3 and k > 100          # the accumulated (negative) dense reward of ~140 steps
                paid[i] = True
        if paid.all():
            break
    assert paid.all()
    assert not env.world.pipeline_error()
    env.close()


@pytest.mark.parametrize("task", ["edge", "surface"])
def test_tcp_position_control(oracle, task):
    """control_mode "TCP_position_control" (robot.py:156-186, base_robot_arm.py:228-279): pose-delta actions, IK from the current
    joints, position motors, blocking move of <= 10 substeps with the pose / speed exit - each step from an identical state"""
    import tactile_gym_b200 as tg

    n = 5
    rng = np.random.RandomState(17)
    if task == "edge":
        modes = dict(BASE, movement_mode="xyzRz", noise_mode="rand_height", control_mode="TCP_position_control", reward_mode="dense")
        env_id = "edge_follow-v0"
        draws = np.stack([rng.uniform(0.0015, 0.0065, (n, 2)), rng.uniform(-np.pi, np.pi, (n, 2))], axis=2)
        mk = lambda: oracle.EdgeFollowOracle(image_size=64, movement_mode="xyzRz", control_mode="TCP_position_control")
    else:
        modes = dict(BASE, movement_mode="xyzRxRy", noise_mode="simplex", control_mode="TCP_position_control", reward_mode="dense")
        env_id = "surface_follow-v0"
        draws = np.stack([rng.randint(0, 10 ** 8, (n, 2)).astype(np.float64), rng.uniform(-np.pi, np.pi, (n, 2))], axis=2)
        mk = lambda: oracle.SurfaceFollowOracle(image_size=64, sensor="tactip", control_mode="TCP_position_control", render=False)
    env = tg.make_vec(env_id, n, env_kwargs={"env_modes": modes, "image_size": [64, 64], "max_steps": 200})
    env.world.set_draws(draws)
    obs = env.reset()["oracle"]
    st = env.world.get_state()
    refs = []
    for i in range(n):
        r = mk()
        r.reset(draws=tuple(draws[i, 0]))
        refs.append(r)
    p0 = obs[:, 0:3].copy()
    moved = 0.0
    for k in range(12):
        act = rng.uniform(-0.25, 0.25, (n, env.world.act_dim)).astype(np.float32)
        if k < 6:
            act[:, 0] = 0.25                       # a steady 1 mm per step along the first action axis
        for i, r in enumerate(refs):
            _sync(r, st[i])
            r.step(act[i])
        o, rew, done, infos = env.step(act)
        st = env.world.get_state()
        for i, r in enumerate(refs):
            assert np.allclose(st[i, :6], np.array(r.s.q[:6]), atol=1e-9), (k, i, np.abs(st[i, :6] - np.array(r.s.q[:6])).max())
            assert np.allclose(st[i, 6:12], np.array(r.s.qd[:6]), atol=1e-7), (k, i)
            assert 1 <= r.last_move_substeps <= 10
            assert abs(rew[i] - r.reward) < 1e-6 and bool(done[i]) == r.done
            assert np.allclose(o["oracle"][i], r.oracle_obs(), atol=2e-5)
        if k == 5:
            axis = 0 if task == "edge" else 2      # edge xyzRz: x; surface-auto xyzRxRy: the first policy action is z
            moved = np.abs(o["oracle"][:, axis] - p0[:, axis])
    assert np.all(moved > 0.004) and np.all(moved < 0.0065), moved      # ~6 mm after six full-scale position steps
    env.close()
