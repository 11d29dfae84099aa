"""observation_mode "oracle" (get_oracle_obs: edge_follow_env.py:454-476, base_surface_env.py:789-819, object_balance_env.py:528-563,
object_push_env.py:571-609, object_roll_env.py:371-409): the state vector the CUDA path writes (tg_bind_oracle_obs, float32)
against the CPU oracle's restatement, every step from an identical state.  Tolerance 2e-5: float32 storage of values up to ~3
plus the 1e-9 / 1e-6 per-step state differences the parity tests establish."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 2e-5
BASE = {"control_mode": "TCP_velocity_control", "observation_mode": "oracle", "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "tactip"}


def _sync(ref, row, nb, with_obj):
    for k in range(nb):
        ref.s.q[k] = row[k]; ref.s.qd[k] = row[nb + k]
    ref.steps = int(row[2 * nb + 9])
    if with_obj:
        o = row[2 * nb + 11:]
        for c in range(3):
            ref.o.pos[c] = o[c]; ref.o.vel[c] = o[7 + c]; ref.o.omg[c] = o[10 + c]
        for c in range(4):
            ref.o.quat[c] = o[3 + c]


def _case(oracle, task, rng, n):
    if task == "edge":
        modes = dict(BASE, movement_mode="xyRz", noise_mode="rand_height")
        draws = np.stack([rng.uniform(0.0015, 0.0065, (n, 2)), rng.uniform(-np.pi, np.pi, (n, 2))], axis=2)
        mk = lambda: oracle.EdgeFollowOracle(image_size=64, movement_mode="xyRz")
        return "edge_follow-v0", modes, draws, mk, 10, False, 200
    if task == "surface":
        modes = dict(BASE, movement_mode="xyzRxRy", noise_mode="simplex", tactile_sensor_name="digit")
        draws = np.stack([rng.randint(0, 10 ** 8, (n, 2)).astype(np.float64), rng.uniform(-np.pi, np.pi, (n, 2))], axis=2)
        mk = lambda: oracle.SurfaceFollowOracle(image_size=64, sensor="digit")
        return "surface_follow-v0", modes, draws, mk, 20, False, 200
    if task == "balance":
        modes = dict(BASE, movement_mode="xyRxRy", object_mode="pole", rand_gravity=True, rand_embed_dist=True)
        draws = np.stack([rng.uniform(-1.0, -0.1, (n, 2)), rng.uniform(0.003, 0.006, (n, 2)),
                          rng.choice([-1, 1], (n, 2)) * rng.rand(n, 2), rng.choice([-1, 1], (n, 2)) * rng.rand(n, 2)], axis=2)
        mk = lambda: oracle.ObjectBalanceOracle(image_size=64, movement_mode="xyRxRy")
        return "object_balance-v0", modes, draws, mk, 26, True, 250
    if task == "push":
        modes = dict(BASE, movement_mode="TyRz", rand_init_orn=True, rand_obj_mass=True, traj_type="simplex")
        draws = np.stack([rng.uniform(-np.pi / 32, np.pi / 32, (n, 2)), rng.uniform(0.4, 0.8, (n, 2)),
                          rng.randint(0, 10 ** 8, (n, 2)).astype(np.float64)], axis=2)
        mk = lambda: oracle.ObjectPushOracle(image_size=64, arm="ur5", sensor="tactip", movement_mode="TyRz", traj_type="simplex",
                                             rand_init_orn=True, rand_obj_mass=True)
        return "object_push-v0", modes, draws, mk, 30, True, 1000
    modes = dict(BASE, movement_mode="xy", rand_init_obj_pos=True, rand_obj_size=True, rand_embed_dist=True)
    draws = np.stack([rng.uniform(1.0, 2.0, (n, 2)), rng.uniform(0.0019, 0.003, (n, 2)), rng.uniform(-0.009, 0.009, (n, 2)),
                      rng.uniform(-0.009, 0.009, (n, 2)), rng.uniform(-np.pi, np.pi, (n, 2)), rng.uniform(0.0, 0.015, (n, 2))], axis=2)
    mk = lambda: oracle.ObjectRollOracle(image_size=64, rand_obj_size=True, rand_embed_dist=True, rand_init_obj_pos=True)
    return "object_roll-v0", modes, draws, mk, 34, True, 250


@pytest.mark.parametrize("task", ["edge", "surface", "balance", "push", "roll"])
def test_oracle_observation_matches_oracle(oracle, task):
    import tactile_gym_b200 as tg

    n, nb = 5, 6
    rng = np.random.RandomState(len(task))
    env_id, modes, draws, mk, k_obs, with_obj, max_steps = _case(oracle, task, rng, n)
    env = tg.make_vec(env_id, n, env_kwargs={"env_modes": modes, "image_size": [64, 64], "max_steps": max_steps})
    assert list(env.observation_space.spaces) == ["oracle"] and env.observation_space["oracle"].shape == (k_obs,)
    env.world.set_draws(draws)
    obs = env.reset()
    assert set(obs) == {"oracle"} and obs["oracle"].shape == (n, k_obs) and obs["oracle"].dtype == np.float32
    l0 = env.world.launch_count()
    st = env.world.get_state()
    refs = []
    for i in range(n):
        r = mk()
        r.reset(draws=draws[i, 0] if draws.shape[2] > 2 else tuple(draws[i, 0]))
        refs.append(r)

    def check(o, tag):
        for i, r in enumerate(refs):
            if task == "push":      # the goal index is a rounding-level tie at reset (test_gpu_push.py): take the device's
                r.targ = int(st[i, 2 * nb + 25]) - 1
                r.update_goal()
            want = r.oracle_obs()
            assert want.shape == (k_obs,)
            assert np.allclose(o[i], want, atol=TOL), (tag, i, np.abs(o[i] - want).max(), int(np.abs(o[i] - want).argmax()))

    for i, r in enumerate(refs):
        _sync(r, st[i], nb, with_obj)
        if task == "surface":
            r.step_data()           # tip_i / tip_j follow the synced TCP
    check(obs["oracle"], "reset")
    for k in range(6):
        act = rng.uniform(-0.25, 0.25, (n, env.world.act_dim)).astype(np.float32)
        if task == "surface":
            act[:, 0] = 0.25
        for i, r in enumerate(refs):
            _sync(r, st[i], nb, with_obj)
            r.step(act[i])
        obs, rew, done, infos = env.step(act)
        st = env.world.get_state()
        assert not done.any()
        check(obs["oracle"], k)
        assert np.abs(obs["oracle"]).max() > 1e-3
    # nothing was rendered: one launch per step (the fused step kernel)
    assert env.world.launch_count() - l0 == 6
    env.close()


def test_oracle_terminal_observation_and_gym_env(oracle):
    """episode turnover: infos[i]["terminal_observation"]["oracle"] is the finished episode's last state, obs the new one's;
    the single-env gym surface returns the same vector"""
    import tactile_gym_b200 as tg

    n, S = 4, 64
    modes = dict(BASE, movement_mode="xy", noise_mode="rand_height")
    env = tg.make_vec("edge_follow-v0", n, env_kwargs={"env_modes": modes, "image_size": [S, S], "max_steps": 3})
    rng = np.random.RandomState(9)
    draws = np.stack([rng.uniform(0.0015, 0.0065, (n, 4)), rng.uniform(-np.pi, np.pi, (n, 4))], axis=2)
    env.world.set_draws(draws)
    env.reset()
    refs = []
    for i in range(n):
        r = oracle.EdgeFollowOracle(image_size=S, max_steps=3)
        r.reset(draws=tuple(draws[i, 0]))
        refs.append(r)
    for ep in range(2):
        for k in range(3):
            act = rng.uniform(-0.25, 0.25, (n, 2)).astype(np.float32)
            obs, rew, done, infos = env.step(act)
            for i, r in enumerate(refs):
                r.step(act[i])
        assert done.all()
        for i, r in enumerate(refs):
            assert np.allclose(infos[i]["terminal_observation"]["oracle"], r.oracle_obs(), atol=1e-4), (ep, i)
            r.reset(draws=tuple(draws[i, ep + 1]))
            assert np.allclose(obs["oracle"][i], r.oracle_obs(), atol=1e-4), (ep, i)
    env.close()

    e1 = tg.make("edge_follow-v0", max_steps=10, image_size=[S, S], env_modes=modes)
    o = e1.reset()
    assert set(o) == {"oracle"} and o["oracle"].shape == (10,)
    o2, r2, d2, _ = e1.step(np.array([0.25, 0.0], np.float32))
    assert abs(o2["oracle"][3] - 0.01) < 1e-3       # TCP x velocity in the work frame = the commanded 0.01 m/s
    assert np.allclose(e1.get_oracle_obs(), o2["oracle"])
    e1.close()
