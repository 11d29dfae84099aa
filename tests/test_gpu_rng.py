"""The device RNG on the GPU (csrc/tg_rng.cuh: every env carries numpy's MT19937 state and its resets draw from it with numpy's
call semantics): (1) an env seeded on the device runs through exactly the episodes the host-drawn ring gives it - same states,
bit for bit, across many episode turnovers; (2) surface_follow's noise_mode "random" (1,024 uniform draws per reset,
base_surface_env.py:302-318 - only possible with the generator on the device) against the CPU oracle seeded the same way."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [
    ("edge_follow-v0", {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "noise_mode": "rand_height", "observation_mode": "tactile",
                        "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "tactip"}, 2, 3),
    ("object_balance-v0", {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "object_mode": "pole", "rand_gravity": True,
                           "rand_embed_dist": True, "observation_mode": "tactile", "reward_mode": "dense", "arm_type": "ur5",
                           "tactile_sensor_name": "tactip"}, 2, 3),
    ("surface_follow-v0", {"movement_mode": "xyzRxRy", "control_mode": "TCP_velocity_control", "noise_mode": "simplex", "observation_mode": "tactile",
                           "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "digit"}, 3, 4),
    ("object_push-v0", {"movement_mode": "TyRz", "control_mode": "TCP_velocity_control", "rand_init_orn": True, "rand_obj_mass": True,
                        "traj_type": "simplex", "observation_mode": "tactile_and_feature", "reward_mode": "dense", "arm_type": "mg400",
                        "tactile_sensor_name": "digitac"}, 2, 3),
    ("object_roll-v0", {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "rand_obj_size": True, "rand_embed_dist": True,
                        "rand_init_obj_pos": True, "observation_mode": "tactile", "reward_mode": "dense", "arm_type": "ur5",
                        "tactile_sensor_name": "tactip"}, 2, 3),
]


@pytest.mark.parametrize("env_id,modes,act_dim,max_steps", CASES)
def test_device_rng_reproduces_the_host_drawn_episodes(env_id, modes, act_dim, max_steps):
    import tactile_gym_b200 as tg

    n, S, steps = 9, 64, 40
    runs = {}
    for rng_mode in ("host", "device"):
        env = tg.make_vec(env_id, n, seed=21, env_kwargs={"env_modes": modes, "image_size": [S, S], "max_steps": max_steps}, rng=rng_mode)
        assert env.world.rng_mode == rng_mode
        env.reset()
        rs = np.random.RandomState(1)
        rec = [env.world.get_state()]
        for k in range(steps):
            o, r, d, _ = env.step(rs.uniform(-0.25, 0.25, (n, act_dim)).astype(np.float32))
            rec.append(env.world.get_state())
        assert not env.world.pipeline_error() and not env.world.draws_exhausted()
        runs[rng_mode] = rec
        env.close()
    for k, (a, b) in enumerate(zip(runs["host"], runs["device"])):
        assert np.array_equal(a, b), (k, np.abs(a - b).max())
    # the runs went through many episodes, each with its own draws
    assert len({tuple(s[0]) for s in runs["device"]}) > steps // max_steps


@pytest.mark.parametrize("movement,variant_id", [("xyzRxRy", "surface_follow-v0"), ("yz", "surface_follow-v1")])
def test_random_noise_surface_matches_oracle(oracle, movement, variant_id):
    import tactile_gym_b200 as tg

    modes = {"movement_mode": movement, "control_mode": "TCP_velocity_control", "noise_mode": "random", "observation_mode": "tactile",
             "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "tactip"}
    n, S = 5, 64
    env = tg.make_vec(variant_id, n, seed=50, env_kwargs={"env_modes": modes, "image_size": [S, S], "max_steps": 6})
    assert env.world.rng_mode == "device"
    obs = env.reset()["tactile"]
    st = env.world.get_state()
    variant = "goal" if variant_id.endswith("v1") else "auto"
    refs = [oracle.SurfaceFollowOracle(image_size=S, sensor="tactip", movement_mode=movement, noise_mode="random", variant=variant, max_steps=6,
                                       seed=50 + i) for i in range(n)]

    def img_close(a, b):
        d = np.abs(a.astype(np.int32) - b.astype(np.int32))
        return d.max(), (d != 0).mean()

    def sync(r, row):
        for k in range(6):
            r.s.q[k] = row[k]; r.s.qd[k] = row[6 + k]
        r.steps = int(row[21])

    for i, r in enumerate(refs):
        r.reset()
        assert 0.0 < r.h.max() <= 0.005 and len(np.unique(r.h)) > 900
        assert np.allclose(st[i, :6], np.array(r.s.q[:6]), atol=2e-6), i
        assert st[i, 22] == r.last_reset_substeps
        sync(r, st[i])
        mx, frac = img_close(r.observation(), obs[i])
        assert mx <= 1 and frac < 2e-3, (i, mx, frac)
    act_dim = env.world.act_dim
    rs = np.random.RandomState(2)
    episode = 0
    for k in range(14):                                      # through two episode ends: the next surfaces are drawn on the device too
        act = rs.uniform(-0.25, 0.25, (n, act_dim)).astype(np.float32)
        act[:, 0 if variant == "auto" else 1] = 0.25          # press towards the surface
        o2, rew, done, infos = env.step(act)
        st2 = env.world.get_state()
        for i, r in enumerate(refs):
            sync(r, st[i])
            _, rr, dd, _ = r.step(act[i])
            assert bool(dd) == bool(done[i]), (k, i)
            assert abs(rr - rew[i]) < 1e-6 * max(1.0, abs(rr)), (k, i, rr, rew[i])
            if dd:
                r.reset()                                    # the oracle's generator moves on exactly like the env's
                assert np.allclose(st2[i, :6], np.array(r.s.q[:6]), atol=2e-6), (k, i)
            else:
                assert np.allclose(st2[i, :6], np.array(r.s.q[:6]), atol=1e-9), (k, i)
            sync(r, st2[i])
            mx, frac = img_close(r.observation(), o2["tactile"][i])
            assert mx <= 1 and frac < 2e-3, (k, i, mx, frac)
        episode += int(done.all())
        st = st2
    assert episode == 2 and not env.world.pipeline_error()
    env.close()
