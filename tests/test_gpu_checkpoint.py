"""Full checkpoint / resume (tg_checkpoint_save / tg_checkpoint_load) and the per-env NaN guard, on the GPU through
the C ABI.  Resume is checked the strict way: a run is checkpointed mid-way, continued, then restored and continued again with
the same actions - every observation byte, reward and done flag of the two continuations must be identical, across episode
turnovers (pre-computed next episodes, partial rebuilds, heightfield double buffers and RNG states are all in the checkpoint)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

EDGE = {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "noise_mode": "rand_height", "observation_mode": "tactile",
        "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "tactip"}
SURFACE = {"movement_mode": "xyzRxRy", "control_mode": "TCP_velocity_control", "noise_mode": "simplex", "observation_mode": "tactile",
           "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "digit"}
PUSH = {"movement_mode": "TyRz", "control_mode": "TCP_velocity_control", "rand_init_orn": True, "rand_obj_mass": True, "traj_type": "simplex",
        "observation_mode": "tactile_and_feature", "reward_mode": "dense", "arm_type": "mg400", "tactile_sensor_name": "digitac"}
BALANCE = {"movement_mode": "xyRxRy", "control_mode": "TCP_velocity_control", "object_mode": "pole", "rand_gravity": True, "rand_embed_dist": True,
           "observation_mode": "tactile", "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "tactip"}


def _run(env, acts):
    out = []
    for a in acts:
        o, r, d, infos = env.step(a)
        rec = {k: v.copy() for k, v in o.items()}
        rec["reward"], rec["done"] = r.copy(), d.copy()
        rec["term"] = [infos[i]["terminal_observation"]["tactile"].copy() for i in np.flatnonzero(d)]
        out.append(rec)
    return out


@pytest.mark.parametrize("env_id,modes,max_steps,rng", [("edge_follow-v0", EDGE, 7, "device"), ("surface_follow-v0", SURFACE, 9, "device"),
                                                        ("object_push-v0", PUSH, 6, "device"), ("object_balance-v0", BALANCE, 8, "device"),
                                                        ("edge_follow-v0", EDGE, 7, "host")])
def test_checkpoint_resume_is_bit_identical(env_id, modes, max_steps, rng):
    import tactile_gym_b200 as tg

    n, S = 37, 64
    env = tg.make_vec(env_id, n, seed=3, env_kwargs={"env_modes": modes, "image_size": [S, S], "max_steps": max_steps}, rng=rng)
    env.reset()
    rs = np.random.RandomState(1)
    acts = rs.uniform(-0.25, 0.25, (60, n, env.world.act_dim)).astype(np.float32)
    _run(env, acts[:17])                       # mid-episode, rebuilds in flight
    ck = env.world.save_checkpoint()
    first = _run(env, acts[17:])
    env.world.load_checkpoint(ck)
    second = _run(env, acts[17:])
    assert sum(r["done"].sum() for r in first) > 3 * n      # several episode turnovers per env
    for k, (a, b) in enumerate(zip(first, second)):
        for key in a:
            if key == "term":
                assert len(a[key]) == len(b[key]) and all(np.array_equal(x, y) for x, y in zip(a[key], b[key])), (k, key)
            else:
                assert np.array_equal(a[key], b[key]), (k, key)
    assert not env.world.pipeline_error()
    # a world of another shape refuses the blob
    other = tg.make_vec(env_id, n + 1, seed=3, env_kwargs={"env_modes": modes, "image_size": [S, S], "max_steps": max_steps}, rng=rng)
    with pytest.raises(Exception):
        other.world.load_checkpoint(ck)
    other.close()
    env.close()


@pytest.mark.parametrize("env_id,modes", [("edge_follow-v0", EDGE), ("object_balance-v0", BALANCE)])
def test_nan_guard_ends_the_episode_of_a_diverged_env(env_id, modes):
    import tactile_gym_b200 as tg

    n, S = 9, 64
    env = tg.make_vec(env_id, n, seed=5, env_kwargs={"env_modes": modes, "image_size": [S, S], "max_steps": 50})
    env.reset()
    a = np.zeros((n, env.world.act_dim), dtype=np.float32)
    env.step(a)
    assert env.world.nan_resets() == 0
    st = env.world.get_state()
    st[4, 2] = np.nan                                  # one joint angle of env 4
    env.world.set_state(st)
    o, r, d, infos = env.step(a)
    assert d[4] and r[4] == 0.0 and d.sum() == 1 and "terminal_observation" in infos[4]
    assert env.world.nan_resets() == 1 and (env.world.lib.tg_pipeline_error(env.world.h, env.world._stream()) & 4)
    st = env.world.get_state()
    assert np.isfinite(st).all()                       # env 4 runs its next episode, the others never noticed
    o, r, d, infos = env.step(a)
    assert not d.any() and np.isfinite(r).all() and env.world.nan_resets() == 1
    env.close()
