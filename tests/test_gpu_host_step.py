"""tg_step_host (host buffers in / out, chunked raster with the device->host copies overlapped on the library's copy stream)
against the plain tg_step + torch copies it replaces in TactileVecEnv: same seeds, same actions -> identical bytes, for every
raster kernel (polygon, heightfield, sphere), with features, with ragged chunk sizes, through episode turnovers."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

EDGE = {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "noise_mode": "rand_height", "observation_mode": "tactile",
        "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "tactip"}
SURF = {"movement_mode": "xyzRxRy", "control_mode": "TCP_velocity_control", "noise_mode": "simplex", "observation_mode": "tactile",
        "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "digit"}
PUSH = {"movement_mode": "TyRz", "control_mode": "TCP_velocity_control", "rand_init_orn": False, "rand_obj_mass": False,
        "traj_type": "simplex", "observation_mode": "tactile_and_feature", "reward_mode": "dense", "arm_type": "ur5",
        "tactile_sensor_name": "tactip"}
ROLL = {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "rand_init_obj_pos": True, "rand_obj_size": True,
        "rand_embed_dist": True, "observation_mode": "tactile_and_feature", "reward_mode": "dense", "arm_type": "ur5",
        "tactile_sensor_name": "tactip"}

CASES = [("edge_follow-v0", EDGE, 37, 64, 3, 16), ("edge_follow-v0", EDGE, 64, 128, 200, 0), ("surface_follow-v0", SURF, 13, 128, 4, 3),
         ("object_push-v0", PUSH, 9, 64, 5, 2), ("object_roll-v0", ROLL, 10, 64, 4, 1), ("edge_follow-v0", EDGE, 3, 256, 2, 16)]


@pytest.mark.parametrize("env_id,modes,n,S,max_steps,chunks", CASES)
def test_host_step_equals_device_step(env_id, modes, n, S, max_steps, chunks):
    import tactile_gym_b200 as tg

    kw = {"env_modes": modes, "image_size": [S, S], "max_steps": max_steps}
    a = tg.make_vec(env_id, n, seed=11, env_kwargs=kw, copy_chunks=chunks)    # tg_step_host
    b = tg.make_vec(env_id, n, seed=11, env_kwargs=kw, copy_chunks=-1)        # tg_step + torch copies
    assert a._host_step and not b._host_step
    oa, ob = a.reset(), b.reset()
    assert np.array_equal(oa["tactile"], ob["tactile"])
    rng = np.random.RandomState(4)
    l0 = a.world.launch_count()
    ends = 0
    for k in range(2 * max_steps + 1 if max_steps < 10 else 6):
        act = rng.uniform(-0.25, 0.25, (n, a.world.act_dim)).astype(np.float32)
        oa, ra, da, ia = a.step(act)
        ob, rb, db, ib = b.step(act)
        assert np.array_equal(oa["tactile"], ob["tactile"]), k
        assert np.array_equal(ra, rb) and np.array_equal(da, db), k
        # the device tensors hold the same step too
        assert np.array_equal(a.world.obs.cpu().numpy(), oa["tactile"]) and np.array_equal(a.world.reward.cpu().numpy(), ra)
        if "extended_feature" in oa:
            assert np.array_equal(oa["extended_feature"], ob["extended_feature"]), k
        for i in np.nonzero(da)[0]:
            ends += 1
            assert np.array_equal(ia[i]["terminal_observation"]["tactile"], ib[i]["terminal_observation"]["tactile"])
            assert ia[i]["episode"]["l"] == ib[i]["episode"]["l"] and ia[i]["episode"]["r"] == ib[i]["episode"]["r"]
    if max_steps < 10:
        assert ends >= 2 * n
    assert np.array_equal(a.world.get_state(), b.world.get_state())
    assert a.world.launch_count() > l0
    a.close(); b.close()


def test_host_step_needs_the_reset_pipeline():
    """max_steps < 2 has no standby pipeline: the VecEnv keeps the torch copy path there, and the C entry point refuses"""
    import torch

    import tactile_gym_b200 as tg
    from tactile_gym_b200 import _lib as L

    env = tg.make_vec("edge_follow-v0", 4, seed=1, env_kwargs={"env_modes": EDGE, "image_size": [64, 64], "max_steps": 1})
    assert not env._host_step
    env.reset()
    obs, rew, done, infos = env.step(np.zeros((4, 2), np.float32))
    assert done.all() and "terminal_observation" in infos[0]
    pin = torch.zeros((4, 2)).pin_memory()
    with pytest.raises(L.TgError):
        env.world.step_host(pin, env._pin_obs, env._pin_rew, env._pin_done)
    env.close()
