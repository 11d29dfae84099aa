"""The scanline raster for convex stimuli (csrc/tg_raster_scan.cuh) against the general raster kernel it replaces on the hot path
(csrc/tg_raster.cuh, itself pinned to the oracle by tests/test_gpu_parity.py - which now run through the scanline kernel too):
same images up to the fp64-vs-fp64 formulation noise (<= 1 LSB on < 0.1 % of the pixels), for the edge box, the cube and the
two-part pole, at 64 / 128 / 256 pixels; and the fallback plumbing (envs the scanline kernel hands back are rendered by the
general kernel in the masked second launch)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys
sys.path.insert(0, %(root)r)
import numpy as np
import tactile_gym_b200 as tg
env_id, S, n = %(env_id)r, %(S)d, 97
modes = %(modes)r
env = tg.make_vec(env_id, n, seed=5, env_kwargs={"env_modes": modes, "image_size": [S, S], "max_steps": %(max_steps)d})
obs = [env.reset()["tactile"].copy()]
rs = np.random.RandomState(0)
for k in range(%(steps)d):
    a = rs.uniform(-0.25, 0.25, (n, env.world.act_dim)).astype(np.float32)
    o, r, d, infos = env.step(a)
    obs.append(o["tactile"].copy())
    if d.any():
        obs.append(np.stack([infos[i]["terminal_observation"]["tactile"] for i in np.flatnonzero(d)]))
np.save(%(out)r, np.concatenate([x.reshape(-1, S * S) for x in obs]))
env.close()
'''

EDGE = {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "noise_mode": "rand_height", "observation_mode": "tactile",
        "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "tactip"}
BALANCE = {"movement_mode": "xyRxRy", "control_mode": "TCP_velocity_control", "object_mode": "pole", "rand_gravity": True, "rand_embed_dist": True,
           "observation_mode": "tactile", "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "tactip"}
PUSH = {"movement_mode": "TyRz", "control_mode": "TCP_velocity_control", "rand_init_orn": True, "rand_obj_mass": False, "traj_type": "simplex",
        "observation_mode": "tactile_and_feature", "reward_mode": "dense", "arm_type": "mg400", "tactile_sensor_name": "digitac"}


def _run(tmp_path, name, env_id, modes, S, extra_env, steps=10, max_steps=7):
    out = str(tmp_path / (name + ".npy"))
    code = CHILD % {"root": ROOT, "env_id": env_id, "modes": modes, "S": S, "out": out, "steps": steps, "max_steps": max_steps}
    env = dict(os.environ, **extra_env)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stderr[-3000:]
    return np.load(out)


@pytest.mark.parametrize("env_id,modes,S", [("edge_follow-v0", EDGE, 128), ("edge_follow-v0", dict(EDGE, tactile_sensor_name="digit"), 64),
                                            ("object_balance-v0", BALANCE, 256), ("object_balance-v0", BALANCE, 128), ("object_push-v0", PUSH, 128)])
def test_scanline_raster_equals_general_raster(tmp_path, env_id, modes, S):
    scan = _run(tmp_path, "scan", env_id, modes, S, {})
    gen = _run(tmp_path, "gen", env_id, modes, S, {"TG_NO_SCAN": "1"})
    half = _run(tmp_path, "half", env_id, modes, S, {"TG_SCAN_TEST_FALLBACK": "1"})     # odd envs through the fallback pass
    assert scan.shape == gen.shape == half.shape and scan.shape[0] > 900
    d = np.abs(scan.astype(np.int32) - gen.astype(np.int32))
    assert d.max() <= 1 and (d != 0).mean() < 1e-3, (d.max(), (d != 0).mean())
    assert (scan > 0).mean() > 0.05                       # the stimulus really shows
    d2 = np.abs(half.astype(np.int32) - gen.astype(np.int32))
    assert d2.max() <= 1 and (d2 != 0).mean() < 1e-3
    assert (half[1::2] != gen[1::2]).mean() <= (scan[1::2] != gen[1::2]).mean()      # the handed-back envs ARE the general kernel's


def test_scanline_raster_takes_tilted_poles(tmp_path):
    """Poles left to tip over for 70 steps (up to the 35 degree termination): the plate's far corners go behind the eye plane and
    beyond the far plane.  The scanline raster keeps those envs (homogeneous edge functions, see scan_setup_kernel) and its images
    equal the general kernel's; only near-plane cuts are handed back, and those stay rare."""
    scan = _run(tmp_path, "scan_t", "object_balance-v0", BALANCE, 128, {}, steps=70, max_steps=250)
    gen = _run(tmp_path, "gen_t", "object_balance-v0", BALANCE, 128, {"TG_NO_SCAN": "1"}, steps=70, max_steps=250)
    assert scan.shape == gen.shape
    d = np.abs(scan.astype(np.int32) - gen.astype(np.int32))
    assert d.max() <= 1 and (d != 0).mean() < 1e-3, (d.max(), (d != 0).mean())
