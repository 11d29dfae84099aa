"""surface_follow-v2 on its own surface (noise_mode "vertical_simplex": upright heightfield, `forward` sensor type) and
TCP_position_control on the MG400 (slaved joints in the IK result, mg400.py:167-172): CUDA path against the CPU oracle.

Both were written after round 1's GPU budget was spent, gated off, and passed on the first B200 they ever saw (4 XPASS in
GPUTEST_r01.json); the gates are gone and these are plain tests now.  The oracle side is pinned to the reference's source on the
CPU (tests/test_oracle_reference_golden.py::test_vertical_surface_geometry_and_rewards).  The bodies stay scripts so that
tests/test_vertical_posctl_scripts_dryrun.py can run them on the CPU against an oracle-backed stand-in."""
import os

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys
sys.path.insert(0, %(root)r)
import numpy as np
import tactile_gym_b200 as tg
from oracle import oracle as O

O.build()
arm, sensor, S, n = %(arm)r, %(sensor)r, %(S)d, 5
modes = {"movement_mode": "xRz", "control_mode": "TCP_velocity_control", "noise_mode": "vertical_simplex", "observation_mode": %(obs)r,
         "reward_mode": "dense", "arm_type": arm, "tactile_sensor_name": sensor}
env = tg.make_vec("surface_follow-v2", n, env_kwargs={"env_modes": modes, "image_size": [S, S], "max_steps": 200})
rng = np.random.RandomState(S)
draws = np.stack([rng.randint(0, 10 ** 8, (n, 2)).astype(np.float64), rng.choice([-1.0, 1.0], (n, 2))], axis=2)
env.world.set_draws(draws)
obs = env.reset()
st = env.world.get_state()
nb = env.world.nb
refs = []


def sync(r, row):
    for k in range(nb):
        r.s.q[k] = row[k]; r.s.qd[k] = row[nb + k]
    r.steps = int(row[2 * nb + 9])


def img_close(a, b):
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    return d.max(), (d != 0).mean()


for i in range(n):
    r = O.SurfaceFollowOracle(image_size=S, arm=arm, sensor=sensor, movement_mode="xRz", variant="vert", noise_mode="vertical_simplex",
                              render=%(render)s)
    r.reset(draws=(draws[i, 0, 0], draws[i, 0, 1]))
    refs.append(r)
    assert np.allclose(st[i, :nb], np.array(r.s.q[:nb]), atol=5e-6), ("reset pose", i, np.abs(st[i, :nb] - np.array(r.s.q[:nb])).max())
    assert st[i, 2 * nb + 10] == r.last_reset_substeps, ("reset substeps", i)
    sync(r, st[i]); r.step_data()
    if %(obs)r == "oracle":
        assert np.allclose(obs["oracle"][i], r.oracle_obs(), atol=2e-5), ("reset oracle obs", i, np.abs(obs["oracle"][i] - r.oracle_obs()).max())
    else:
        mx, frac = img_close(r.observation(), obs["tactile"][i])
        assert mx <= 1 and frac < 2e-3, ("reset image", i, mx, frac)
touched = 0
for k in range(%(steps)d):
    act = rng.uniform(-0.25, 0.25, (n, 2)).astype(np.float32)
    for i, r in enumerate(refs):
        sync(r, st[i])
        r.step(act[i])
    o2, rew, done, infos = env.step(act)
    st = env.world.get_state()
    for i, r in enumerate(refs):
        assert np.allclose(st[i, :nb], np.array(r.s.q[:nb]), atol=1e-9), ("joints", k, i)
        assert abs(rew[i] - r.reward) < 1e-6 * max(1.0, abs(r.reward)) and bool(done[i]) == r.done, ("reward", k, i, rew[i], r.reward)
        if %(obs)r == "oracle":
            assert np.allclose(o2["oracle"][i], r.oracle_obs(), atol=2e-5), ("oracle obs", k, i)
        else:
            img = r.observation()
            mx, frac = img_close(img, o2["tactile"][i])
            assert mx <= 1 and frac < 2e-3, ("image", k, i, mx, frac)
            touched += int((img[..., 0][r.ref[2] == 0] > 0).sum() > 50)
assert %(obs)r == "oracle" or touched > %(steps)d, "the surface never showed in the images"
assert not env.world.pipeline_error()
env.close()
print("VERTICAL-OK")
'''


@pytest.mark.parametrize("arm,sensor,S,obs_mode", [("mg400", "tactip", 128, "tactile"), ("ur5", "digit", 128, "tactile"), ("mg400", "tactip", 64, "oracle")])
def test_vertical_surface_matches_oracle(capsys, arm, sensor, S, obs_mode):
    code = CHILD % {"root": ROOT, "arm": arm, "sensor": sensor, "S": S, "obs": obs_mode, "render": "True" if obs_mode == "tactile" else "False",
                    "steps": 8}
    exec(compile(code, "vertical-child", "exec"), {"__name__": "child"})
    assert "VERTICAL-OK" in capsys.readouterr().out


CHILD_POSCTL = r'''
import sys
sys.path.insert(0, %(root)r)
import numpy as np
import tactile_gym_b200 as tg
from oracle import oracle as O

O.build()
n, nb = 5, 8
modes = {"movement_mode": "xyzRz", "control_mode": "TCP_position_control", "noise_mode": "rand_height", "observation_mode": "oracle",
         "reward_mode": "dense", "arm_type": "mg400", "tactile_sensor_name": "digitac"}
env = tg.make_vec("edge_follow-v0", n, env_kwargs={"env_modes": modes, "image_size": [64, 64], "max_steps": 200})
rng = np.random.RandomState(17)
draws = np.stack([rng.uniform(0.0015, 0.0045, (n, 2)), rng.uniform(-np.pi, np.pi, (n, 2))], axis=2)
env.world.set_draws(draws)
obs = env.reset()["oracle"]
st = env.world.get_state()
refs = []
for i in range(n):
    r = O.EdgeFollowOracle(image_size=64, arm="mg400", sensor="digitac", movement_mode="xyzRz", control_mode="TCP_position_control")
    r.reset(draws=tuple(draws[i, 0]))
    refs.append(r)
p0 = obs[:, 0:3].copy()
for k in range(8):
    act = rng.uniform(-0.25, 0.25, (n, 4)).astype(np.float32)
    if k < 5:
        act[:, 0] = 0.25
    for i, r in enumerate(refs):
        for j in range(nb):
            r.s.q[j] = st[i, j]; r.s.qd[j] = st[i, nb + j]
        r.steps = int(st[i, 2 * nb + 9])
        r.step(act[i])
    o, rew, done, infos = env.step(act)
    st = env.world.get_state()
    for i, r in enumerate(refs):
        assert np.allclose(st[i, :nb], np.array(r.s.q[:nb]), atol=1e-9), ("joints", k, i, np.abs(st[i, :nb] - np.array(r.s.q[:nb])).max())
        assert abs(rew[i] - r.reward) < 1e-6 and bool(done[i]) == r.done, ("reward", k, i)
        assert np.allclose(o["oracle"][i], r.oracle_obs(), atol=2e-5), ("oracle obs", k, i)
    if k == 4:
        moved = np.abs(o["oracle"][:, 0] - p0[:, 0])
        assert np.all(moved > 0.004) and np.all(moved < 0.0055), moved
env.close()
print("POSCTL-OK")
'''


def test_mg400_position_control_matches_oracle(capsys):
    exec(compile(CHILD_POSCTL % {"root": ROOT}, "posctl-child", "exec"), {"__name__": "child"})
    assert "POSCTL-OK" in capsys.readouterr().out
