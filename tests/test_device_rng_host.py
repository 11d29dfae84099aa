"""CPU: the device RNG (tactile_gym_b200/csrc/tg_rng.cuh, MT19937 with numpy's legacy call semantics) compiled for the host
and run against numpy's RandomState itself - uniform, randint(1e8) with its masked rejection, choice([-1, 1]), choice * rand,
across several regenerations of the 624-word state - and the per-task draw programs (TgTask.draw_kind) against the host draw
functions they replace (engine.*_draws), which tests/test_oracle_reference_golden.py pins to the reference's own reset code."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("rng") / "rng_harness")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "csrc", "rng_host_harness.cpp")])
    return exe


def _run(exe, rng, program, rounds):
    st = rng.get_state()
    inp = " ".join(str(int(k)) for k in st[1]) + "\n%d %d %d\n" % (st[2], len(program), rounds)
    inp += "\n".join("%d %.17g %.17g" % p for p in program) + "\n"
    out = subprocess.run([exe], input=inp, capture_output=True, text=True, check=True).stdout.split()
    return np.array([float(x) for x in out[:-1]]).reshape(rounds, len(program)), int(out[-1])


def test_primitives_match_numpy(harness):
    from tactile_gym_b200 import _lib as L, seeding

    prog = [(L.TG_DRAW_UNIFORM, -np.pi, np.pi), (L.TG_DRAW_RANDINT, 0.0, 1e8), (L.TG_DRAW_CHOICE_PM1, 0.0, 0.0),
            (L.TG_DRAW_CHOICE_RAND, 0.0, 0.0), (L.TG_DRAW_CONST, 0.0, 0.0), (L.TG_DRAW_UNIFORM, 0.0015, 0.0065)]
    rounds = 400                                           # ~2,900 outputs: the state is regenerated four times on the way
    a, b = seeding.np_random(42)[0], seeding.np_random(42)[0]
    got, pos = _run(harness, a, prog, rounds)
    for r in range(rounds):
        want = [b.uniform(-np.pi, np.pi), float(b.randint(1e8)), float(b.choice([-1, 1])), b.choice([-1, 1]) * b.rand(), -7.0,
                b.uniform(0.0015, 0.0065)]
        assert list(got[r]) == want, r
    assert pos == b.get_state()[2]
    # a state taken mid-stream (pos != 624) continues the same way
    got2, _ = _run(harness, b, prog[:2], 5)
    c = seeding.np_random(42)[0]
    c.set_state(b.get_state())
    assert list(got2[0]) == [c.uniform(-np.pi, np.pi), float(c.randint(1e8))]


@pytest.mark.parametrize("env_id,modes", [
    ("edge_follow-v0", {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "noise_mode": "rand_height", "observation_mode": "tactile",
                        "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "tactip"}),
    ("edge_follow-v0", {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "noise_mode": "fixed_height", "observation_mode": "tactile",
                        "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "digit"}),
    ("object_balance-v0", {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "object_mode": "pole", "rand_gravity": True,
                           "rand_embed_dist": True, "observation_mode": "tactile", "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "tactip"}),
    ("object_balance-v0", {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "object_mode": "pole", "rand_gravity": False,
                           "rand_embed_dist": True, "observation_mode": "tactile", "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "digit"}),
    ("surface_follow-v0", {"movement_mode": "xyzRxRy", "control_mode": "TCP_velocity_control", "noise_mode": "simplex", "observation_mode": "tactile",
                           "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "digit"}),
    ("surface_follow-v0", {"movement_mode": "yz", "control_mode": "TCP_velocity_control", "noise_mode": "none", "observation_mode": "tactile",
                           "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "digit"}),
    ("object_push-v0", {"movement_mode": "TyRz", "control_mode": "TCP_velocity_control", "rand_init_orn": True, "rand_obj_mass": True, "traj_type": "simplex",
                        "observation_mode": "tactile_and_feature", "reward_mode": "dense", "arm_type": "mg400", "tactile_sensor_name": "digitac"}),
    ("object_push-v0", {"movement_mode": "TyRz", "control_mode": "TCP_velocity_control", "rand_init_orn": False, "rand_obj_mass": False, "traj_type": "straight",
                        "observation_mode": "tactile_and_feature", "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "tactip"}),
    ("object_roll-v0", {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "rand_obj_size": True, "rand_embed_dist": True, "rand_init_obj_pos": True,
                        "observation_mode": "tactile", "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "tactip"}),
    ("object_roll-v0", {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "rand_obj_size": False, "rand_embed_dist": False, "rand_init_obj_pos": False,
                        "observation_mode": "tactile", "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "tactip"}),
])
def test_draw_programs_equal_the_host_draw_functions(harness, env_id, modes):
    """what the device RNG produces for a task's program == what the host-side draw function (the path every parity test has used
    so far) produces from the same generator, draw for draw, 70 resets deep"""
    from tactile_gym_b200 import seeding
    from tactile_gym_b200.engine import edge_follow_draws
    from tactile_gym_b200.vec_env import CONFIG_BUILDERS

    built = CONFIG_BUILDERS[env_id](modes, [64, 64], 200, 2)
    cfg, draw = built[0], (built[2] if len(built) == 3 else edge_follow_draws(built[0].task))
    t = cfg.task
    prog = [(int(t.draw_kind[d]), float(t.draw_lo[d]), float(t.draw_hi[d])) for d in range(t.n_draws)]
    a, b = seeding.np_random(9)[0], seeding.np_random(9)[0]
    got, pos = _run(harness, a, prog, 70)
    want = draw(b, 70)
    for d in range(t.n_draws):                              # constants: the harness prints its marker, the device takes draw_default
        if prog[d][0] == 0:
            got[:, d] = t.draw_default[d]
    assert np.array_equal(got, want), np.abs(got - want).max()
    assert pos == b.get_state()[2]
