"""CPU: host logic - scene compiler, config builder, registry, seeding, C-ABI exports (no GPU calls)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reduced_model_matches_link_table(oracle, edge_modes):
    """Folding the fixed joints keeps total mass and the TCP / camera frames of the 11-link table."""
    from tactile_gym_b200 import scene
    from tactile_gym_b200.engine import edge_follow_config

    cfg, keep = edge_follow_config(edge_modes, [128, 128], 200, 8)
    a = cfg.arm
    assert a.nb == 6 and a.topo == 0
    mj = scene.load_model_json("ur5", "tactip", "standard")
    moving_or_below = sum(l["mass"] for l in mj["links"][1:])  # all but base_link
    assert abs(sum(a.mass[i] for i in range(6)) - moving_or_below) < 1e-12
    # world TCP pose from the reduced model (python FK) == oracle's link-table FK
    m = oracle.load_model("ur5", "tactip", "standard", [0.65, 0, 0.035], [-np.pi, 0, np.pi / 2], np.zeros((6, 2)))
    q = np.array(keep[4])
    R, p = np.eye(3), np.zeros(3)
    for b in range(6):
        p = p + R @ np.array(a.jpos[b][:])
        ax = np.array(a.axis[b][:]); c, s = np.cos(q[b]), np.sin(q[b])
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        R = R @ np.array(a.jrot[b][:]).reshape(3, 3) @ (np.eye(3) + s * K + (1 - c) * K @ K)
    tcp = p + R @ np.array(a.tcp_pos[:])
    P, Q = oracle.link_states(m, q)
    assert np.allclose(tcp, P[m.tcp_link], atol=1e-12)
    eye, *_ = oracle.camera_frame(m, q)
    assert np.allclose(p + R @ np.array(a.cam_pos[:]), eye, atol=1e-12)


def test_config_rejects_bad_modes(edge_modes):
    from tactile_gym_b200.engine import edge_follow_config

    bad = dict(edge_modes); del bad["tactile_sensor_name"]
    with pytest.raises(KeyError):
        edge_follow_config(bad, [128, 128], 200, 1)
    bad = dict(edge_modes, arm_type="pr2")
    with pytest.raises(ValueError):
        edge_follow_config(bad, [128, 128], 200, 1)
    cfg, _ = edge_follow_config(dict(edge_modes, control_mode="TCP_position_control", arm_type="mg400"), [128, 128], 200, 1)
    assert (cfg.task.control_mode, cfg.arm.nb) == (1, 8)                             # mg400.py:131-175
    with pytest.raises(ValueError):
        edge_follow_config(dict(edge_modes, control_mode="joint_torque_control"), [128, 128], 200, 1)
    cfg, _ = edge_follow_config(dict(edge_modes, control_mode="TCP_position_control"), [128, 128], 200, 1)
    assert (cfg.task.control_mode, cfg.task.pos_max_steps, cfg.task.act_hi[0]) == (1, 10, 0.001)     # edge_follow_env.py:143-153


def test_object_tasks_take_position_control():
    """control_mode="TCP_position_control" on the object tasks: pose-delta action ranges of the reference (1 mm / 1 degree per step:
    object_balance_env.py:129-139, object_push_env.py:137-147, object_roll_env.py:112-122) and its 10-step blocking move"""
    from tactile_gym_b200.engine import object_balance_config, object_push_config, object_roll_config

    base = {"control_mode": "TCP_position_control", "observation_mode": "tactile", "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "tactip"}
    deg = np.pi / 180
    cfg = object_balance_config(dict(base, movement_mode="xyRxRy", object_mode="pole", rand_gravity=True, rand_embed_dist=True), [64, 64], 250, 2)[0]
    assert (cfg.task.control_mode, cfg.task.pos_max_steps) == (1, 10)
    assert np.allclose(list(cfg.task.act_hi), [0.001, 0.001, 0.001, deg, deg, 0.0]) and np.allclose(list(cfg.task.act_lo), [-0.001, -0.001, -0.001, -deg, -deg, 0.0])
    cfg = object_push_config(dict(base, movement_mode="TyRz", traj_type="simplex", rand_init_orn=False, rand_obj_mass=False), [64, 64], 1000, 2)[0]
    assert (cfg.task.control_mode, cfg.task.pos_max_steps) == (1, 10) and np.allclose(list(cfg.task.act_hi), [0.001, 0.001, 0, 0, 0, deg])
    cfg = object_roll_config(dict(base, movement_mode="xy", rand_init_obj_pos=True, rand_obj_size=True, rand_embed_dist=True), [64, 64], 250, 2)[0]
    assert (cfg.task.control_mode, cfg.task.pos_max_steps) == (1, 10) and np.allclose(list(cfg.task.act_hi), [0.001, 0.001, 0, 0, 0, 0])
    vel = object_push_config(dict(base, control_mode="TCP_velocity_control", movement_mode="TyRz", traj_type="simplex", rand_init_orn=False,
                                  rand_obj_mass=False), [64, 64], 1000, 2)[0]
    assert vel.task.control_mode == 0 and np.allclose(list(vel.task.act_hi), [0.01, 0.01, 0, 0, 0, 5 * deg])
    with pytest.raises(ValueError):
        object_balance_config(dict(base, control_mode="joint_velocity_control", movement_mode="xy", object_mode="pole"), [64, 64], 250, 2)


def test_registry_ids():
    import tactile_gym_b200 as tg

    # the reference's ids (tactile_gym/rl_envs/__init__.py:3-41)
    for env_id in ("edge_follow-v0", "surface_follow-v0", "surface_follow-v1", "surface_follow-v2", "object_roll-v0", "object_push-v0", "object_balance-v0"):
        assert env_id in tg.REGISTRY
    with pytest.raises(KeyError):
        tg.make("no_such_env-v0")
    with pytest.raises(NotImplementedError):
        tg.make("edge_follow_aotu-v0")      # registered by the reference for a class that its sources do not define


def test_surface_config_modes():
    """host-side mapping of the surface_follow variants / modes to the task description (no GPU needed)"""
    from tactile_gym_b200.engine import surface_follow_config, surface_follow_goal_config, surface_follow_vert_config

    m = {"movement_mode": "xRz", "control_mode": "TCP_velocity_control", "noise_mode": "simplex", "observation_mode": "oracle",
         "reward_mode": "sparse", "arm_type": "ur5", "tactile_sensor_name": "digit"}
    cfg, _, draw = surface_follow_vert_config(m, [64, 64], 200, 2)
    t = cfg.task
    assert (t.act_dim, list(t.act_index[:2]), t.surf_mode, t.surf_dir_mode, t.surf_drive_y_only, t.sparse_reward) == (2, [0, 5], 2, 1, 1, 1)
    assert (t.surf_w_goal, t.surf_w_surf, t.surf_w_norm) == (0.0, 10.0, 3.0) and abs(t.surf_drive - 0.25 * 0.7) < 1e-15
    assert t.act_hi[5] == 0.0                      # horizontal surface: no yaw range (base_surface_env.py:197-206)
    cfg, _, _ = surface_follow_vert_config(dict(m, noise_mode="vertical_simplex", arm_type="mg400", tactile_sensor_name="tactip"), [64, 64], 200, 2)
    assert (cfg.task.surf_mode, cfg.task.surf_vertical, cfg.task.act_hi[5] > 0) == (3, 1, True)     # base_surface_env.py:83-107, 183-194
    cfg, _, _ = surface_follow_config(dict(m, movement_mode="yzRx", reward_mode="dense"), [64, 64], 200, 2)
    assert (cfg.task.act_dim, cfg.task.surf_mode, cfg.task.surf_dir_mode, cfg.task.sparse_reward) == (2, 1, 1, 0)
    cfg, _, _ = surface_follow_goal_config(dict(m, movement_mode="yz", noise_mode="none"), [64, 64], 200, 2)
    assert (cfg.task.act_dim, list(cfg.task.act_index[:2]), cfg.task.surf_mode) == (2, [1, 2], 2)
    cfg, _, _ = surface_follow_config(dict(m, movement_mode="xyz", noise_mode="random", reward_mode="dense"), [64, 64], 200, 2)
    assert (cfg.task.surf_mode, cfg.task.draw_kind[0], cfg.task.draw_kind[1]) == (4, 0, 1)     # base_surface_env.py:302-318: device RNG only
    with pytest.raises(ValueError):
        surface_follow_config(dict(m, movement_mode="xyz", noise_mode="perlin"), [64, 64], 200, 2)
    with pytest.raises(ValueError):
        surface_follow_config(dict(m, movement_mode="xyz", reward_mode="shaped"), [64, 64], 200, 2)
    # draws follow the reference's RNG call order: randint(1e8) only for simplex, choice([-1, 1]) for the 1-d modes
    from tactile_gym_b200 import seeding

    a, b = seeding.np_random(7)[0], seeding.np_random(7)[0]
    d = draw(a, 3)
    for r in range(3):
        assert d[r, 0] == b.randint(1e8) and d[r, 1] == b.choice([-1, 1])


def test_seeding_matches_oracle_restatement(oracle):
    from tactile_gym_b200 import seeding

    for s in (0, 1, 12345):
        a, _ = seeding.np_random(s)
        assert np.array_equal(a.uniform(size=4), oracle.gym_np_random(s).uniform(size=4))
    with pytest.raises(ValueError):
        seeding.np_random(-1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tactile_gym_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert not re.search(r"#include\s*[<\"].*oracle", src), f
                assert "libtg_oracle" not in src, f


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads and exports every function include/tactile_gym_b200.h declares."""
    import __graft_entry__ as g
    from tactile_gym_b200 import _lib

    g.build()
    header = open(os.path.join(ROOT, "include", "tactile_gym_b200.h")).read()
    declared = set(re.findall(r"\b(tg_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.tg_version() == 100


def test_no_gpu_fails_loudly(edge_modes):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import tactile_gym_b200 as tg
    from tactile_gym_b200._lib import TgError

    with pytest.raises(TgError):
        tg.make("edge_follow-v0", env_modes=edge_modes, image_size=[64, 64])


def test_quantize_shortcut_is_exact(tmp_path):
    """tg_raster.cuh replaces pen / 0.05f by a reciprocal + 2 FMA; the uint8 result must never differ.
    (every 7th float of [0, 0.05f] here - 147 M values; run tools/check_quantize.c without arguments for all 1.03e9)"""
    exe = str(tmp_path / "check_quantize")
    subprocess.check_call(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tools", "check_quantize.c"), "-lm"])
    out = subprocess.run([exe, "7"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert " 0 uint8 mismatches" in out.stdout


def test_balance_draws_emulate_the_numpy_call_sequence(oracle):
    """object_balance draws are generated in batches from the raw MT19937 stream; they must equal the reference's
    sequence of np_random.uniform / choice / rand calls (object_balance_env.py:300-313, 366-371)."""
    from tactile_gym_b200 import seeding
    from tactile_gym_b200.engine import object_balance_draws

    for rg, re_ in [(True, True), (False, True), (True, False), (False, False)]:
        fn = object_balance_draws(rg, re_, 0.003, 0.006, 0.0035)
        a = fn(seeding.np_random(11)[0], 9)
        rng = seeding.np_random(11)[0]
        b = np.array([oracle.balance_draws(rng, "tactip", rg, re_) for _ in range(9)])
        assert np.array_equal(a, b)


def test_pole_asset_composite():
    import json

    from tactile_gym_b200 import scene

    pole = json.load(open(os.path.join(scene.ASSETS, "objects", "pole.json")))
    assert abs(pole["mass"] - 0.11) < 1e-15
    assert abs(pole["com_off"][2] - ((0.01 * 0.00125 + 0.1 * 0.05) / 0.11 - 0.00125)) < 1e-15
    prims, nv = scene.merge_coplanar(scene.load_stimulus("pole"))
    assert prims.shape == (12, 4, 3) and (nv == 4).all()


def test_ctypes_structs_mirror_the_header(tmp_path):
    """Every field of the ctypes Structures in _lib.py sits at the offset the C compiler gives the header's struct (a drifted
    mirror would corrupt the task description silently): sizeof + offsetof of all fields, compiled from include/*.h."""
    from tactile_gym_b200 import _lib

    structs = ["TgArm", "TgPhysics", "TgTask", "TgSensor", "TgConfig", "TgHostStep"]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "tactile_gym_b200.h"', "int main(void) {"]
    for sn in structs:
        lines.append('printf("%s . %%zu\\n", sizeof(%s));' % (sn, sn))
        for fname, *_ in getattr(_lib, sn)._fields_:
            lines.append('printf("%s %s %%zu\\n", offsetof(%s, %s));' % (sn, fname, sn, fname))
    lines += ["return 0; }"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = str(tmp_path / "layout")
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, str(src)])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    n = 0
    for line in out.splitlines():
        sn, fname, val = line.split()
        cls = getattr(_lib, sn)
        want = C.sizeof(cls) if fname == "." else getattr(cls, fname).offset
        assert int(val) == want, (sn, fname, int(val), want)
        n += 1
    assert n >= 150


def test_reference_training_params_are_accepted():
    """Every env construction the reference's own training set-ups use (tests/golden/reference_params.json: the rl_params_ppo /
    rl_params_sac dicts of tactile_gym/sb3_helpers/params/*_params.py, extracted by tools/make_reference_params.py) builds an
    engine task description as it stands - or fails the way it is documented to."""
    import json

    from tactile_gym_b200.vec_env import CONFIG_BUILDERS

    sets = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_params.json")))
    assert len(sets) == 21      # 14 training parameter sets + the 7 demo scripts' (examples/demo_*_env.py)
    built = 0
    for p in sets:
        modes, env_id = p["env_modes"], p["env_name"]
        args = (modes, p["image_size"], p["max_ep_len"], 2)
        if "arm_type" not in modes or "tactile_sensor_name" not in modes:
            # the SAC dicts predate the arm_type / tactile_sensor_name keys: the reference's own constructors raise KeyError on
            # them (edge_follow_env.py:41,59), and so does this engine
            with pytest.raises(KeyError):
                CONFIG_BUILDERS[env_id](*args)
            continue
        out = CONFIG_BUILDERS[env_id](*args)
        cfg = out[0]
        assert cfg.n_envs == 2 and cfg.task.max_steps == p["max_ep_len"] and cfg.sensor.image_size == p["image_size"][0]
        built += 1
    # every complete PPO set-up (edge, balance, push on the MG400 + mini TacTip, roll, surface -v0, -v1, -v2 on its vertical
    # surface) and every demo script
    assert built == 14
