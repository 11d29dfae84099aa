"""Raster work counters (diagnostic build with -DTG_RASTER_STATS, see tg_raster.cuh RSTAT): where do the exact-path
pixels and the partial candidates come from?  usage: raster_stats.py build | run [edge|balance] [steps]"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tools", "_diag", "libtg_stats.so")
NAMES = ["tiles", "tiles_with_dominator", "in_prims", "part_prims", "unc_edge_px", "unc_px", "unc_px_x_cands", "spans_shaded",
         "span_part_prims", "steep_prims", "valid_prims", "clipped_images"]

if sys.argv[1] == "build":
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--shared",
                           "-DTG_RASTER_STATS", "-Xcompiler", "-fPIC", "-o", LIB, os.path.join(ROOT, "tactile_gym_b200", "csrc", "tg_world.cu")])
    sys.exit(0)

os.environ["TG_LIB_OVERRIDE"] = LIB
sys.path.insert(0, ROOT)
import torch
import tactile_gym_b200 as tg
from tactile_gym_b200 import _lib

task = sys.argv[2] if len(sys.argv) > 2 else "balance"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
if task == "balance":
    env_id, n, S = "object_balance-v0", 2048, 256
    modes = {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "object_mode": "pole", "rand_gravity": True,
             "rand_embed_dist": True, "observation_mode": "tactile", "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "tactip"}
else:
    env_id, n, S = "edge_follow-v0", 4096, 128
    modes = {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "noise_mode": "rand_height",
             "observation_mode": "tactile", "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "tactip"}
env = tg.make_vec(env_id, n, seed=1, env_kwargs={"env_modes": modes, "image_size": [S, S], "max_steps": 250})
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(0)
lib = _lib.load()
out = (C.c_ulonglong * 48)()
for k in range(steps):
    env.step_tensor((torch.rand((n, 2), device="cuda", generator=g) - 0.5) * 0.5)
    if k in (0, steps // 2, steps - 1):
        lib.tg_debug_raster_stats(out)          # clear
        env.world.raster_only()
        lib.tg_debug_raster_stats(out)
        d = dict(zip(NAMES, list(out)))
        print("step %d (per image, %d images):" % (k + 1, n), {kk: round(v / n, 2) for kk, v in d.items()})
        print("   partial tiles per prim:", [round(out[16 + t] / n, 1) for t in range(12)], " in tiles per prim:", [round(out[32 + t] / n, 1) for t in range(12)])
env.close()
