"""Short workload for ncu: BASELINE config 2 (edge_follow-v0, UR5+TacTip 128x128, 4096 envs) or config 5
(object_balance-v0, 256x256, 2048 envs), a few steps, then optionally 3 raster-only launches.
usage: prof_run.py [n] [S] [steps] [raster|-] [edge|balance|surface|push]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tactile_gym_b200 as tg

task = sys.argv[5] if len(sys.argv) > 5 else "edge"
if task == "balance":
    env_id, act_dim = "object_balance-v0", 2
    modes = {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "object_mode": "pole", "rand_gravity": True,
             "rand_embed_dist": True, "observation_mode": "tactile", "reward_mode": "dense", "arm_type": "ur5",
             "tactile_sensor_name": "tactip"}
elif task == "surface":
    env_id, act_dim = "surface_follow-v0", 3
    modes = {"movement_mode": "xyzRxRy", "control_mode": "TCP_velocity_control", "noise_mode": "simplex", "observation_mode": "tactile",
             "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "digit"}
elif task == "push":
    env_id, act_dim = "object_push-v0", 2
    modes = {"movement_mode": "TyRz", "control_mode": "TCP_velocity_control", "rand_init_orn": False, "rand_obj_mass": False,
             "traj_type": "simplex", "observation_mode": "tactile_and_feature", "reward_mode": "dense", "arm_type": "mg400",
             "tactile_sensor_name": "digitac"}
else:
    env_id, act_dim = "edge_follow-v0", 2
    modes = {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "noise_mode": "rand_height",
             "observation_mode": "tactile", "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "tactip"}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
env = tg.make_vec(env_id, n, seed=1, env_kwargs={"env_modes": modes, "image_size": [S, S], "max_steps": 200})
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(0)
for k in range(steps):
    a = (torch.rand((n, act_dim), device="cuda", generator=g) - 0.5) * 0.5
    env.step_tensor(a)
torch.cuda.synchronize()
if len(sys.argv) > 4 and sys.argv[4] == "raster":
    for k in range(3):
        env.world.raster_only()
    torch.cuda.synchronize()
env.close()
