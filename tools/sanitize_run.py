"""Workload of tools/sanitize.sh: a few steps of one task family on a small world (TG_SANITIZE_TASK)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import tactile_gym_b200 as tg

task = os.environ.get("TG_SANITIZE_TASK", "edge")
base = {"control_mode": "TCP_velocity_control", "observation_mode": "tactile", "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "tactip"}
cases = {
    "edge": ("edge_follow-v0", dict(base, movement_mode="xy", noise_mode="rand_height")),
    "surface": ("surface_follow-v0", dict(base, movement_mode="xyzRxRy", noise_mode="simplex", tactile_sensor_name="digit")),
    "balance": ("object_balance-v0", dict(base, movement_mode="xyRxRy", object_mode="pole", rand_gravity=True, rand_embed_dist=True)),
    "push": ("object_push-v0", dict(base, movement_mode="TyRz", rand_init_orn=True, rand_obj_mass=True, traj_type="simplex", arm_type="mg400",
                                    tactile_sensor_name="digitac", observation_mode="tactile_and_feature")),
    "roll": ("object_roll-v0", dict(base, movement_mode="xy", rand_init_obj_pos=True, rand_obj_size=True, rand_embed_dist=True,
                                    observation_mode="tactile_and_feature")),
}
env_id, modes = cases[task]
n = 33
env = tg.make_vec(env_id, n, seed=1, env_kwargs={"env_modes": modes, "image_size": [64, 64], "max_steps": 3})
env.reset()
rs = np.random.RandomState(0)
for k in range(8):
    env.step(rs.uniform(-0.25, 0.25, (n, env.world.act_dim)).astype(np.float32))
ck = env.world.save_checkpoint()
env.world.load_checkpoint(ck)
env.step(np.zeros((n, env.world.act_dim), dtype=np.float32))
env.close()
print("sanitize_run", task, "done")
