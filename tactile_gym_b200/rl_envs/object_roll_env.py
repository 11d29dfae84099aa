"""object_roll-v0 (tactile_gym/rl_envs/nonprehensile_manipulation/object_roll/object_roll_env.py) on the batched engine."""
from ..engine import TactileWorld, object_roll_config
from .base_tactile_env import BaseTactileEnv

env_modes_default = {
    "movement_mode": "xy",
    "control_mode": "TCP_velocity_control",
    "rand_init_obj_pos": False,
    "rand_obj_size": False,
    "rand_embed_dist": False,
    "observation_mode": "tactile",
    "reward_mode": "dense",
    "arm_type": "ur5",
    "tactile_sensor_name": "tactip",
}


class ObjectRollEnv(BaseTactileEnv):
    """Observation modes built: 'tactile' and 'tactile_and_feature' (the goal position in the TCP frame,
    object_roll_env.py:402-408)."""

    def __init__(self, max_steps=1000, image_size=(64, 64), env_modes=env_modes_default, show_gui=False, show_tactile=False, device=0):
        super().__init__(max_steps, image_size, show_gui, show_tactile, arm_type=env_modes["arm_type"])
        self.movement_mode = env_modes["movement_mode"]
        self.control_mode = env_modes["control_mode"]
        self.rand_init_obj_pos = env_modes.get("rand_init_obj_pos", False)
        self.rand_obj_size = env_modes.get("rand_obj_size", False)
        self.rand_embed_dist = env_modes.get("rand_embed_dist", False)
        self.observation_mode = env_modes["observation_mode"]
        self.reward_mode = env_modes["reward_mode"]
        if self.reward_mode not in ("dense", "sparse"):
            raise ValueError("Incorrect reward_mode specified: {}".format(self.reward_mode))
        self.t_s_name = env_modes["tactile_sensor_name"]
        cfg, keep, draw = object_roll_config(env_modes, image_size, max_steps, n_envs=1)
        self.world = TactileWorld(cfg, keep, device=device, draw_fn=draw)
        self._finish_init()
