"""object_balance-v0 (tactile_gym/rl_envs/nonprehensile_manipulation/object_balance/object_balance_env.py, object_mode
"pole") on the batched engine."""
from ..engine import TactileWorld, object_balance_config
from .base_tactile_env import BaseTactileEnv

env_modes_default = {
    "movement_mode": "xy",
    "control_mode": "TCP_velocity_control",
    "object_mode": "pole",
    "rand_gravity": False,
    "rand_embed_dist": False,
    "observation_mode": "tactile",
    "reward_mode": "dense",
    "arm_type": "ur5",
    "tactile_sensor_name": "tactip",
}


class ObjectBalanceEnv(BaseTactileEnv):
    def __init__(self, max_steps=1000, image_size=(64, 64), env_modes=env_modes_default, show_gui=False, show_tactile=False, device=0):
        super().__init__(max_steps, image_size, show_gui, show_tactile, arm_type=env_modes["arm_type"])
        self.movement_mode = env_modes["movement_mode"]
        self.control_mode = env_modes["control_mode"]
        self.object_mode = env_modes.get("object_mode", "pole")
        self.observation_mode = env_modes["observation_mode"]
        self.reward_mode = env_modes["reward_mode"]
        if self.reward_mode not in ("dense", "sparse"):
            raise ValueError("Incorrect reward_mode specified: {}".format(self.reward_mode))
        self.t_s_name = env_modes["tactile_sensor_name"]
        cfg, keep, draw = object_balance_config(env_modes, image_size, max_steps, n_envs=1)
        self.world = TactileWorld(cfg, keep, device=device, draw_fn=draw)
        self._finish_init()
