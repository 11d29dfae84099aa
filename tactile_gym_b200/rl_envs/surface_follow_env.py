"""surface_follow-v0 (tactile_gym/rl_envs/exploration/surface_follow/surface_follow_auto/surface_follow_auto_env.py on
base_surface_env.py, noise_mode "simplex") on the batched engine."""
from ..engine import TactileWorld, surface_follow_config, surface_follow_goal_config, surface_follow_vert_config
from .base_tactile_env import BaseTactileEnv

env_modes_default = {
    "movement_mode": "xyzRxRy",
    "control_mode": "TCP_velocity_control",
    "noise_mode": "simplex",
    "observation_mode": "tactile",
    "reward_mode": "dense",
    "arm_type": "ur5",
    "tactile_sensor_name": "tactip",
}


class SurfaceFollowAutoEnv(BaseTactileEnv):
    def __init__(self, max_steps=200, image_size=(64, 64), env_modes=env_modes_default, show_gui=False, show_tactile=False, device=0):
        super().__init__(max_steps, image_size, show_gui, show_tactile, arm_type=env_modes["arm_type"])
        self.movement_mode = env_modes["movement_mode"]
        self.control_mode = env_modes["control_mode"]
        self.noise_mode = env_modes.get("noise_mode", "simplex")
        self.observation_mode = env_modes["observation_mode"]
        self.reward_mode = env_modes["reward_mode"]
        if self.reward_mode not in ("dense", "sparse"):
            raise ValueError("Incorrect reward_mode specified: {}".format(self.reward_mode))
        self.t_s_name = env_modes["tactile_sensor_name"]
        cfg, keep, draw = surface_follow_config(env_modes, image_size, max_steps, n_envs=1)
        self.world = TactileWorld(cfg, keep, device=device, draw_fn=draw)
        self._finish_init()


class SurfaceFollowGoalEnv(BaseTactileEnv):
    """surface_follow-v1 (tactile_gym/rl_envs/exploration/surface_follow/surface_follow_goal/surface_follow_goal_env.py).
    Observation modes built: 'tactile' and 'tactile_and_feature' (TCP + goal position in the work frame, :83-97)."""

    def __init__(self, max_steps=200, image_size=(64, 64), env_modes=env_modes_default, show_gui=False, show_tactile=False, device=0):
        super().__init__(max_steps, image_size, show_gui, show_tactile, arm_type=env_modes["arm_type"])
        self.movement_mode = env_modes["movement_mode"]
        self.control_mode = env_modes["control_mode"]
        self.noise_mode = env_modes.get("noise_mode", "simplex")
        self.observation_mode = env_modes["observation_mode"]
        self.reward_mode = env_modes["reward_mode"]
        if self.reward_mode not in ("dense", "sparse"):
            raise ValueError("Incorrect reward_mode specified: {}".format(self.reward_mode))
        self.t_s_name = env_modes["tactile_sensor_name"]
        cfg, keep, draw = surface_follow_goal_config(env_modes, image_size, max_steps, n_envs=1)
        self.world = TactileWorld(cfg, keep, device=device, draw_fn=draw)
        self._finish_init()


env_modes_default_vert = {
    "movement_mode": "xRz",
    "control_mode": "TCP_velocity_control",
    "noise_mode": "simplex",
    "observation_mode": "oracle",
    "reward_mode": "dense",
    "arm_type": "ur5",
    "tactile_sensor_name": "tactip",
}


class SurfaceFollowVertEnv(BaseTactileEnv):
    """surface_follow-v2 (tactile_gym/rl_envs/exploration/surface_follow/surface_follow_vert/surface_follow_vert_env.py) on the
    horizontal surfaces; noise_mode 'vertical_simplex' (vertical heightfield + `forward` sensors) raises NotImplementedError."""

    def __init__(self, max_steps=200, image_size=(64, 64), env_modes=env_modes_default_vert, show_gui=False, show_tactile=False, device=0):
        super().__init__(max_steps, image_size, show_gui, show_tactile, arm_type=env_modes["arm_type"])
        self.movement_mode = env_modes["movement_mode"]
        self.control_mode = env_modes["control_mode"]
        self.noise_mode = env_modes.get("noise_mode", "simplex")
        self.observation_mode = env_modes["observation_mode"]
        self.reward_mode = env_modes["reward_mode"]
        if self.reward_mode not in ("dense", "sparse"):
            raise ValueError("Incorrect reward_mode specified: {}".format(self.reward_mode))
        self.t_s_name = env_modes["tactile_sensor_name"]
        cfg, keep, draw = surface_follow_vert_config(env_modes, image_size, max_steps, n_envs=1)
        self.world = TactileWorld(cfg, keep, device=device, draw_fn=draw)
        self._finish_init()
