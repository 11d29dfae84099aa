"""surface_follow-v0 (tactile_gym/rl_envs/exploration/surface_follow/surface_follow_auto/surface_follow_auto_env.py on
base_surface_env.py, noise_mode "simplex") on the batched engine."""
import numpy as np

from .. import spaces
from ..engine import TactileWorld, surface_follow_config, surface_follow_goal_config
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
        if self.reward_mode != "dense":
            raise NotImplementedError("reward_mode %r: only 'dense' is built" % self.reward_mode)
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
        if self.reward_mode != "dense":
            raise NotImplementedError("reward_mode %r: only 'dense' is built" % self.reward_mode)
        self.t_s_name = env_modes["tactile_sensor_name"]
        cfg, keep, draw = surface_follow_goal_config(env_modes, image_size, max_steps, n_envs=1)
        self.world = TactileWorld(cfg, keep, device=device, draw_fn=draw)
        self.min_action, self.max_action = -0.25, 0.25
        self.act_dim = self.world.act_dim
        self.action_space = spaces.Box(low=self.min_action, high=self.max_action, shape=(self.act_dim,), dtype=np.float32)
        if self.observation_mode not in ("tactile", "tactile_and_feature"):
            raise NotImplementedError("observation_mode %r: only 'tactile' and 'tactile_and_feature' are built" % self.observation_mode)
        S = self._image_size[0]
        sp = {"tactile": spaces.Box(low=0, high=255, shape=(S, S, 1), dtype=np.uint8)}
        if self.observation_mode == "tactile_and_feature":
            sp["extended_feature"] = spaces.Box(low=-np.inf, high=np.inf, shape=(6,), dtype=np.float32)
        self.observation_space = spaces.Dict(sp)
        self.reset()

    def _obs(self):
        o = {"tactile": self.world.obs[0].cpu().numpy()}
        if self.observation_mode == "tactile_and_feature":
            o["extended_feature"] = self.world.feat[0, :6].cpu().numpy()
        return o

    def get_extended_feature_array(self):
        return self.world.feat[0, :6].cpu().numpy()
