"""gym.Env-shaped single-env view on the batched engine (N = 1).

Mirrors BaseTactileEnv (tactile_gym/rl_envs/base_tactile_env.py:12-324): old-gym API
reset() -> obs dict, step(a) -> (obs dict, float, bool, {}), seed(), close(), observation_space,
action_space, min_action / max_action.  Observation modes built: `oracle`, `tactile`, `tactile_and_feature`.
"""
import numpy as np

from .. import spaces

_Base = object
for _mod in ("gym", "gymnasium"):   # the reference's envs are old-gym Envs (4-tuple step); either base class will do
    try:  # pragma: no cover
        _Base = __import__(_mod).Env
        break
    except Exception:  # noqa: BLE001
        pass


class BaseTactileEnv(_Base):
    metadata = {"render.modes": ["rgb_array"]}

    def __init__(self, max_steps=250, image_size=(64, 64), show_gui=False, show_tactile=False, arm_type="ur5"):
        if show_gui or show_tactile:
            raise ValueError("show_gui / show_tactile need a display server; not available in the batched engine")
        self._max_steps = max_steps
        self._image_size = list(image_size)
        self.arm_type = arm_type
        self._seed = None
        self.world = None

    # subclasses build self.world (a TactileWorld with n_envs == 1) and call this
    def _finish_init(self):
        self.min_action, self.max_action = -0.25, 0.25
        self.act_dim = self.world.act_dim
        self.action_space = spaces.Box(low=self.min_action, high=self.max_action, shape=(self.act_dim,), dtype=np.float32)
        # observation modes (base_tactile_env.py:76-114, 247-282): "oracle" (the task's state vector), "tactile", and
        # "tactile_and_feature" for the tasks that define an extended feature; the visual modes are not built
        nf = self.world.nfeat
        built = ["oracle", "tactile"] + (["tactile_and_feature"] if nf else [])
        if self.observation_mode not in built:
            raise NotImplementedError("observation_mode %r: built modes are %s (SURVEY 8(f) item 3)" % (self.observation_mode, built))
        S = self._image_size[0]
        sp = {}
        if "oracle" in self.observation_mode:
            self.world.bind_oracle_obs()
            sp["oracle"] = spaces.Box(low=-np.inf, high=np.inf, shape=(self.world.n_oracle,), dtype=np.float32)
        if "tactile" in self.observation_mode:
            sp["tactile"] = spaces.Box(low=0, high=255, shape=(S, S, 1), dtype=np.uint8)
        if "feature" in self.observation_mode:
            sp["extended_feature"] = spaces.Box(low=-np.inf, high=np.inf, shape=(nf,), dtype=np.float32)
        self.observation_space = spaces.Dict(sp)
        # the reference resets inside the constructor, before any seed() (edge_follow_env.py:131)
        self.reset()

    def seed(self, seed=None):
        self._seed = seed
        return [self.world.seed([seed])[0]]

    def _obs(self):
        o = {}
        if "oracle" in self.observation_mode:
            o["oracle"] = self.get_oracle_obs()
        if "tactile" in self.observation_mode:
            o["tactile"] = self.world.obs[0].cpu().numpy()
        if "feature" in self.observation_mode:
            o["extended_feature"] = self.get_extended_feature_array()
        return o

    def get_oracle_obs(self):
        return self.world.bind_oracle_obs()[0, : self.world.n_oracle].cpu().numpy()

    def get_extended_feature_array(self):
        return self.world.feat[0, : self.world.nfeat].cpu().numpy()

    def reset(self):
        self.world.reset(render="tactile" in self.observation_mode)
        return self._obs()

    def step(self, action):
        torch = self.world.torch
        a = torch.as_tensor(np.asarray(action, dtype=np.float32).reshape(1, -1), device=self.world.device)
        # a single env must not auto-reset: gym semantics leave that to the caller
        self.world.physics_only(a)
        if "tactile" in self.observation_mode:
            self.world.raster_only()
        torch.cuda.synchronize(self.world.device)
        return self._obs(), float(self.world.reward[0].item()), bool(self.world.done[0].item()), {}

    def get_tactile_obs(self):
        if "tactile" not in self.observation_mode:      # the reference renders on demand (base_tactile_env.py:200-210)
            self.world.raster_only()
        return self.world.obs[0].cpu().numpy()

    def render(self, mode="rgb_array"):
        if mode != "rgb_array":
            return np.array([])
        t = self.get_tactile_obs()
        return np.repeat(t, 3, axis=2)

    def close(self):
        if self.world is not None:
            self.world.close()
            self.world = None
