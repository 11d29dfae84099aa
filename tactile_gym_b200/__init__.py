"""tactile_gym_b200: batched, B200-native tactile-RL environment engine with tactile_gym's env surface."""
from . import _lib
from .rl_envs import REGISTRY, make
from .vec_env import TactileVecEnv

__all__ = ["make", "make_vec", "TactileVecEnv", "REGISTRY"]


def make_vec(env_id, n_envs, seed=None, env_kwargs=None, device=0, **kw):
    """Counterpart of stable_baselines3's make_vec_env(env_id, n_envs, seed, vec_env_cls=SubprocVecEnv, env_kwargs=...)
    as used in tactile_gym/sb3_helpers/rl_utils.py:17-30."""
    return TactileVecEnv(env_id, n_envs, seed=seed, env_kwargs=env_kwargs, device=device, **kw)
