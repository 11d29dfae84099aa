"""gym.spaces when gym/gymnasium is installed, else minimal duck-typed stand-ins (Box, Dict).

The reference builds its observation / action spaces from gym.spaces (rl_envs/base_tactile_env.py:76-114,
rl_envs/exploration/edge_follow/edge_follow_env.py:167-174).  gym is not a hard dependency here.
"""
import numpy as np

_spaces = None
for _mod in ("gymnasium", "gym"):   # SB3 >= 2.0 checks isinstance against gymnasium's spaces; the reference itself is on old gym
    try:  # pragma: no cover - depends on the environment
        _spaces = __import__(_mod, fromlist=["spaces"]).spaces
        break
    except Exception:  # noqa: BLE001
        _spaces = None
HAVE_GYM = _spaces is not None
if HAVE_GYM:  # pragma: no cover
    Box, Dict = _spaces.Box, _spaces.Dict
else:

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            self.shape = tuple(shape) if shape is not None else np.shape(low)
            self.low = np.full(self.shape, low, dtype=self.dtype) if np.isscalar(low) else np.asarray(low, dtype=self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype) if np.isscalar(high) else np.asarray(high, dtype=self.dtype)
            self._rng = np.random.RandomState()

        def seed(self, seed=None):
            self._rng = np.random.RandomState(seed)
            return [seed]

        def sample(self):
            if np.issubdtype(self.dtype, np.integer):
                return self._rng.randint(self.low, self.high.astype(np.int64) + 1, size=self.shape).astype(self.dtype)
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return self._rng.uniform(lo, hi, size=self.shape).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return "Box(%s, %s, %s, %s)" % (self.low.min(), self.high.max(), self.shape, self.dtype)

    class Dict:
        def __init__(self, spaces):
            self.spaces = dict(spaces)

        def __getitem__(self, k):
            return self.spaces[k]

        def keys(self):
            return self.spaces.keys()

        def items(self):
            return self.spaces.items()

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def __repr__(self):
            return "Dict(%s)" % ", ".join("%s: %r" % kv for kv in self.spaces.items())
