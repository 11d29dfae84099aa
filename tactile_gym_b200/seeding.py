"""gym (<= 0.21) seeding, restated so seeds reproduce the reference's RNG stream without gym installed.

BaseTactileEnv.seed (rl_envs/base_tactile_env.py:61-64) calls gym.utils.seeding.np_random(seed); the envs
then use RandomState-only methods (uniform, randint), so the RandomState era of gym is the relevant one:
    hash_seed(seed) = first 8 bytes of sha512(str(seed)), read as little-endian uint32 words
    rng = np.random.RandomState(); rng.seed([words...])
"""
import hashlib
import os
import struct

import numpy as np


def create_seed(seed=None):
    if seed is None:
        return int.from_bytes(os.urandom(4), "little")
    if not (isinstance(seed, (int, np.integer)) and seed >= 0):
        raise ValueError("Seed must be a non-negative integer or omitted, not %r" % (seed,))
    return int(seed)


def np_random(seed=None):
    seed = create_seed(seed)
    h = hashlib.sha512(str(seed).encode("utf8")).digest()[:8]
    h += b"\0" * (4 - len(h) % 4)  # gym pads even when already aligned
    words = struct.unpack("%dI" % (len(h) // 4), h)
    big = sum(2 ** (32 * i) * v for i, v in enumerate(words))
    ints = []
    while big > 0:
        big, mod = divmod(big, 2 ** 32)
        ints.append(mod)
    rng = np.random.RandomState()
    rng.seed(ints or [0])
    return rng, seed
