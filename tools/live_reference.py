"""Locate and drive the UNMODIFIED reference (ac-93/tactile_gym on PyBullet) when a box happens to have it.

Neither this container nor the GPU pool has a pybullet wheel, so in practice `probe()` answers "unavailable" and the callers
(tests/test_live_pybullet.py, `bench.py --impl reference`) skip / fall back to the CPU oracle port.  The hook exists so that a pod
WITH the wheel is noticed ("probe at runtime, never assume"): the reference package itself travels with the repo in
baseline/_ref (pip --no-deps install of /root/reference, git-ignored), only its dependencies are missing.
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = [os.path.join(ROOT, "baseline", "_ref"), "/root/reference"]


def probe():
    """-> (tactile_gym module, pybullet module, None) or (None, None, reason)"""
    try:
        pb = importlib.import_module("pybullet")
    except Exception as e:  # noqa: BLE001
        return None, None, "pybullet not importable (%s)" % type(e).__name__
    try:
        importlib.import_module("gym")
    except Exception:  # noqa: BLE001
        return None, None, "gym not importable (the reference's envs subclass gym.Env and use gym <= 0.21 seeding)"
    for c in CANDIDATES:
        if os.path.isdir(os.path.join(c, "tactile_gym", "assets")):
            if c not in sys.path:
                sys.path.insert(0, c)
            try:
                return importlib.import_module("tactile_gym"), pb, None
            except Exception as e:  # noqa: BLE001
                return None, None, "tactile_gym found in %s but not importable: %r" % (c, e)
    return None, None, "tactile_gym package (baseline/_ref or /root/reference) not found"


def make_env(env_id, env_modes, image_size, max_steps):
    """The reference's own constructor through its own registry (tactile_gym/rl_envs/__init__.py:3-41), DIRECT mode, no GUI."""
    import gym

    importlib.import_module("tactile_gym.rl_envs")
    return gym.make(env_id, max_steps=max_steps, image_size=list(image_size), env_modes=dict(env_modes), show_gui=False, show_tactile=False)


def asset_path(*parts):
    for c in CANDIDATES:
        p = os.path.join(c, "tactile_gym", "assets", *parts)
        if os.path.exists(p):
            return p
    return None
