#!/usr/bin/env python3
"""Golden trajectories of the CPU oracle (tests/golden/oracle_trajectories.npz): for each task, seeded env + seeded actions ->
joint states, rewards, dones and the tactile images' digests over a reset and 8 steps.  tests/test_oracle_golden.py replays them,
so an edit of oracle/ that changes an existing behaviour shows up without a GPU.  The committed file was generated from the
oracle of commit 2581c9a (before the oracle-obs / sparse / position-control additions), i.e. it also shows those additions left
the BASELINE configurations' arithmetic untouched.
usage: make_oracle_golden.py [repo root to import oracle/ from] [output.npz]"""
import os
import sys
import zlib

root = os.path.abspath(sys.argv[1]) if len(sys.argv) > 1 else os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import numpy as np  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = [
    ("edge", lambda: O.EdgeFollowOracle(image_size=64, seed=11), 2),
    ("edge_mg400_digitac", lambda: O.EdgeFollowOracle(image_size=64, arm="mg400", sensor="digitac", seed=12), 2),
    ("balance", lambda: O.ObjectBalanceOracle(image_size=64, seed=13), 2),
    ("surface", lambda: O.SurfaceFollowOracle(image_size=64, sensor="digit", seed=14), 3),
    ("surface_goal", lambda: O.SurfaceFollowOracle(image_size=64, sensor="tactip", seed=15, variant="goal"), 5),
    ("push", lambda: O.ObjectPushOracle(image_size=64, seed=16), 2),
    ("roll", lambda: O.ObjectRollOracle(image_size=64, seed=17, rand_obj_size=True, rand_embed_dist=True, rand_init_obj_pos=True), 2),
]


def digest(obs):
    img = obs["tactile"] if isinstance(obs, dict) else obs
    img = np.ascontiguousarray(img)
    return np.array([zlib.crc32(img.tobytes()), int(img.sum()), int((img > 0).sum())], dtype=np.int64)


def run(make, act_dim, steps=8):
    env = make()
    o = env.reset()
    n = env.m.ndof
    q, rew, done, dig = [np.array(env.s.q[:n])], [float(env.reward)], [bool(env.done)], [digest(o)]
    rng = np.random.RandomState(act_dim * 7 + 1)
    for k in range(steps):
        a = rng.uniform(-0.25, 0.25, act_dim).astype(np.float32)
        o, r, d, _ = env.step(a)
        q.append(np.array(env.s.q[:n])); rew.append(float(r)); done.append(bool(d)); dig.append(digest(o))
    return np.array(q), np.array(rew), np.array(done), np.array(dig)


def main():
    O.build()
    out = {}
    for name, make, act_dim in CASES:
        q, rew, done, dig = run(make, act_dim)
        out[name + "_q"], out[name + "_reward"], out[name + "_done"], out[name + "_image"] = q, rew, done, dig
    path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "oracle_trajectories.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.endswith("_q")})


if __name__ == "__main__":
    main()
