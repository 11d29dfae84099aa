"""debug: replay the surface_follow parity test for one config and print where the images differ"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tactile_gym_b200 as tg
from oracle import oracle as O

sensor, S, movement = sys.argv[1], int(sys.argv[2]), sys.argv[3]
modes = {"movement_mode": movement, "control_mode": "TCP_velocity_control", "noise_mode": "simplex",
         "observation_mode": "tactile", "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": sensor}
n = 5
env = tg.make_vec("surface_follow-v0", n, env_kwargs={"env_modes": modes, "image_size": [S, S], "max_steps": 200})
rng = np.random.RandomState(S + len(sensor))
draws = np.stack([rng.randint(0, 10 ** 8, (n, 2)).astype(np.float64), rng.uniform(-np.pi, np.pi, (n, 2))], axis=2)
env.world.set_draws(draws)
obs = env.reset()["tactile"]
refs = []
for i in range(n):
    r = O.SurfaceFollowOracle(image_size=S, sensor=sensor, movement_mode=movement)
    r.reset(draws=(draws[i, 0, 0], draws[i, 0, 1]))
    refs.append(r)
act_dim = env.world.act_dim
for k in range(30):
    act = rng.uniform(-0.25, 0.25, (n, act_dim)).astype(np.float32)
    act[:, 0] = 0.25 if k < 12 else act[:, 0]
    o2, rew, done, infos = env.step(act)
    st = env.world.get_state()
    for i, r in enumerate(refs):
        for q in range(6):
            r.s.q[q] = st[i, q]; r.s.qd[q] = st[i, 6 + q]
        r.steps = int(st[i, 21])
        img = r.observation()[..., 0]
        g = o2["tactile"][i][..., 0]
        d = np.abs(img.astype(int) - g.astype(int))
        if d.max() > 1:
            rr, cc = np.nonzero(d > 1)
            print("step", k, "env", i, "bad px", len(rr), "rows", rr.min(), rr.max(), "cols", cc.min(), cc.max(), "pipeline_error", env.world.pipeline_error())
            for a, b in list(zip(rr, cc))[:12]:
                print("   px", a, b, "oracle", img[a, b], "gpu", g[a, b])
            # full-surface oracle (all cells) for comparison
            tw = O.heightfield_tris(r.V, 0, 63, 0, 63)
            full = O.tactile_image(r.m, np.array(r.s.q[:6]), S, tw, r.ref, border_on=True)
            print("   oracle(full surface) vs oracle(patch) max diff", np.abs(full.astype(int) - img.astype(int)).max(), " vs gpu", np.abs(full.astype(int) - g.astype(int)).max())
            sys.exit(0)
print("no mismatch")
