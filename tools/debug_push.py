"""GPU-side debugging aid: step object_push on the device and in the oracle from identical states, print the differences."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tactile_gym_b200 as tg
from oracle import oracle as O
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_push import PUSH_MODES, _sync, _set_goal

arm, sensor, S, movement = "mg400", "digitac", 128, "TyRz"
n, nb = 4, 8
env = tg.make_vec("object_push-v0", n, env_kwargs={"env_modes": PUSH_MODES, "image_size": [S, S], "max_steps": 1000})
rng = np.random.RandomState(0)
draws = np.stack([np.zeros((n, 2)), np.full((n, 2), 0.491), rng.randint(0, 10 ** 8, (n, 2)).astype(np.float64)], axis=2)
env.world.set_draws(draws)
env.reset()
st = env.world.get_state()
refs = []
for i in range(n):
    r = O.ObjectPushOracle(image_size=S)
    r.reset(draws=draws[i, 0]); refs.append(r)
    _set_goal(r, st[i, 2 * nb + 25]); _sync(r, st[i], nb)
for k in range(int(sys.argv[1]) if len(sys.argv) > 1 else 12):
    act = rng.uniform(-0.25, 0.25, (n, 2)).astype(np.float32)
    o2, rew, done, infos = env.step(act)
    st = env.world.get_state()
    for i, r in enumerate(refs):
        o, rr, dd, _ = r.step(act[i])
        ob = st[i, 2 * nb + 11:]
        print(k, i, "dq %.1e dpos %.1e dquat %.1e dvel %.1e drew %.1e goal %d/%d nc %d it %d" % (
            np.abs(st[i, :nb] - np.array(r.s.q[:nb])).max(), np.abs(ob[:3] - np.array(r.o.pos[:])).max(),
            np.abs(ob[3:7] - np.array(r.o.quat[:])).max(), np.abs(ob[7:13] - np.array(list(r.o.vel[:]) + list(r.o.omg[:]))).max(),
            abs(rr - rew[i]), int(ob[14]), r.targ, r.p.n_contacts, r.p.n_iters), "quat", ob[3:7], np.array(r.o.quat[:]))
        _sync(r, st[i], nb)
