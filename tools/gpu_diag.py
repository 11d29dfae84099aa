"""Diagnostics on the GPU box: parity drift statistics + raw kernel timings."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tactile_gym_b200 as tg
from oracle import oracle as O

modes = {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "noise_mode": "rand_height",
         "observation_mode": "tactile", "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "tactip"}

def drift():
    n, S = 8, 64
    env = tg.make_vec("edge_follow-v0", n, env_kwargs={"env_modes": modes, "image_size": [S, S], "max_steps": 200})
    rng = np.random.RandomState(0)
    draws = np.stack([rng.uniform(0.0015, 0.0065, (n, 2)), rng.uniform(-np.pi, np.pi, (n, 2))], axis=2)
    env.world.set_draws(draws); env.reset()
    refs = [O.EdgeFollowOracle(image_size=S) for _ in range(n)]
    for i, r in enumerate(refs): r.reset(draws=tuple(draws[i, 0]))
    st = env.world.get_state()
    print("reset dq", max(np.abs(st[i, :6] - np.array(r.s.q[:6])).max() for i, r in enumerate(refs)))
    for k in range(30):
        act = rng.uniform(-0.25, 0.25, (n, 2)).astype(np.float32)
        env.step(act)
        for i, r in enumerate(refs): r.step(act[i])
        st = env.world.get_state()
        if k % 5 == 0 or k == 29:
            dq = max(np.abs(st[i, :6] - np.array(r.s.q[:6])).max() for i, r in enumerate(refs))
            dqd = max(np.abs(st[i, 6:12] - np.array(r.s.qd[:6])).max() for i, r in enumerate(refs))
            dp = max(np.abs(st[i, 12:15] - r.tcp_world()[0]).max() for i, r in enumerate(refs))
            print("step", k, "dq %.2e dqd %.2e dtcp %.2e" % (dq, dqd, dp))
    env.close()

def timing(n=4096, S=128, lanes=0, iters=20):
    env = tg.make_vec("edge_follow-v0", n, env_kwargs={"env_modes": modes, "image_size": [S, S], "max_steps": 200}, lanes_per_warp=lanes)
    env.seed(1); env.reset()
    w = env.world
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    acts = (torch.rand((iters + 5, n, 2), device="cuda", generator=g) - 0.5) * 0.5
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for k in range(5): w.step(acts[k])
    torch.cuda.synchronize()
    def t(fn, reps):
        torch.cuda.synchronize(); ev[0].record()
        for k in range(reps): fn(k)
        ev[1].record(); torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]) / reps
    tp = t(lambda k: w.physics_only(acts[k % iters]), iters)
    tr = t(lambda k: w.raster_only(), iters)
    ts = t(lambda k: w.step(acts[k % iters]), iters)
    full = torch.ones(n, dtype=torch.uint8, device="cuda")
    trs = t(lambda k: tg._lib.check(w.lib.tg_reset_only(w.h, full.data_ptr(), w._stream())), 3)
    st = w.get_state()
    print("N=%d S=%d lanes=%d: physics %.3f ms, raster %.3f ms (%.0f GB/s), full step %.3f ms -> %.0f steps/s; reset-all %.3f ms (mean substeps %.1f)"
          % (n, S, w.cfg.lanes_per_warp if lanes else -1, tp, tr, n * S * S / tr / 1e6, ts, n / ts * 1e3, trs, st[:, 22].mean()))
    env.close()

if __name__ == "__main__":
    drift()
    for lanes in (1, 2, 4, 8, 32):
        timing(4096, 128, lanes)
    timing(16384, 128, 32)
    timing(4096, 64, 0)
    timing(1024, 256, 0)
