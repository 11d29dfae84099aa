"""where the time of one VecEnv.step (host numpy in, host numpy out) goes"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tactile_gym_b200 as tg
modes = {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "noise_mode": "rand_height",
         "observation_mode": "tactile", "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "tactip"}
n = 4096
env = tg.make_vec("edge_follow-v0", n, seed=1, env_kwargs={"env_modes": modes, "image_size": [128, 128], "max_steps": 200})
env.reset()
w = env.world
st = w.get_state(); st[:, 2 * w.nb + 9] = np.random.RandomState(0).randint(0, 200, size=n); w.set_state(st)
rng = np.random.RandomState(0)
acts = rng.uniform(-0.25, 0.25, (80, n, 2)).astype(np.float32)
for k in range(30): env.step(acts[k])
T = {"async_host": 0, "wait_sync": 0, "wait_host": 0, "total": 0}
for k in range(30, 80):
    t0 = time.perf_counter(); env.step_async(acts[k]); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    env.step_wait(); t3 = time.perf_counter()
    T["async_host"] += t1 - t0; T["wait_sync"] += t2 - t1; T["wait_host"] += t3 - t2; T["total"] += t3 - t0
print({k: "%.3f ms" % (v / 50 * 1e3) for k, v in T.items()}, "-> %.2f M steps/s" % (n * 50 / T["total"] / 1e6))
# device-side split
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
a = torch.from_numpy(acts[0]).cuda()
torch.cuda.synchronize()
ev[0].record(); w.step(a, want_terminal_obs=True); ev[1].record(); env._pin_obs.copy_(w.obs, non_blocking=True); ev[2].record(); torch.cuda.synchronize()
print("kernels %.3f ms, obs D2H %.3f ms" % (ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])))
