import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import tactile_gym_b200 as tg
modes = {"movement_mode": "xyzRxRy", "control_mode": "TCP_velocity_control", "noise_mode": "simplex", "observation_mode": "tactile",
         "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "digit"}
n = 1024
env = tg.make_vec("surface_follow-v0", n, env_kwargs={"env_modes": modes, "image_size": [128, 128], "max_steps": 200})
w = env.world
env.world.seed([1 + i for i in range(n)])
env.reset()
def run(label, K=60):
    gen = torch.Generator(device="cuda"); gen.manual_seed(0)
    acts = (torch.rand((K, n, 3), device="cuda", generator=gen) - 0.5) * 0.5
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    s0 = w.pipeline_stalls()
    for k in range(K):
        ev[k][0].record(); w.physics_only(acts[k]) if False else None
        w.step(acts[k]); ev[k][2].record()
    torch.cuda.synchronize()
    ts = np.array([a.elapsed_time(c) for a, b, c in ev])
    print(label, "step ms: median %.3f mean %.3f max %.3f | stalls %d" % (np.median(ts), ts.mean(), ts.max(), w.pipeline_stalls() - s0), "first 12:", np.round(ts[:12], 2))
run("sync phases")
st = w.get_state(); st[:, 2 * w.nb + 9] = np.random.RandomState(0).randint(0, 200, size=n); w.set_state(st)
run("staggered (first 60 after staggering)")
run("staggered (next 60)")
run("staggered (next 60)")
