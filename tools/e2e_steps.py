"""per-step wall times of the e2e loop exactly as bench.py runs it (where do the slow steps come from?)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import tactile_gym_b200 as tg
n = 4096
chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 0
W = bench.workload("edge")
env = tg.make_vec(W["env_id"], n, env_kwargs={"env_modes": W["modes"], "image_size": [128, 128], "max_steps": 200}, copy_chunks=chunks)
env.world.seed([1 + i for i in range(n)])
env.reset()
w = env.world
st = w.get_state(); st[:, 2 * w.nb + 9] = np.random.RandomState(1000).randint(0, 200, size=n); w.set_state(st)
g0 = torch.Generator(device=w.device); g0.manual_seed(12345)
for _ in range(40):
    w.step((torch.rand((n, 2), device=w.device, generator=g0) - 0.5) * 0.5)
torch.cuda.synchronize()
acts = np.random.RandomState(0).uniform(-0.25, 0.25, (120, n, 2)).astype(np.float32)
for k in range(10): env.step(acts[k])
ts, ta, tw, nd, tsy = [], [], [], [], []
import gc
gc0 = [g["collections"] for g in gc.get_stats()]
for k in range(10, 110):
    t0 = time.perf_counter(); env.step_async(acts[k]); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    o, r, d, i = env.step_wait(); t3 = time.perf_counter()
    ts.append(t3 - t0); ta.append(t1 - t0); tw.append(t3 - t2); nd.append(int(d.sum())); tsy.append(t2 - t1)
ts = np.array(ts) * 1e3
print("chunks", chunks, "mean %.3f median %.3f p90 %.3f max %.3f ms -> %.2f M steps/s" % (ts.mean(), np.median(ts), np.percentile(ts, 90), ts.max(), n / ts.mean() / 1e3))
print("async mean %.3f, wait_host mean %.3f, done/step mean %.1f" % (np.mean(ta) * 1e3, np.mean(tw) * 1e3, np.mean(nd)))
print("slowest (idx, total, async, sync, wait_host, ndone):", [(int(k), round(float(ts[k]), 2), round(ta[k] * 1e3, 2), round(tsy[k] * 1e3, 2), round(tw[k] * 1e3, 2), nd[k]) for k in np.argsort(ts)[-4:]])
print("gc collections during loop:", [g["collections"] - a for g, a in zip(gc.get_stats(), gc0)], "pipeline stalls", w.pipeline_stalls())
