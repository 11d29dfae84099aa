"""object_balance on the g8 kernel vs the one-thread kernel (TG_G8=0/1 per process): states after 30 steps and timing"""
import os, sys; sys.path.insert(0, ".")
import numpy as np, torch, bench, tactile_gym_b200 as tg
W = bench.workload("balance")
n = int(sys.argv[1]) if len(sys.argv) > 1 else W["n"]
env = tg.make_vec(W["env_id"], n, seed=1, env_kwargs={"env_modes": W["modes"], "image_size": [64, 64], "max_steps": W["max_steps"]})
env.reset(); w = env.world
g = torch.Generator(device="cuda"); g.manual_seed(0)
acts = (torch.rand((40, n, w.act_dim), device="cuda", generator=g) - 0.5) * 0.5
for k in range(30): w.physics_only(acts[k])
st = w.get_state()
np.save(sys.argv[2], st)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
ts = []
for k in range(10):
    flush.fill_(k); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); a.record(); w.physics_only(acts[30 + k]); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print("TG_G8=%s n=%d physics %.4f ms (L2 flushed)" % (os.environ.get("TG_G8"), n, sum(ts) / len(ts)), "finite", np.isfinite(st).all())
