"""object_balance 2048 x 256^2: how many envs does the scanline raster hand to the general kernel as the episodes go on, and what
does a raster pass cost then?"""
import sys; sys.path.insert(0, ".")
import torch, bench, tactile_gym_b200 as tg
W = bench.workload("balance")
env = tg.make_vec(W["env_id"], W["n"], seed=1, env_kwargs={"env_modes": W["modes"], "image_size": [W["img"], W["img"]], "max_steps": W["max_steps"]})
env.reset(); w = env.world
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
g = torch.Generator(device="cuda"); g.manual_seed(0)
def t_raster():
    ts = []
    for k in range(8):
        flush.fill_(k); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); a.record(); w.raster_only(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sum(ts) / len(ts)
print("after reset: raster %.4f ms, fallbacks %d" % (t_raster(), w.scan_fallbacks()))
for rnd in range(5):
    for k in range(50):
        w.step((torch.rand((W["n"], w.act_dim), device="cuda", generator=g) - 0.5) * 0.5)
    print("after %d steps: raster %.4f ms, fallbacks %d of %d" % (50 * (rnd + 1), t_raster(), w.scan_fallbacks(), W["n"]), "reasons", w.scan_fallback_reasons().tolist())
