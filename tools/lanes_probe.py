"""step_kernel time against the lane packing (envs per warp) at N = 4096: is one warp per scheduler really the best point?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, tactile_gym_b200 as tg
n = int(os.environ.get("TG_PROBE_N", "4096"))
flush = None
W = bench.workload(os.environ.get("TG_PROBE_WORKLOAD", "edge"))
for lanes in [int(x) for x in sys.argv[1:]] or [-8, 8, 16, 32]:      # -8: one env per 8-lane group (tg_g8.cuh)
    env = tg.make_vec(W["env_id"], n, env_kwargs={"env_modes": W["modes"], "image_size": [W["img"], W["img"]], "max_steps": W["max_steps"]}, lanes_per_warp=lanes)
    env.world.seed([1 + i for i in range(n)]); env.reset(); w = env.world
    if flush is None:
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=w.device)
    g0 = torch.Generator(device=w.device); g0.manual_seed(1)
    acts = (torch.rand((40, n, w.act_dim), device=w.device, generator=g0) - 0.5) * 0.5
    for k in range(5): w.physics_only(acts[k])
    cold, warm = [], []
    for k in range(5, 25):
        flush.fill_(k & 255); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); w.physics_only(acts[k]); b.record(); torch.cuda.synchronize(); cold.append(a.elapsed_time(b))
    for k in range(25, 40):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); w.physics_only(acts[k]); b.record(); torch.cuda.synchronize(); warm.append(a.elapsed_time(b))
    print("lanes_per_warp %2d: step_kernel cold-L2 %.4f ms, warm %.4f ms" % (lanes, np.mean(cold), np.mean(warm)))
    env.close()
