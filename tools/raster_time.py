import sys; sys.path.insert(0,".")
import torch, bench, tactile_gym_b200 as tg
for name in ("edge","balance","push"):
    W=bench.workload(name)
    env=tg.make_vec(W["env_id"],W["n"],seed=1,env_kwargs={"env_modes":W["modes"],"image_size":[W["img"],W["img"]],"max_steps":W["max_steps"]}); env.reset(); w=env.world
    flush=torch.empty(256*1024*1024,dtype=torch.uint8,device="cuda")
    for _ in range(3): w.raster_only()
    ts=[]
    for k in range(20):
        flush.fill_(k); a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); a.record(); w.raster_only(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    t=sum(ts)/len(ts); print("raster_only %s %dx%d: %.4f ms  %.0f GB/s  frac %.3f"%(name,W["n"],W["img"],t,W["n"]*W["alg_bytes"]/t/1e6, W["n"]*W["alg_bytes"]/t/1e6/6530.3))
    env.close(); del flush
