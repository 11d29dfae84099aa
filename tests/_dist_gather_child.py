"""torchrun child of tests/test_gpu_distributed.py: every rank steps its shard of the envs with the collated-batch all-gather
(NCCL) and checks the gathered batch, in global env order, against a single-GPU world holding ALL the envs on this rank's GPU."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import tactile_gym_b200 as tg

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
task = sys.argv[1] if len(sys.argv) > 1 else "edge"
n, S = 24, 64
if task == "push":
    env_id = "object_push-v0"
    modes = {"movement_mode": "TyRz", "control_mode": "TCP_velocity_control", "rand_init_orn": False, "rand_obj_mass": False, "traj_type": "simplex",
             "observation_mode": "tactile_and_feature", "reward_mode": "dense", "arm_type": "mg400", "tactile_sensor_name": "digitac"}
    nd, act_dim, ms = 3, 2, 5
else:
    env_id = "edge_follow-v0"
    modes = {"movement_mode": "xy", "control_mode": "TCP_velocity_control", "noise_mode": "rand_height", "observation_mode": "tactile",
             "reward_mode": "dense", "arm_type": "ur5", "tactile_sensor_name": "tactip"}
    nd, act_dim, ms = 2, 2, 4
kw = {"env_modes": modes, "image_size": [S, S], "max_steps": ms}
rng = np.random.RandomState(0)            # the same global draws / actions on every rank
G = n * world
if task == "push":
    draws = np.stack([np.zeros((G, 6)), np.full((G, 6), 0.491), rng.randint(0, 10 ** 8, (G, 6)).astype(np.float64)], axis=2)
else:
    draws = np.stack([rng.uniform(0.0015, 0.0065, (G, 6)), rng.uniform(-np.pi, np.pi, (G, 6))], axis=2)
acts = torch.tensor(rng.uniform(-0.25, 0.25, (9, G, act_dim)).astype(np.float32), device=dev)

shard = tg.make_vec(env_id, n, env_kwargs=kw, device=local)
shard.world.set_draws(draws[rank * n:(rank + 1) * n])
shard.reset()
full = tg.make_vec(env_id, G, env_kwargs=kw, device=local)
full.world.set_draws(draws)
full.reset()
prev = None
any_done = False
for k in range(9):
    h = shard.step_collated(acts[k, rank * n:(rank + 1) * n])
    o, r, d = full.world.step(acts[k])
    want = (o.clone(), r.clone(), d.clone(), full.world.feat.clone() if full.world.nfeat else None)
    any_done = any_done or bool(d.any())
    if prev is not None:
        # consume batch k-1 while step k is in flight (the overlap the double buffer exists for)
        go, gr, gd, gf = shard.collated_wait(prev[0])
        wo, wr, wd_, wf = prev[1]
        assert go.shape == (world, n, S, S, 1)
        assert torch.equal(go.flatten(0, 1), wo), ("obs", k - 1)
        assert torch.equal(gr.flatten(), wr) and torch.equal(gd.flatten(), wd_), ("reward/done", k - 1)
        if wf is not None:
            assert torch.equal(gf.flatten(0, 1), wf), ("feat", k - 1)
    prev = (h, want)
go, gr, gd, gf = shard.collated_wait(prev[0])
assert torch.equal(go.flatten(0, 1), prev[1][0]) and torch.equal(gr.flatten(), prev[1][1])
assert any_done                                   # episode ends went through the gather too
torch.cuda.synchronize(dev)
shard.close(); full.close()
dist.barrier()
dist.destroy_process_group()
print("GATHER-OK rank %d" % rank)
