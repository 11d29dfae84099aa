"""CPU, world_size 2, gloo: the N > 1 plumbing (index sharding, seeds, the optional all-gather, max-over-ranks)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world)})
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tactile_gym_b200 import distributed as D

    n_local, S = 3, 4
    lo, hi = D.shard_range(rank, world, n_local)
    seeds = D.shard_seeds(100, rank, n_local)
    obs = torch.full((n_local, S, S, 1), rank, dtype=torch.uint8) + torch.arange(n_local, dtype=torch.uint8).view(-1, 1, 1, 1) * 10
    rew = torch.arange(lo, hi, dtype=torch.float32)
    done = torch.tensor([(i % 2) for i in range(lo, hi)], dtype=torch.uint8)
    g_obs, g_rew, g_done = D.all_gather_batch(obs, rew, done)
    mx = D.max_over_ranks([1.0 + rank, 5.0 - rank], "cpu")
    # the packed, in-place, double-buffered gather: the "kernels" write straight into this rank's slot of the receive buffer
    cb = D.CollatedBatch(n_local, S, nfeat=2, device="cpu")
    packed = []
    for step in range(3):                       # 3 steps over 2 buffers: the flip and the reuse are exercised
        o, r, d, f = cb.local_views()
        o.copy_(obs + step); r.copy_(rew * (step + 1)); d.copy_(done); f.copy_(torch.stack([rew, -rew], dim=1) + step)
        go, gr, gd, gf = cb.wait(cb.gather())
        packed.append((tuple(go.shape), go[:, :, 0, 0, 0].flatten().tolist(), gr.flatten().tolist(), gd.flatten().tolist(), gf.flatten(0, 1)[:, 1].tolist()))
    q.put((rank, lo, hi, seeds, g_obs.shape, g_rew.tolist(), g_done.tolist(), g_obs[:, 0, 0, 0].tolist(), mx, packed))
    dist.destroy_process_group()


def test_sharding_and_gather_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, s0, shp0, rew0, done0, o0, mx0, pk0), (r1, lo1, hi1, s1, shp1, rew1, done1, o1, mx1, pk1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 3, 3, 6)
    assert s0 + s1 == [100, 101, 102, 103, 104, 105]          # seeds follow the global env index
    assert tuple(shp0) == (6, 4, 4, 1)
    assert rew0 == rew1 == [0.0, 1.0, 2.0, 3.0, 4.0, 5.0]      # rank-major == global index order
    assert done0 == done1 == [0, 1, 0, 1, 0, 1]
    assert o0 == o1 == [0, 10, 20, 1, 11, 21]
    assert mx0 == mx1 == [2.0, 5.0]
    assert pk0 == pk1
    for step, (shp, o, r, d, f1) in enumerate(pk0):
        assert shp == (2, 3, 4, 4, 1)                          # [world, n_local, S, S, 1]
        assert o == [x + step for x in [0, 10, 20, 1, 11, 21]]
        assert r == [x * (step + 1) for x in [0.0, 1.0, 2.0, 3.0, 4.0, 5.0]] and d == [0, 1, 0, 1, 0, 1]
        assert f1 == [-x + step for x in [0.0, 1.0, 2.0, 3.0, 4.0, 5.0]]


def test_single_process_is_a_no_op():
    from tactile_gym_b200 import distributed as D

    o, r, d = torch.zeros(2, 4, 4, 1), torch.zeros(2), torch.zeros(2)
    assert D.all_gather_batch(o, r, d)[0] is o
    assert D.max_over_ranks([3.0], "cpu") == [3.0]
    cb = D.CollatedBatch(2, 4, nfeat=0, device="cpu")
    o, r, d, f = cb.local_views()
    o.fill_(7); r.fill_(1.5); d.fill_(1)
    go, gr, gd, gf = cb.wait(cb.gather())
    assert go.shape == (1, 2, 4, 4, 1) and int(go.sum()) == 7 * 32 and gr.tolist() == [[1.5, 1.5]] and gd.tolist() == [[1, 1]] and gf is None
    assert cb.local_views()[0].data_ptr() != o.data_ptr()      # flipped to the other buffer
