"""Multi-GPU plumbing: one process per GPU, envs sharded by index range, no data-path collective.

SURVEY.md 8(e): envs never interact, so rank r of G owns global env indices [r*n_local, (r+1)*n_local) and the only
collective is the OPTIONAL all-gather of (obs, reward, done) for callers that want one collated batch.  The same code
runs over NCCL (GPU tensors) and gloo (CPU tensors, used by the world_size-2 tests).
"""
import torch
import torch.distributed as dist


def shard_range(rank, world, n_local):
    """global env index range owned by `rank`"""
    return rank * n_local, (rank + 1) * n_local


def shard_seeds(base_seed, rank, n_local):
    """per-env seeds that depend on the GLOBAL env index only, so results do not depend on the GPU count
    (the make_vec_env convention is seed + env index, sb3_helpers/rl_utils.py:17-30)"""
    lo, hi = shard_range(rank, 0, n_local)
    return [base_seed + i for i in range(lo, hi)]


def all_gather_batch(obs, reward, done, group=None):
    """[n_local, ...] per rank -> [world * n_local, ...] on every rank, rank-major (= global env index order)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return obs, reward, done
    world = dist.get_world_size(group)
    out = []
    for t in (obs, reward, done):
        t = t.contiguous()
        g = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(g, t, group=group)
        out.append(g)
    return tuple(out)


def max_over_ranks(values, device, group=None):
    """device-side max over ranks of a list of floats (kernel timings are reported as the slowest rank's)"""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return [float(x) for x in t.tolist()]
