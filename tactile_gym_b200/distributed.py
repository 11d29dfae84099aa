"""Multi-GPU plumbing: one process per GPU, envs sharded by index range, no data-path collective.

SURVEY.md 8(e): envs never interact, so rank r of G owns global env indices [r*n_local, (r+1)*n_local) and the only collective
is the OPTIONAL all-gather of the step's results for callers that want one collated batch (what SubprocVecEnv's pipes do in the
reference, tactile_gym/sb3_helpers/rl_utils.py:17-30).  `CollatedBatch` does it with ONE `all_gather_into_tensor` per step on a
packed, pre-allocated buffer, IN PLACE: every rank's slot

    [ obs u8 n*S*S | reward f32 n | done u8 n (padded to 4) | feat f32 n*F ]

lives inside the receive buffer, the step / raster kernels write their outputs straight into it (the world's obs / reward / done /
feat tensors ARE views of the slot), so nothing is staged or copied before NCCL reads it.  Two such buffers alternate, and the
collective runs on a side stream behind an event, so the gather of step k overlaps the physics and raster of step k+1.
The same code runs over NCCL (GPU tensors) and gloo (CPU tensors, used by the world_size-2 tests).
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


def _align(x, a=16):
    return (x + a - 1) // a * a


class PackedSlot:
    """byte layout of one rank's slot"""

    def __init__(self, n, S, nfeat=0):
        self.n, self.S, self.nfeat = n, S, nfeat
        self.obs_off, self.obs_bytes = 0, n * S * S
        self.rew_off = _align(self.obs_bytes)
        self.done_off = _align(self.rew_off + 4 * n)
        self.feat_off = _align(self.done_off + n)
        self.bytes = _align(self.feat_off + 4 * n * nfeat)

    def views(self, slot):
        """typed views of a [bytes] uint8 tensor (one rank's slot)"""
        n, S = self.n, self.S
        obs = slot[self.obs_off:self.obs_off + self.obs_bytes].view(n, S, S, 1)
        rew = slot[self.rew_off:self.rew_off + 4 * n].view(torch.float32)
        done = slot[self.done_off:self.done_off + n]
        feat = slot[self.feat_off:self.feat_off + 4 * n * self.nfeat].view(torch.float32).view(n, self.nfeat) if self.nfeat else None
        return obs, rew, done, feat

    def global_views(self, buf, world):
        """views of the whole gathered buffer: obs [world, n, S, S, 1], reward [world, n], done [world, n], feat [world, n, F]
        (rank-major = global env index order: env g is [g // n, g % n]).  They are strided views into the packed buffer; use
        `.flatten(0, 1)` for a contiguous [world * n, ...] copy."""
        n, S = self.n, self.S
        b = buf.view(world, self.bytes)
        obs = b[:, self.obs_off:self.obs_off + self.obs_bytes].unflatten(1, (n, S, S, 1))
        rew = torch.as_strided(buf.view(torch.float32), (world, n), (self.bytes // 4, 1), self.rew_off // 4)
        done = b[:, self.done_off:self.done_off + n]
        feat = None
        if self.nfeat:
            feat = torch.as_strided(buf.view(torch.float32), (world, n, self.nfeat), (self.bytes // 4, self.nfeat, 1), self.feat_off // 4)
        return obs, rew, done, feat


class CollatedBatch:
    """One in-place packed all-gather per step, double-buffered, overlapped with the next step.

        cb = CollatedBatch(n_local, S, nfeat, device)            # after dist.init_process_group
        obs, rew, done, feat = cb.local_views()                  # hand these to the world as its output tensors
        ... step writes them ...
        work = cb.gather()                                       # async; flips to the other buffer for the next step
        g_obs, g_rew, g_done, g_feat = cb.wait(work)             # [world, n, ...] views, valid until the step after next
    """

    def __init__(self, n_local, S, nfeat=0, device="cpu", group=None, depth=2):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.slot = PackedSlot(n_local, S, nfeat)
        self.device = torch.device(device)
        self.bufs = [torch.zeros(self.world * self.slot.bytes, dtype=torch.uint8, device=self.device) for _ in range(depth)]
        self.cur = 0
        self.cuda = self.device.type == "cuda"
        self.side = torch.cuda.Stream(self.device) if self.cuda else None
        self.ready = [torch.cuda.Event() for _ in range(depth)] if self.cuda else None

    def _my_slot(self, k):
        b = self.slot.bytes
        return self.bufs[k][self.rank * b:(self.rank + 1) * b]

    def local_views(self, k=None):
        """this rank's output tensors inside buffer k (default: the one the next step should write)"""
        return self.slot.views(self._my_slot(self.cur if k is None else k))

    def gather(self):
        """start the all-gather of the current buffer (its local slot must have been written on the current stream) and
        flip; returns a handle for wait()"""
        k = self.cur
        self.cur = (self.cur + 1) % len(self.bufs)
        if self.world == 1:
            return (k, None, None)
        if self.cuda:
            ev = self.ready[k]
            ev.record(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.side):
                self.side.wait_event(ev)
                work = dist.all_gather_into_tensor(self.bufs[k], self._my_slot(k), group=self.group, async_op=True)
            return (k, work, None)
        work = dist.all_gather_into_tensor(self.bufs[k], self._my_slot(k), group=self.group, async_op=True)
        return (k, work, None)

    def wait(self, handle):
        """make the current stream wait for the gather; returns the global views of that buffer"""
        k, work, _ = handle
        if work is not None:
            work.wait()          # NCCL: stream-level wait (no host block); gloo: blocks
        return self.slot.global_views(self.bufs[k], self.world)


def all_gather_batch(obs, reward, done, group=None):
    """[n_local, ...] per rank -> [world * n_local, ...] on every rank, rank-major (= global env index order).  Convenience
    form for tensors that do not live in a CollatedBatch: packs them into one buffer, ONE collective, unpacks."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return obs, reward, done
    world = dist.get_world_size(group)
    parts = [t.contiguous().view(-1).view(torch.uint8) for t in (obs, reward, done)]
    sizes = [p.numel() for p in parts]
    offs = [0, _align(sizes[0]), _align(_align(sizes[0]) + sizes[1])]
    total = _align(offs[2] + sizes[2])
    send = torch.zeros(total, dtype=torch.uint8, device=obs.device)
    for p, o in zip(parts, offs):
        send[o:o + p.numel()] = p
    recv = torch.empty(world * total, dtype=torch.uint8, device=obs.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    r = recv.view(world, total)
    out = []
    for t, o, s in zip((obs, reward, done), offs, sizes):
        g = r[:, o:o + s].contiguous().view(-1).view(t.dtype)
        out.append(g.view((world * t.shape[0],) + tuple(t.shape[1:])))
    return tuple(out)


def max_over_ranks(values, device, group=None):
    """device-side max over ranks of a list of floats (kernel timings are reported as the slowest rank's)"""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return [float(x) for x in t.tolist()]
