"""Multi-GPU plumbing for the ray-rendering path: one process per GPU (torch.distributed).

Rays are independent units, so the path shards with NO data-path collective:
rank r renders the contiguous ray range ``shard_bounds(n, r, world)``; random draws made on the
device are keyed by the GLOBAL ray index (``ray_id_offset``), so an N-GPU job is numerically the
same job as the 1-GPU one.  The only exchange step is the training gradient all-reduce: both
networks' gradients are packed into ONE flat fp32 buffer (2 x 595 844 floats = 4.77 MB) and reduced
with a single NCCL all-reduce per step on the compute stream (``allreduce_gradients``).
``gather_rays`` optionally reassembles rendered outputs on every rank (52 B/ray).

Backend: "nccl" on GPUs (NVLink 5 / NVSwitch), "gloo" in the CPU tests (tests/test_dist_gloo.py).
"""
import torch
import torch.distributed as dist


def is_initialized():
    return dist.is_available() and dist.is_initialized()


def world():
    return (dist.get_rank(), dist.get_world_size()) if is_initialized() else (0, 1)


def shard_bounds(n, rank=None, world_size=None):
    """Contiguous, balanced split of n rays: the first n % world ranks get one extra ray."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(rays_flat, rank=None, world_size=None):
    """-> (local rays, global index of local ray 0).  rays_flat [n, 8|11] (already packed)."""
    lo, hi = shard_bounds(rays_flat.shape[0], rank, world_size)
    return rays_flat[lo:hi], lo


def render_sharded(render_rays_fn, rays_flat, chunk=1024 * 32, **kwargs):
    """batchify_rays over this rank's shard; returns the local dict of outputs.  Device-side draws
    stay keyed by global ray id through ``ray_id_offset``."""
    local, lo = shard_rays(rays_flat)
    out = {}
    for i in range(0, local.shape[0], chunk):
        r = render_rays_fn(local[i:i + chunk], ray_id_offset=lo + i, **kwargs)
        for k, v in r.items():
            out.setdefault(k, []).append(v)
    return {k: torch.cat(v, 0) for k, v in out.items()}


def gather_rays(local, n_total):
    """All-gather a per-ray tensor [n_local, ...] back into global ray order [n_total, ...]."""
    rank, ws = world()
    if ws == 1:
        return local
    sizes = [shard_bounds(n_total, r, ws) for r in range(ws)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:hi - lo] for b, (lo, hi) in zip(bufs, sizes)], 0)


class FlatGradBucket:
    """One flat fp32 buffer aliasing the .grad of every parameter of the given modules, so that the
    data-parallel reduction is a single all-reduce (never one per layer or per network)."""

    def __init__(self, modules, extra=0):
        self.params = [p for m in modules if m is not None for p in m.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        # `extra` floats behind the gradients are cleared by the same fill but stay out of the all-reduce
        # (train.TrainStep keeps its loss accumulators there)
        self._store = torch.zeros(n + extra, dtype=torch.float32, device=dev)
        self.flat, self.extra = self._store[:n], self._store[n:]
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def zero_(self):
        self._store.zero_()

    def allreduce_mean(self):
        """SUM over ranks then divide by world: the loss is a mean over the global ray batch."""
        rank, ws = world()
        if ws > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(ws)
        return self.flat

    def allreduce_sum(self):
        """SUM over ranks, no division: for steps whose local loss gradient is already scaled by the GLOBAL batch
        size (train.TrainStep)."""
        rank, ws = world()
        if ws > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        return self.flat


def allreduce_gradients(bucket):
    return bucket.allreduce_mean()
