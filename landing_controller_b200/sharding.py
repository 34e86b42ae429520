"""Scenario sharding for multi-GPU sweeps: contiguous blocks of the sweep per rank,
no data-path collective, ONE all-gather of the per-scenario result records at the end."""
import numpy as np


def shard_bounds(n_scenarios, world, rank):
    """Contiguous block [lo, hi) of rank `rank`; blocks differ by at most one scenario."""
    base, rem = divmod(n_scenarios, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_records(x, f, status, iters):
    """[B, nx+3] float64 record per scenario: x*, f*, status, iters (what the all-gather moves)."""
    B = x.shape[0]
    rec = np.empty((B, x.shape[1] + 3))
    rec[:, :-3] = x
    rec[:, -3] = f
    rec[:, -2] = status
    rec[:, -1] = iters
    return rec


def unpack_records(rec):
    return dict(x=rec[:, :-3], f=rec[:, -3], status=rec[:, -2].astype(np.int32), iters=rec[:, -1].astype(np.int32))


def gather_records(rec_local, n_scenarios, world, rank):
    """All-gather ragged shards with torch.distributed (any backend): pads to the largest shard."""
    import torch
    import torch.distributed as dist
    sizes = [shard_bounds(n_scenarios, world, r) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    dev = rec_local.device if isinstance(rec_local, torch.Tensor) else None
    t = rec_local if isinstance(rec_local, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(rec_local))
    pad = torch.zeros(mx, t.shape[1], dtype=t.dtype, device=dev)
    pad[:t.shape[0]] = t
    out = torch.zeros(world * mx, t.shape[1], dtype=t.dtype, device=dev)
    dist.all_gather_into_tensor(out, pad)
    parts = [out[r * mx:r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts, dim=0)
