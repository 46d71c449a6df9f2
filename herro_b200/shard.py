"""Read-id sharding across the GPUs of one box.

Target reads are independent units: every window of a read depends only on that read's
alignments and on read-only access to the read store.  The reference pulls targets from one MPMC
channel per process (src/lib.rs:154-187) and scales across processes with read clusters
(`-c`, scripts/create_clusters.py).  Here each rank owns a contiguous chunk of targets balanced by
window count; there is NO collective on the data path — only the final throughput reduction."""
from __future__ import annotations

import numpy as np


def shard_targets(read_len: np.ndarray, window_size: int, rank: int, world: int) -> np.ndarray:
    """Contiguous rid ranges with (nearly) equal numbers of windows.  Returns the rids of `rank`."""
    nwin = (np.asarray(read_len, dtype=np.int64) + window_size - 1) // window_size
    csum = np.cumsum(nwin)
    total = int(csum[-1]) if len(csum) else 0
    bounds = [int(np.searchsorted(csum, total * r / world, side="left")) for r in range(world)] + [len(nwin)]
    bounds[0] = 0
    return np.arange(bounds[rank], bounds[rank + 1], dtype=np.int64)


def reduce_throughput(dist, seconds: float, units: float, device=None):
    """(max seconds over ranks, sum of units) — the only cross-rank traffic of a run."""
    import torch
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    u = torch.tensor([units], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t[0]), float(u[0])
