"""Host-side sharding of independent streams over the GPUs of one box.

Channels (and FFT blocks) are independent, so each rank owns a contiguous channel range and
there is no data-path collective; torch.distributed is used only for the barrier and for
reducing the timings (max over ranks) and unit counts (sum over ranks).  Backend agnostic:
NCCL on the GPU box, gloo in the CPU tests."""
from __future__ import annotations

import numpy as np


def partition(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced [start, start+count) of `total` items for `rank` of `world`."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def channel_tuning(first: int, count: int, lo: float = 2000.0, hi: float = 90000.0, seed: int = 7) -> np.ndarray:
    """Per-channel tuning of BASELINE config 4 (uniform in [lo, hi], seed 7) addressed by the
    GLOBAL channel index, so that a channel's tuning does not depend on how many ranks there are."""
    out = np.empty(count, dtype=np.float64)
    for i in range(count):
        rng = np.random.Generator(np.random.PCG64([seed, first + i]))
        out[i] = rng.uniform(lo, hi)
    return out


def reduce_timing(dist, device, ms_local: list[float], units_local: float) -> tuple[list[float], float]:
    """(max over ranks of each time, sum over ranks of the units).  `dist` is the initialised
    torch.distributed module or None for a single process."""
    if dist is None:
        return list(ms_local), float(units_local)
    import torch
    t = torch.tensor(list(ms_local), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    u = torch.tensor([float(units_local)], dtype=torch.float64, device=device)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()], float(u.item())
