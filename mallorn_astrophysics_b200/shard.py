"""Object sharding across the GPUs of one box and the single all-gather of the feature table.

Objects are independent: each rank packs and featurises a contiguous range of the
dense object index, chosen so the ranges carry equal *cost* (observation counts by default; the GP pass
costs ~N^3 per object, so ``cost='cubic'`` balances that instead).  Nothing is exchanged during compute;
at the end one ``all_gather_into_tensor`` of the padded ``[rows_max, n_cols]`` float64 table (NCCL over
NVLink on the B200 box, gloo in the CPU tests) gives every rank the full table in object order.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def balanced_ranges(weights: Sequence[float], world: int) -> List[Tuple[int, int]]:
    """Split ``range(len(weights))`` into ``world`` contiguous ranges of (nearly) equal total weight.
    Every object lands in exactly one range; ranges may be empty when there are fewer objects than ranks."""
    w = np.asarray(weights, dtype=np.float64)
    n = int(w.shape[0])
    if world <= 0:
        raise ValueError("world must be positive")
    if n == 0:
        return [(0, 0)] * world
    csum = np.cumsum(w)
    total = float(csum[-1])
    if total <= 0.0:                                   # degenerate weights: split by count
        cuts = [round(n * r / world) for r in range(world + 1)]
    else:
        targets = total * np.arange(1, world) / world
        inner = np.searchsorted(csum, targets, side="left") + 1
        cuts = [0] + [int(min(max(c, 0), n)) for c in inner] + [n]
        for i in range(1, len(cuts)):                  # monotone
            cuts[i] = max(cuts[i], cuts[i - 1])
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def object_costs(obj: np.ndarray, n_obj: int, cost: str = "linear") -> np.ndarray:
    """Per-object cost from the flat ``obj`` column: observation count (``linear``) or its cube (``cubic``)."""
    cnt = np.bincount(obj[obj >= 0], minlength=n_obj).astype(np.float64)
    if cost == "linear":
        return cnt
    if cost == "cubic":
        return cnt ** 3
    raise ValueError(cost)


def shard_rows(obj: np.ndarray, lo: int, hi: int):
    """Row mask and re-based object index of the shard ``[lo, hi)``."""
    keep = (obj >= lo) & (obj < hi)
    return keep, (obj[keep] - lo).astype(np.int32)


def allgather_rows(local, rows_per_rank: Sequence[int], group=None):
    """All-gather row blocks of unequal height.  ``local``: torch tensor ``[rows_per_rank[rank], C]`` (CPU for
    gloo, CUDA for NCCL).  Returns ``[sum(rows_per_rank), C]`` in rank order on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if len(rows_per_rank) != world or int(rows_per_rank[rank]) != int(local.shape[0]):
        raise ValueError("rows_per_rank does not describe this rank's block")
    rmax = max(int(r) for r in rows_per_rank)
    cols = int(local.shape[1])
    if rmax == 0:
        return local.new_empty((0, cols))
    padded = local
    if local.shape[0] != rmax:
        padded = local.new_zeros((rmax, cols))
        padded[: local.shape[0]] = local
    out = local.new_empty((world * rmax, cols))
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    if all(int(r) == rmax for r in rows_per_rank):
        return out
    return torch.cat([out[r * rmax: r * rmax + int(rows_per_rank[r])] for r in range(world)], dim=0)


def featurise_sharded(obj, band, t, flux, err, n_obj: int, modules: Sequence[str], z=None, cost: str = "linear",
                      group=None, device=None):
    """Every rank calls this with the same flat host columns; returns ``{module: [n_obj, cols] CUDA tensor}``
    holding the full table on every rank.  One store per rank, passes on the local shard, one all-gather per
    module table."""
    import torch
    import torch.distributed as dist
    from . import _native
    from .store import LightCurveStore
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = torch.cuda.current_device() if device is None else device
    ranges = balanced_ranges(object_costs(np.asarray(obj), n_obj, cost), world)
    lo, hi = ranges[rank]
    keep, local_obj = shard_rows(np.asarray(obj), lo, hi)
    out = {}
    with LightCurveStore(local_obj, np.asarray(band)[keep], np.asarray(t)[keep], np.asarray(flux)[keep],
                         np.asarray(err)[keep], hi - lo, device=dev) as store:
        for m in modules:
            rows = torch.empty((hi - lo, _native.COLS[m]), dtype=torch.float64, device=f"cuda:{dev}")
            zz = None if (z is None or m != "physics") else np.asarray(z)[lo:hi]
            store.features(m, z=zz, out=rows)
            torch.cuda.synchronize(dev)
            out[m] = rows if world == 1 else allgather_rows(rows, [b - a for a, b in ranges], group)
    return out
