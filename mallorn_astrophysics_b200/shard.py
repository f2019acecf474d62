"""Object sharding across the GPUs of one box and the single all-gather of the feature table.

Objects are independent: each rank packs and featurises a contiguous range of the
dense object index, chosen so the ranges carry equal *cost* (observation counts by default; the GP pass
costs ~N^3 per object, so ``cost='cubic'`` balances that instead).  Nothing is exchanged during compute;
at the end ONE ``all_gather_into_tensor`` of the padded ``[rows_max, n_cols]`` float64 table -- all requested module
tables side by side -- (NCCL over NVLink on the B200 box, gloo in the CPU tests) gives every rank the full table in object
order.  The gathered table is float64 x every module column (434 for the eight passes of the v34a/v92d superset): the
reference's frames are float64 and the 222 columns are selected downstream by name, so neither the precision nor the
width is cut on the wire (27.8 GB per rank at 8 x 10^6 objects, 65 ms over NVLink 5 in round 1).
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


def equal_ranges(n_obj: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous ranges of (nearly) equal object counts: the cost-balanced split when per-object costs are i.i.d.
    (a synthetic stream generated per shard, where no rank ever sees the whole set)."""
    cuts = [round(n_obj * r / world) for r in range(world + 1)]
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def batch_ranges(ranges: Sequence[Tuple[int, int]], per_batch: int) -> List[List[Tuple[int, int]]]:
    """Split every rank's contiguous range into the SAME number of contiguous batches of at most ``per_batch`` objects
    (the tallest range decides the number): ``out[rank][batch] = (lo, hi)``.  A set too large to featurise in one go is
    worked through batch by batch, batch ``j`` on every rank at the same time, with the all-gather of batch ``j``'s rows
    after it; every object is in exactly one batch."""
    if per_batch <= 0:
        raise ValueError("per_batch must be positive")
    tallest = max((b - a for a, b in ranges), default=0)
    n_batches = max(1, -(-tallest // per_batch))
    return [[(a + x, a + y) for x, y in equal_ranges(b - a, n_batches)] for a, b in ranges]


def table_width(modules: Sequence[str]) -> int:
    from . import _native
    return sum(_native.COLS[m] for m in modules)


def featurise_shard(cols: dict, n_local: int, modules: Sequence[str], table=None, device=None, stream: int = 0,
                    status: dict = None, on_pass=None):
    """One rank's share of the work: pack the shard's flat columns (``obj`` already re-based to ``0..n_local``; host arrays or
    CUDA tensors) into one store and run every pass of ``modules`` on it, writing the rows side by side into ONE
    ``[n_local, table_width(modules)]`` float64 CUDA table (module order = column order).  Nothing is synchronised:
    everything is ordered on ``stream``; ``on_pass(name)`` is called after each launch sequence and, as ``on_pass('table')``,
    after the copy of a pass's rows into the wide table (the bench records events there).  ``cols['z']`` (per-object redshift) is used by the physics pass when present."""
    import torch
    from . import _native
    from .store import LightCurveStore
    dev = torch.cuda.current_device() if device is None else int(device)
    width = table_width(modules)
    if table is None:
        table = torch.empty((n_local, width), dtype=torch.float64, device=f"cuda:{dev}")
    elif tuple(table.shape) != (n_local, width) or table.dtype != torch.float64 or not table.is_contiguous():
        raise ValueError(f"table must be a contiguous float64 CUDA tensor of shape {(n_local, width)}")
    store = LightCurveStore(cols["obj"], cols["band"], cols["t"], cols["flux"], cols["err"], n_local, device=dev, stream=stream)
    if on_pass:
        on_pass("pack")
    try:
        offs, c0 = {}, 0
        for m in modules:
            offs[m] = c0
            c0 += _native.COLS[m]
        pair = "bazin" in offs and "powerlaw" in offs          # both fit passes: one call, their straggler stages overlap (mlrn_fits)
        for m in modules:
            w = _native.COLS[m]
            if pair and m in ("bazin", "powerlaw"):
                if m != next(x for x in modules if x in ("bazin", "powerlaw")):
                    continue                                   # done with the first of the two
                rb = torch.empty((n_local, _native.COLS["bazin"]), dtype=torch.float64, device=f"cuda:{dev}")
                rp = torch.empty((n_local, _native.COLS["powerlaw"]), dtype=torch.float64, device=f"cuda:{dev}")
                store.fit_passes(rb, rp, status_bazin=(status or {}).get("bazin"), status_powerlaw=(status or {}).get("powerlaw"))
                if on_pass:
                    on_pass("bazin+powerlaw")
                table[:, offs["bazin"]:offs["bazin"] + rb.shape[1]].copy_(rb)
                table[:, offs["powerlaw"]:offs["powerlaw"] + rp.shape[1]].copy_(rp)
                if on_pass:
                    on_pass("table")
                continue
            rows = torch.empty((n_local, w), dtype=torch.float64, device=f"cuda:{dev}")     # dense rows of this pass; freed (stream-ordered) after the copy
            z = cols.get("z") if m in ("physics", "research") else None
            if z is not None and not (hasattr(z, "is_cuda") and z.is_cuda):
                z = torch.from_numpy(np.array(z, dtype=np.float64)).to(f"cuda:{dev}")
            store.features(m, z=z, out=rows, status=(status or {}).get(m))
            if on_pass:
                on_pass(m)                                     # the pass's own launches ...
            table[:, offs[m]:offs[m] + w].copy_(rows)
            if on_pass:
                on_pass("table")                               # ... and the copy of its rows into the wide table, marked separately
    finally:
        # destroying the store frees its arrays with cudaFree, which waits for the device: the passes above are complete
        store.close()
    return table


def allgather_table(local, rows_per_rank: Sequence[int], out=None, group=None):
    """The ONE collective of the path: all-gather of the shards' tables (padded to the tallest shard) so that every rank
    holds the full ``[sum(rows), width]`` table in object order.  ``out``: optional preallocated ``[world * max(rows), width]``
    receive buffer (returned as is when all shards are equally tall)."""
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
    if out is None:
        out = local.new_empty((world * rmax, cols))
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    if all(int(r) == rmax for r in rows_per_rank):
        return out
    return torch.cat([out[r * rmax: r * rmax + int(rows_per_rank[r])] for r in range(world)], dim=0)


def featurise_sharded(source, n_obj: int, modules: Sequence[str], cost: str = "linear", group=None, device=None,
                      gather: bool = True, on_pass=None):
    """The multi-GPU entry: every rank calls this; returns ``(table, ranges)`` with ``table`` the full
    ``[n_obj, table_width(modules)]`` float64 CUDA table on every rank (``gather=False``: only this rank's rows).

    ``source`` is either
      * a dict of flat columns over the WHOLE set (``obj, band, t, flux, err`` and optionally ``z``; host arrays), split
        into contiguous object ranges of equal ``cost`` (observation counts, or their cube for the GP) -- every rank then
        masks out its own rows; or
      * a callable ``source(lo, hi) -> dict`` producing the flat columns of objects ``[lo, hi)`` with ``obj`` re-based to
        ``0..hi-lo`` (a per-shard reader or generator: at 10^7 objects no rank ever holds the whole set); ranges are then
        equal object counts.
    One store per rank, no collective during compute, ONE all-gather of the concatenated table at the end."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = torch.cuda.current_device() if device is None else device
    if callable(source):
        ranges = equal_ranges(n_obj, world)
        lo, hi = ranges[rank]
        cols = source(lo, hi)
    else:
        obj = np.asarray(source["obj"])
        ranges = balanced_ranges(object_costs(obj, n_obj, cost), world)
        lo, hi = ranges[rank]
        keep, local_obj = shard_rows(obj, lo, hi)
        cols = {"obj": local_obj}
        for k in ("band", "t", "flux", "err"):
            cols[k] = np.asarray(source[k])[keep]
        if source.get("z") is not None:
            cols["z"] = np.asarray(source["z"])[lo:hi]
    table = featurise_shard(cols, hi - lo, modules, device=dev, on_pass=on_pass)
    if world == 1 or not gather:
        return table, ranges
    full = allgather_table(table, [b - a for a, b in ranges], group=group)
    if on_pass:
        on_pass("allgather")
    return full, ranges


allgather_rows = allgather_table      # round-1 name
