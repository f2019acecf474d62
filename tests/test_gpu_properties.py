"""Size-independent properties of the GPU path at a scale the CPU oracle cannot check (tens of thousands of
objects, millions of observations): determinism, independence of the batch composition, and invariance under
re-interleaving of the input rows (each object's own row order -- which the reference's features depend on -- kept)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MODULES = ("statistical", "colors", "shape", "tde", "physics", "bazin", "powerlaw", "gp2d")
N_OBJ = 30000


@pytest.fixture(scope="module")
def table():
    from mallorn_astrophysics_b200.synth import make_table
    return make_table(N_OBJ, config_index=4)


def _run(tb, obj=None, rows=None, n_obj=None, z=None):
    from mallorn_astrophysics_b200.store import LightCurveStore
    sel = slice(None) if rows is None else rows
    o = tb.obj[sel] if obj is None else obj
    with LightCurveStore(o, tb.band[sel], tb.t[sel], tb.flux[sel], tb.err[sel], tb.n_obj if n_obj is None else n_obj) as s:
        return {m: s.features(m, z=(tb.z if z is None else z) if m == "physics" else None) for m in MODULES}, s.segments()


def _same(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def test_runs_are_bit_identical(table):
    a, sa = _run(table)
    b, sb = _run(table)
    assert np.array_equal(sa, sb)
    for m in MODULES:
        assert _same(a[m], b[m]), m


def test_rows_do_not_depend_on_the_rest_of_the_batch(table):
    """Featurising the second half of the objects alone gives the rows it got inside the full batch."""
    full, _ = _run(table)
    half = table.n_obj // 2
    keep = np.flatnonzero(table.obj >= half)
    part, _ = _run(table, obj=(table.obj[keep] - half).astype(np.int32), rows=keep, n_obj=table.n_obj - half, z=table.z[half:])
    for m in MODULES:
        assert _same(part[m], full[m][half:]), m


def test_reinterleaving_the_input_rows_changes_nothing(table):
    """The flat table is concatenated split files upstream: how objects interleave is arbitrary.  A permutation of
    the rows that keeps every object's own rows in their original relative order must not change a bit."""
    full, seg = _run(table)
    rng = np.random.default_rng(5)
    order = np.argsort(table.obj, kind="stable")                  # rows grouped per object, original order kept
    obj_of_slot = table.obj[rng.permutation(table.n_rows)]        # which object each output slot belongs to
    out_order = np.argsort(obj_of_slot, kind="stable")            # output slots grouped by object, ascending slot
    perm = np.empty(table.n_rows, dtype=np.int64)
    perm[out_order] = order                                       # i-th slot of object o <- its i-th original row
    assert np.array_equal(np.sort(perm), np.arange(table.n_rows))
    assert np.array_equal(table.obj[perm], obj_of_slot)
    shuffled, seg2 = _run(table, rows=perm)
    assert np.array_equal(seg, seg2)
    for m in MODULES:
        assert _same(shuffled[m], full[m]), m


def test_the_overlapped_fit_call_equals_the_two_passes(table):
    """mlrn_fits (Bazin on the caller's stream, decline models on an internal side stream) returns the rows and status
    tables of mlrn_bazin + mlrn_powerlaw bit for bit, with host buffers and with device buffers."""
    import torch
    from mallorn_astrophysics_b200.store import LightCurveStore
    sub = table.subset(6000)
    with LightCurveStore.from_table(sub) as s:
        sb, sp = np.zeros((s.n_obj, 6, 2), np.int32), np.zeros((s.n_obj, 27, 2), np.int32)
        want_b, want_p = s.features("bazin", status=sb), s.features("powerlaw", status=sp)
        hb, hp = np.empty_like(want_b), np.empty_like(want_p)
        hsb, hsp = np.zeros_like(sb), np.zeros_like(sp)
        s.fit_passes(hb, hp, status_bazin=hsb, status_powerlaw=hsp)
        db = torch.empty(want_b.shape, dtype=torch.float64, device="cuda")
        dp = torch.empty(want_p.shape, dtype=torch.float64, device="cuda")
        dsb = torch.zeros(sb.shape, dtype=torch.int32, device="cuda")
        s.fit_passes(db, dp, status_bazin=dsb)
        torch.cuda.synchronize()
    assert _same(hb, want_b) and _same(hp, want_p)
    assert np.array_equal(hsb, sb) and np.array_equal(hsp, sp)
    assert _same(db.cpu().numpy(), want_b) and _same(dp.cpu().numpy(), want_p)
    assert np.array_equal(dsb.cpu().numpy(), sb)
