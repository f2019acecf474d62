"""GPU parity tests (run on the B200 box: ``pytest -m gpu``).  Everything goes through the C-ABI
(ctypes -> libmallorn_b200.so); the CPU oracle and the committed reference goldens are the checkers."""
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TIE_OBJECT = 15       # ref_edge object with duplicated times (sort order of ties is unspecified upstream)


def _store(g):
    from mallorn_astrophysics_b200.store import LightCurveStore
    return LightCurveStore(g["obj"], g["band"], g["t"], g["flux"], g["err"], int(g["n_obj"]))


def _compare(got, want, rtol, atol=0.0, what=""):
    assert got.shape == want.shape, what
    nan_mismatch = np.isnan(got) != np.isnan(want)
    assert not nan_mismatch.any(), f"{what}: NaN pattern differs at {np.argwhere(nan_mismatch)[:5]}"
    ok = ~np.isnan(want)
    np.testing.assert_allclose(got[ok], want[ok], rtol=rtol, atol=atol, err_msg=what)


@pytest.mark.parametrize("fixture", ["ref_sparse", "ref_dense", "ref_edge", "ref_c1_256"])
def test_pack_is_bit_exact(golden, fixture):
    """K1: offsets, counts, ordering, file ranks and peak indices are integer/bit-exact."""
    from packref import pack_reference
    from oracle.extract import run_module
    g = golden[fixture]
    n_obj = int(g["n_obj"])
    want = pack_reference(g["obj"], g["band"], g["t"], g["flux"], g["err"], n_obj)
    with _store(g) as s:
        got = s.export()
        seg = s.segments()
        assert s.n_rows == len(want["t"]) and s.n_obj == n_obj
    assert np.array_equal(got["seg_off"], want["seg_off"])
    for k in ("t", "flux", "err"):
        assert np.array_equal(got[k].view(np.uint64), want[k].view(np.uint64)), k
    assert np.array_equal(got["rank"], want["rank"])
    assert np.array_equal(got["trank"], want["trank"])
    present, meta = run_module("segments", g["obj"], g["band"], g["t"], g["flux"], g["err"], n_obj)
    assert np.array_equal(seg[present].astype(np.int64), meta)
    absent = np.setdiff1d(np.arange(n_obj), present)
    assert (seg[absent][:, 0:21:3] == 0).all()


@pytest.mark.parametrize("fixture", ["ref_sparse", "ref_dense", "ref_edge", "ref_c1_256"])
def test_statistical_matches_reference_golden(golden, fixture):
    """K2 vs the unmodified reference's own output; moments to rtol 1e-5 (north star)."""
    g = golden[fixture]
    with _store(g) as s:
        got = s.features("statistical")
    present = g["statistical_present"]
    got, want = got[present], g["statistical_rows"]
    if fixture == "ref_edge":
        keep = present != TIE_OBJECT
        got, want = got[keep], want[keep]
    _compare(got, want, rtol=1e-5, atol=1e-9, what=f"statistical/{fixture}")
    # integer-valued columns are exact
    from mallorn_astrophysics_b200 import schema
    idx = [schema.STAT_COLUMNS.index(c) for c in schema.STAT_INT_COLUMNS]
    assert np.array_equal(got[:, idx], want[:, idx])


@pytest.mark.parametrize("module", ["colors", "shape", "tde", "physics"])
@pytest.mark.parametrize("fixture", ["ref_sparse", "ref_dense", "ref_edge", "ref_c1_256"])
def test_scalar_passes_match_reference_golden(golden, fixture, module):
    """K3 vs the unmodified reference's own output (rtol 1e-5; atol covers slopes that are 0 up to rounding)."""
    g = golden[fixture]
    with _store(g) as s:
        got = s.features(module, z=g["z"])
    present = g[f"{module}_present"]
    got, want = got[present], g[f"{module}_rows"]
    if fixture == "ref_edge":
        keep = present != TIE_OBJECT
        got, want = got[keep], want[keep]
    _compare(got, want, rtol=1e-5, atol=1e-9, what=f"{module}/{fixture}")


@pytest.mark.parametrize("module", ["colors", "shape", "tde", "physics"])
def test_scalar_passes_match_oracle_at_scale(module):
    from mallorn_astrophysics_b200.store import LightCurveStore
    from mallorn_astrophysics_b200.synth import make_table
    from oracle.extract import run_table
    table = make_table(2048, config_index=2, count_scale=1.5)
    with LightCurveStore.from_table(table) as s:
        got = s.features(module, z=table.z)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        present, want = run_table(module, table)
    _compare(got[present], want, rtol=1e-5, atol=1e-9, what=f"{module}/2048")


def test_statistical_matches_oracle_at_scale():
    """4096 seeded objects against the oracle (seconds on the CPU side)."""
    from mallorn_astrophysics_b200.store import LightCurveStore
    from mallorn_astrophysics_b200.synth import make_table
    from oracle.extract import run_table
    table = make_table(4096, config_index=1)
    with LightCurveStore.from_table(table) as s:
        got = s.features("statistical")
        seg = s.segments()
    present, want = run_table("statistical", table)
    _compare(got[present], want, rtol=1e-5, atol=1e-9, what="statistical/4096")
    _, meta = run_table("segments", table)
    assert np.array_equal(seg[present].astype(np.int64), meta)


def test_empty_and_degenerate_inputs():
    from mallorn_astrophysics_b200.store import LightCurveStore
    e = np.empty(0)
    with LightCurveStore(np.empty(0, np.int32), np.empty(0, np.uint8), e, e, e, 0) as s:
        assert s.n_rows == 0 and s.features("statistical").shape == (0, 123)
    # three objects, only the middle one has rows; row of an unrequested object is skipped
    obj = np.array([1, 1, -1, 1], np.int32)
    band = np.array([2, 2, 2, 9], np.uint8)
    t = np.array([60000.5, 60001.5, 60002.5, 60003.5])
    f = np.array([1.0, 3.0, 100.0, 2.0])
    err = np.array([0.1, 0.1, 0.1, 0.1])
    with LightCurveStore(obj, band, t, f, err, 3) as s:
        assert s.n_rows == 3 and s.max_obs == 3
        seg = s.segments()
        st = s.features("statistical")
    assert seg[1, 6] == 2 and seg[1, 18] == 1 and seg[0, 0:21:3].sum() == 0
    assert st[1, 2 * 17 + 0] == 2 and st[1, 6 * 17 + 0] == 3 and st[1, 6 * 17 + 4] == 3.0
    assert st[1, 122] == 2 and st[0, 122] == -1


def test_requested_objects_without_any_row():
    """An empty light-curve frame with requested ids (n_rows == 0, n_obj > 0): every segment count is 0, peak positions
    are -1, no object is present, and the passes return rows for it (NaN / n_obs 0) instead of reading garbage."""
    from mallorn_astrophysics_b200.store import LightCurveStore
    e = np.empty(0)
    with LightCurveStore(np.empty(0, np.int32), np.empty(0, np.uint8), e, e, e, 5) as s:
        assert s.n_rows == 0 and s.n_obj == 5
        seg = s.segments()
        assert s.present().size == 0
        st = s.features("statistical")
        bz = s.features("bazin")
    assert (seg[:, 0:21:3] == 0).all() and (seg[:, 1:21:3] == -1).all() and (seg[:, 2:21:3] == -1).all() and (seg[:, 21] == -1).all()
    assert (st[:, 0:7 * 17:17] == 0).all() and np.isnan(st[:, 1]).all() and np.isnan(bz).all()


def test_output_buffers_must_match_exactly():
    """Outputs are never converted: a column slice of a wider table, a wrong dtype, shape or side raises instead of
    filling a temporary."""
    import torch
    from mallorn_astrophysics_b200.store import LightCurveStore
    from mallorn_astrophysics_b200.synth import make_table
    table = make_table(8, seed=5)
    with LightCurveStore.from_table(table) as s:
        wide = torch.zeros((8, 200), dtype=torch.float64, device="cuda")
        with pytest.raises(ValueError):
            s.features("statistical", out=wide[:, :123])
        with pytest.raises(ValueError):
            s.features("statistical", out=torch.zeros((8, 123), dtype=torch.float32, device="cuda"))
        with pytest.raises(ValueError):
            s.features("statistical", out=np.zeros((7, 123)))
        with pytest.raises(ValueError):
            s.features("bazin", out=torch.zeros((8, 52), dtype=torch.float64, device="cuda"), status=np.zeros((8, 12), np.int32))
        with pytest.raises(ValueError):
            s.features("bazin", status=np.zeros((8, 11), np.int32))
        assert (wide == 0).all()
        good = torch.zeros((8, 123), dtype=torch.float64, device="cuda")
        s.features("statistical", out=good)
        torch.cuda.synchronize()
        assert np.array_equal(good.cpu().numpy(), s.features("statistical"), equal_nan=True)


def test_object_above_cap_is_rejected():
    from mallorn_astrophysics_b200 import _native
    from mallorn_astrophysics_b200.store import LightCurveStore
    n = _native.MAX_OBS + 1
    with pytest.raises(_native.MallornNativeError):
        LightCurveStore(np.zeros(n, np.int32), np.zeros(n, np.uint8), np.arange(n, dtype=float),
                        np.ones(n), np.ones(n), 1)


def test_one_very_long_object_does_not_fail_the_batch():
    """The reference takes any length (statistical.py:163-190).  One 3,000-observation object among ordinary ones: K1 stays
    bit-exact, the scalar passes match the oracle on every object, the fit and GP passes return rows for it.  (Most of
    its errors are zero, which the GP drops as invalid, so the 2-D GP sees ~400 points instead of an N^3 = 2.7e10 job.)"""
    from packref import pack_reference
    from mallorn_astrophysics_b200.store import LightCurveStore
    from mallorn_astrophysics_b200.synth import LightCurveTable, make_table
    from oracle.extract import run_table
    base = make_table(40, seed=99, count_scale=2.0)              # ~300 observations each
    obj = base.obj.copy()
    obj[obj < 10] = 0                                             # the first ten objects become ONE object of ~3,000 observations
    obj[obj >= 10] -= 9
    err = base.err.copy()
    rng = np.random.default_rng(5)
    long_rows = np.flatnonzero(obj == 0)
    err[rng.choice(long_rows, size=len(long_rows) - 400, replace=False)] = 0.0
    table = LightCurveTable(obj.astype(np.int32), base.band, base.t, base.flux, err, 31, base.z[9:].copy(), base.ebv[9:].copy(),
                            base.label[9:].copy(), "vlong")
    n_long = int(np.bincount(table.obj)[0])
    assert n_long > 2500
    want_pack = pack_reference(table.obj, table.band, table.t, table.flux, table.err, table.n_obj)
    with LightCurveStore.from_table(table) as s:
        assert s.max_obs == n_long
        got_pack, seg = s.export(), s.segments()
        got = {m: s.features(m, z=table.z if m in ("physics", "research") else None)
               for m in ("statistical", "colors", "shape", "tde", "physics", "research", "bazin", "powerlaw", "gp2d", "gp1d")}
    assert np.array_equal(got_pack["seg_off"], want_pack["seg_off"])
    for k in ("t", "flux", "err"):
        assert np.array_equal(got_pack[k].view(np.uint64), want_pack[k].view(np.uint64)), k
    assert np.array_equal(got_pack["rank"], want_pack["rank"]) and np.array_equal(got_pack["trank"], want_pack["trank"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        present, meta = run_table("segments", table)
        assert np.array_equal(seg[present].astype(np.int64), meta)
        for m in ("statistical", "colors", "shape", "tde", "physics", "research"):
            present, want = run_table(m, table)
            _compare(got[m][present], want, rtol=1e-5, atol=1e-9, what=f"{m}/very long")
        _, want_gp = run_table("gp2d", table.subset(1))
    _compare(got["gp2d"][:1], want_gp, rtol=1e-3, atol=1e-6, what="gp2d/very long")
    assert np.isfinite(got["bazin"][0, 8 * 2]) or np.isnan(got["bazin"][0, 8 * 2])
    assert np.isfinite(got["powerlaw"][0]).any() and np.isfinite(got["gp1d"][0]).any()


def test_long_objects_up_to_the_cap():
    """A few objects of 512-1024 observations (far beyond survey cadence; they take the global-scratch launches): every pass
    must still run, and the scalar passes must still match the oracle."""
    from mallorn_astrophysics_b200.store import LightCurveStore
    from mallorn_astrophysics_b200.synth import make_table
    from oracle.extract import run_table
    from mallorn_astrophysics_b200.synth import LightCurveTable
    base = make_table(12, seed=321, count_scale=2.5)           # ~375 observations each (the generator clips at 400) ...
    table = LightCurveTable((base.obj // 2).astype(np.int32), base.band, base.t, base.flux, base.err, 6,   # ... merged in pairs
                            base.z[::2].copy(), base.ebv[::2].copy(), base.label[::2].copy(), "long")
    assert 512 < np.bincount(table.obj).max() <= 1024
    with LightCurveStore.from_table(table) as s:
        assert s.max_obs > 512
        got = {m: s.features(m, z=table.z if m == "physics" else None)
               for m in ("statistical", "colors", "shape", "tde", "physics", "bazin", "powerlaw", "gp2d")}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for m in ("statistical", "colors", "shape", "tde", "physics"):
            present, want = run_table(m, table)
            _compare(got[m][present], want, rtol=1e-5, atol=1e-9, what=f"{m}/long")
    assert np.isfinite(got["gp2d"][:, 0]).all() and np.isfinite(got["bazin"][:, 8 * 2]).any()


def test_statistical_fast_path_corner_cases():
    """stat_warp.cuh: 24-bit key ties repaired on the exact values (a huge outlier squeezes every other flux of the object into
    a handful of keys, in reversed order), 255 observations (last object of the fast path) next to 256 (first of the general
    kernel), constant fluxes, one- and two-point bands, an unknown-filter segment: all against the oracle."""
    from mallorn_astrophysics_b200.store import LightCurveStore
    from mallorn_astrophysics_b200.synth import LightCurveTable
    from oracle.extract import run_table
    rng = np.random.default_rng(11)
    obj, band, t, f, e = [], [], [], [], []

    def add(o, n, flux, bands=None, times=None):
        obj.append(np.full(n, o)); f.append(np.asarray(flux, dtype=float))
        band.append(rng.integers(0, 6, n) if bands is None else np.asarray(bands))
        t.append(59000.0 + np.sort(rng.uniform(0.0, 900.0, n)) if times is None else np.asarray(times, dtype=float))
        e.append(rng.uniform(0.5, 2.0, n))

    add(0, 200, np.concatenate([1000.0 - 1e-6 * np.arange(199), [1e9]]))             # ties, reversed order, one outlier
    add(1, 255, rng.normal(50.0, 5.0, 255))
    add(2, 256, rng.normal(50.0, 5.0, 256))
    add(3, 120, np.full(120, 3.25))                                                    # constant
    add(4, 1, [7.0], bands=[2])
    add(5, 12, rng.normal(0.0, 1.0, 12), bands=[0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5])   # two points per band
    add(6, 90, rng.normal(-3.0, 1.0, 90), bands=rng.integers(0, 7, 90))               # with unknown-filter rows
    add(7, 64, np.repeat(rng.normal(10.0, 2.0, 8), 8), bands=np.tile(np.arange(4), 16))   # exact duplicates
    add(8, 150, rng.normal(1e-3, 1e-6, 150) + np.where(np.arange(150) == 77, 5e3, 0.0))  # outlier, ordinary order
    n_obj = 9
    o = np.concatenate(obj).astype(np.int32)
    perm = rng.permutation(len(o))                                                       # file order is not object order
    cols = [np.concatenate(x)[perm] for x in (obj, band, t, f, e)]
    table = LightCurveTable(cols[0].astype(np.int32), cols[1].astype(np.uint8), cols[2], cols[3], cols[4], n_obj,
                            np.zeros(n_obj), np.zeros(n_obj), np.zeros(n_obj, np.int64), "corners")
    with LightCurveStore.from_table(table) as s:
        got = s.features("statistical")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        present, want = run_table("statistical", table)
    _compare(got[present], want, rtol=1e-5, atol=1e-9, what="statistical/corner cases")


def test_non_finite_fluxes_stay_with_their_object():
    """An object with NaN / inf fluxes takes the general moment kernel (the fast path declines it) and every pass still
    returns: the rows of all OTHER objects are bit for bit those of the same batch with that object's fluxes finite."""
    from mallorn_astrophysics_b200.store import LightCurveStore
    from mallorn_astrophysics_b200.synth import make_table
    tb = make_table(64, seed=5)
    bad = tb.flux.copy()
    rows = np.flatnonzero(tb.obj == 17)
    bad[rows[3]] = np.nan
    bad[rows[7]] = np.inf
    bad[rows[11]] = -np.inf
    out = {}
    for name, flux in (("clean", tb.flux), ("bad", bad)):
        with LightCurveStore(tb.obj, tb.band, tb.t, flux, tb.err, tb.n_obj) as s:
            out[name] = {m: s.features(m, z=tb.z if m == "physics" else None)
                         for m in ("statistical", "colors", "shape", "tde", "physics", "bazin", "powerlaw", "gp2d")}
    others = np.arange(tb.n_obj) != 17
    for m, a in out["clean"].items():
        b = out["bad"][m]
        assert np.array_equal(a[others].view(np.uint64), b[others].view(np.uint64)), m
    assert out["bad"]["statistical"][17, 0] >= 0          # n_obs of the first band is still a count
