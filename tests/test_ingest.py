"""Host ingest (store.frame_columns): the DataFrame -> flat-column factorisation that replaces the reference's
``groupby('object_id')`` dicts, checked on the CPU against a plain per-row restatement."""
import numpy as np
import pandas as pd
import pytest

from mallorn_astrophysics_b200.schema import BAND_CODE, UNKNOWN_BAND
from mallorn_astrophysics_b200.store import frame_columns


def _frame(n_rows=5000, n_ids=37, seed=3):
    rng = np.random.default_rng(seed)
    ids = np.array([f"Name_{k:03d}_x" for k in range(n_ids)], dtype=object)
    filt = np.array(["u", "g", "r", "i", "z", "y", "w", None], dtype=object)
    lc = pd.DataFrame({"object_id": ids[rng.integers(0, n_ids, n_rows)],
                       "Time (MJD)": 60000 + rng.uniform(0, 1000, n_rows),
                       "Flux": rng.normal(0, 5, n_rows), "Flux_err": rng.uniform(0.1, 2, n_rows),
                       "Filter": filt[rng.integers(0, len(filt), n_rows)]})
    return lc, ids


def _plain(lc, object_ids):
    """row by row, as the docstring states it"""
    if object_ids is None:
        object_ids = list(lc["object_id"].dropna().unique())
    ids = list(dict.fromkeys(object_ids))
    pos = {s: k for k, s in enumerate(ids)}
    obj = np.array([pos.get(s, -1) if isinstance(s, str) else -1 for s in lc["object_id"]], dtype=np.int32)
    band = np.array([BAND_CODE.get(b, UNKNOWN_BAND) if isinstance(b, str) else UNKNOWN_BAND for b in lc["Filter"]], dtype=np.uint8)
    return obj, band, ids


@pytest.mark.parametrize("mode", ["default", "subset_reversed", "with_duplicates_and_strangers"])
def test_frame_columns_matches_the_row_by_row_rule(mode):
    lc, ids = _frame()
    if mode == "default":
        req = None
    elif mode == "subset_reversed":
        req = list(ids[::-2])
    else:
        req = [ids[5], "Not_in_frame", ids[2], ids[5], ids[30]]
    before = lc.copy()
    obj, band, t, f, e, out_ids = frame_columns(lc, req)
    want_obj, want_band, want_ids = _plain(lc, req)
    assert list(out_ids) == want_ids
    assert obj.dtype == np.int32 and band.dtype == np.uint8
    assert np.array_equal(obj, want_obj) and np.array_equal(band, want_band)
    assert np.array_equal(t, lc["Time (MJD)"].to_numpy()) and np.array_equal(f, lc["Flux"].to_numpy())
    assert lc.equals(before)


def test_frame_columns_missing_ids_and_empty_frames():
    lc, ids = _frame(200, 5)
    lc.loc[[3, 17], "object_id"] = None
    obj, band, *_ , out_ids = frame_columns(lc, None)
    assert obj[3] == -1 and obj[17] == -1 and (obj[[0, 1, 2]] >= 0).all()
    assert len(out_ids) == 5
    empty = lc.iloc[0:0]
    obj, band, t, f, e, out_ids = frame_columns(empty, ["a", "b"])
    assert len(obj) == 0 and len(band) == 0 and list(out_ids) == ["a", "b"]
    obj, band, t, f, e, out_ids = frame_columns(empty, None)
    assert len(obj) == 0 and len(out_ids) == 0


# ---- native factorisation of Arrow-backed string columns (store.frame_columns' fast path) --------------------------------
def _as_arrow_series(values, typ, n_chunks=1, slice_from=0):
    import pandas as pd
    import pyarrow as pa
    vals = list(values)
    cuts = np.linspace(0, len(vals), n_chunks + 1).astype(int)
    chunks = [pa.array(["pad"] * slice_from + vals[a:b], type=typ).slice(slice_from) for a, b in zip(cuts[:-1], cuts[1:])]
    return pd.Series(pd.arrays.ArrowExtensionArray(pa.chunked_array(chunks, type=typ)))


@pytest.mark.parametrize("typ_name", ["large_string", "string"])
@pytest.mark.parametrize("n_chunks,slice_from", [(1, 0), (5, 0), (3, 7)])
def test_native_factorize_equals_pandas(typ_name, n_chunks, slice_from):
    import pandas as pd
    import pyarrow as pa
    from mallorn_astrophysics_b200 import ingest
    rng = np.random.default_rng(5)
    pool = np.array([f"obj_{i:08d}" for i in range(700)] + ["x", "", "a-much-longer-identifier-than-the-others-0123456789", "naïve✓"], dtype=object)
    vals = pool[rng.integers(0, len(pool), 60_000)].tolist()
    for k in rng.integers(0, len(vals), 300):
        vals[k] = None                                          # missing ids -> code -1
    s = _as_arrow_series(vals, getattr(pa, typ_name)(), n_chunks, slice_from)
    codes, uniq = ingest.factorize_strings(s, threads=4)
    want_codes, want_uniq = pd.factorize(pd.Series(vals, dtype=object), sort=False)
    assert codes.dtype == np.int32 and np.array_equal(codes, want_codes)
    assert list(uniq) == list(want_uniq)
    fil = np.array(["u", "g", "r", "i", "z", "y", "Y", "gr", "", None], dtype=object)[rng.integers(0, 10, 60_000)].tolist()
    band = ingest.band_codes(_as_arrow_series(fil, getattr(pa, typ_name)(), n_chunks, slice_from), threads=3)
    lut = {"u": 0, "g": 1, "r": 2, "i": 3, "z": 4, "y": 5}
    assert band.dtype == np.uint8 and band.tolist() == [lut.get(x, 6) for x in fil]


def test_native_factorize_edge_cases_and_fallback():
    import pandas as pd
    import pyarrow as pa
    from mallorn_astrophysics_b200 import ingest
    empty = _as_arrow_series([], pa.large_string())
    codes, uniq = ingest.factorize_strings(empty)
    assert len(codes) == 0 and len(uniq) == 0 and len(ingest.band_codes(empty)) == 0
    codes, uniq = ingest.factorize_strings(_as_arrow_series([None, None], pa.large_string()))
    assert codes.tolist() == [-1, -1] and len(uniq) == 0
    grouped = _as_arrow_series(["b"] * 5 + ["a"] * 3 + ["b"] * 2, pa.string())       # runs of one id (the reference's files)
    codes, uniq = ingest.factorize_strings(grouped, threads=1)
    assert codes.tolist() == [0] * 5 + [1] * 3 + [0] * 2 and list(uniq) == ["b", "a"]
    # not Arrow-backed: the caller falls back to pandas
    assert ingest.factorize_strings(pd.Series(["a", "b"], dtype=object)) is None
    assert ingest.band_codes(pd.Series(["u", "g"], dtype=object)) is None
    assert ingest.factorize_strings(pd.Series([1, 2])) is None


@pytest.mark.parametrize("backing", ["arrow", "object"])
def test_frame_columns_same_result_for_arrow_and_object_strings(backing):
    """frame_columns (every drop-in extractor's host half) gives the same columns whichever way the frame stores its strings."""
    import pandas as pd
    from mallorn_astrophysics_b200.store import frame_columns
    from mallorn_astrophysics_b200.synth import make_table
    tb = make_table(300, seed=11)
    lc = tb.to_frame()
    lc.loc[lc.index[::97], "Filter"] = None
    lc.loc[lc.index[5::211], "object_id"] = None
    if backing == "object":
        lc = lc.astype({"object_id": object, "Filter": object})
    else:
        lc = lc.astype({"object_id": "string[pyarrow]", "Filter": "string[pyarrow]"})
    ids = list(tb.object_ids[::-1][:250]) + ["not-in-the-frame"]
    ref = lc.astype({"object_id": object, "Filter": object})
    code, uniq = pd.factorize(ref["object_id"], sort=False)
    pos = {v: k for k, v in enumerate(ids)}
    want_obj = np.array([pos.get(uniq[c], -1) if c >= 0 else -1 for c in code], dtype=np.int32)
    want_band = np.array([{"u": 0, "g": 1, "r": 2, "i": 3, "z": 4, "y": 5}.get(x, 6) for x in ref["Filter"]], dtype=np.uint8)
    obj, band, t, flux, err, out_ids = frame_columns(lc, ids)
    assert obj.dtype == np.int32 and np.array_equal(obj, want_obj) and np.array_equal(band, want_band)
    assert list(out_ids) == ids and np.array_equal(t, lc["Time (MJD)"].to_numpy())
    obj2, _, _, _, _, ids2 = frame_columns(lc)                       # default: order of first appearance
    assert list(ids2) == list(uniq) and np.array_equal(obj2, code.astype(np.int32))
