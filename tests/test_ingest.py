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
