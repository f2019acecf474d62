"""DataFrame assembly of the drop-in extractors (features/_frames.py) on the CPU: the packed store is replaced by a
stand-in whose rows come from the CPU oracle, so everything around the kernels -- request order, omitted ids, object_id
position, integer dtypes, redshift lookup, the CSV entry -- is checked against the reference's own frames without a GPU."""
import warnings

import numpy as np
import pandas as pd
import pytest

from mallorn_astrophysics_b200 import schema
from mallorn_astrophysics_b200.features import _frames
from oracle.extract import run_module

BANDS = np.array(["u", "g", "r", "i", "z", "y", "X"], dtype=object)


class OracleStore:
    """Duck-types LightCurveStore for frames_from_store: rows from the oracle, per dense object index."""

    def __init__(self, obj, band, t, flux, err, n_obj, device=0, stream=0):
        keep = obj >= 0
        self.cols = (obj[keep].astype(np.int32), np.asarray(band)[keep], np.asarray(t)[keep], np.asarray(flux)[keep], np.asarray(err)[keep])
        self.n_obj = int(n_obj)
        self.closed = False

    @classmethod
    def from_frame(cls, lightcurves, object_ids=None, device=0, stream=0):
        from mallorn_astrophysics_b200.store import frame_columns
        obj, band, t, f, e, ids = frame_columns(lightcurves, object_ids)
        return cls(obj, band, t, f, e, len(ids)), ids

    def features(self, module, z=None):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            present, rows = run_module(module, *self.cols, self.n_obj, z)
        out = np.full((self.n_obj, rows.shape[1]), np.nan)
        out[present] = rows
        if module == "statistical":                       # the kernel writes n_obs = 0 / peak_band = -1 for empty objects
            empty = np.setdiff1d(np.arange(self.n_obj), present)
            out[empty] = 0.0
        return out

    def segments(self):
        seg = np.zeros((self.n_obj, 22), np.int32)
        np.add.at(seg, (self.cols[0], 3 * np.minimum(self.cols[1].astype(int), 6)), 1)
        return seg

    def close(self):
        self.closed = True


@pytest.fixture()
def oracle_store(monkeypatch):
    monkeypatch.setattr(_frames, "LightCurveStore", OracleStore)
    return OracleStore


def _frames_of(g):
    n_obj = int(g["n_obj"])
    ids = np.array([f"Obj_{k:04d}_name" for k in range(n_obj)], dtype=object)
    lc = pd.DataFrame({"object_id": ids[g["obj"]], "Time (MJD)": g["t"], "Flux": g["flux"], "Flux_err": g["err"],
                       "Filter": BANDS[np.minimum(g["band"], 6)]})
    meta = pd.DataFrame({"object_id": ids, "Z": g["z"], "EBV": g["ebv"]})
    return lc, meta, ids


@pytest.mark.parametrize("module", ["statistical", "colors", "shape", "tde", "physics"])
def test_frames_match_the_reference_frames(golden, oracle_store, module):
    g = golden["ref_edge"]                                   # object 20 has no rows, object 15 has tied times
    lc, meta, ids = _frames_of(g)
    request = [i for i in ids[::-1] if i != ids[15]]
    df = _frames.run_frame_modules([module], lc, request, meta)[module]
    cols = schema.MODULE_COLUMNS[module]
    assert list(df.columns) == (["object_id"] + cols if module == "statistical" else cols + ["object_id"])
    assert list(df["object_id"]) == [i for i in request if i != ids[20]]
    if module == "statistical":
        assert all(str(df[c].dtype).startswith("int") for c in schema.STAT_INT_COLUMNS)
    ref_rows, ref_present = g[f"{module}_rows"], g[f"{module}_present"]
    order = [int(np.flatnonzero(ids == i)[0]) for i in df["object_id"]]
    want = np.array([ref_rows[np.flatnonzero(ref_present == k)[0]] for k in order])
    got = df[cols].to_numpy(dtype=float)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    np.testing.assert_allclose(got[ok], want[ok], rtol=1e-12)


def test_csv_entry_equals_the_frame_entry(golden, oracle_store, tmp_path, monkeypatch):
    import __graft_entry__ as entry
    entry.build()
    g = golden["ref_sparse"]
    lc, meta, ids = _frames_of(g)
    lc = lc[lc["Filter"] != "X"]
    paths = []
    for i, part in enumerate((lc.iloc[: len(lc) // 3], lc.iloc[len(lc) // 3:])):
        p = tmp_path / f"p{i}.csv"
        part.to_csv(p, index=False)
        paths.append(p)
    frame = pd.concat([pd.read_csv(p) for p in paths], ignore_index=True)
    for request in (None, list(ids[10:30][::-1]) + ["Not_there"]):
        a = _frames.run_frame_modules(["statistical", "physics"], frame, request, meta)
        b = _frames.run_csv_modules(["statistical", "physics"], paths, request, meta)
        for m in a:
            pd.testing.assert_frame_equal(a[m], b[m])


def test_public_csv_pipeline_checks_its_arguments(golden, oracle_store, tmp_path):
    from mallorn_astrophysics_b200.features import pipeline
    g = golden["ref_sparse"]
    lc, meta, ids = _frames_of(g)
    p = tmp_path / "lc.csv"
    lc[lc["Filter"] != "X"].to_csv(p, index=False)
    with pytest.raises(KeyError):
        pipeline.extract_feature_tables_from_csv(p, meta, modules=["nope"])
    with pytest.raises(ValueError):
        pipeline.extract_feature_tables_from_csv(p, None, modules=["physics"])
    t = pipeline.extract_feature_tables_from_csv(p, meta, list(ids[:7]), modules=["colors", "physics"])
    assert list(t) == ["colors", "physics"] and list(t["colors"]["object_id"]) == list(ids[:7])
