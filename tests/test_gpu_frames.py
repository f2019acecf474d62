"""The drop-in boundary itself: the reference's function names, signatures and DataFrame schema
(SURVEY.md section 8b), served from the GPU through ctypes -> libmallorn_b200.so."""
import importlib
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BANDS = np.array(["u", "g", "r", "i", "z", "y", "X"], dtype=object)


def _frames(g):
    import pandas as pd
    n_obj = int(g["n_obj"])
    ids = np.array([f"Obj_{k:04d}_name" for k in range(n_obj)], dtype=object)
    lc = pd.DataFrame({"object_id": ids[g["obj"]], "Time (MJD)": g["t"], "Flux": g["flux"], "Flux_err": g["err"],
                       "Filter": BANDS[np.minimum(g["band"], 6)]})
    meta = pd.DataFrame({"object_id": ids, "Z": g["z"], "EBV": g["ebv"]})
    return lc, meta, ids


@pytest.fixture(scope="module")
def features_pkg():
    """`import features` exactly as a reference driver would after the one-line sys.path switch."""
    sys.path.insert(0, os.path.join(ROOT, "mallorn_astrophysics_b200"))
    try:
        for name in [m for m in sys.modules if m == "features" or m.startswith("features.")]:
            del sys.modules[name]
        yield importlib.import_module("features")
    finally:
        sys.path.remove(os.path.join(ROOT, "mallorn_astrophysics_b200"))


CASES = [("statistical", "features.statistical", "extract_statistical_features", False),
         ("colors", "features.colors", "extract_color_features", False),
         ("shape", "features.lightcurve_shape", "extract_shape_features", False),
         ("tde", "features.tde_physics", "extract_tde_physics_features", False),
         ("physics", "features.physics_based", "extract_physics_features", True),
         ("bazin", "features.bazin_fitting", "extract_bazin_features", False),
         ("gp2d", "features.multiband_gp", "extract_multiband_gp_features", True),
         ("gp1d", "features.gaussian_process", "extract_gp_features", True),
         ("powerlaw", "features.powerlaw", "extract_powerlaw_features", False)]


@pytest.mark.parametrize("module,modname,fn,needs_meta", CASES)
def test_extractors_keep_the_reference_schema(golden, features_pkg, module, modname, fn, needs_meta):
    from mallorn_astrophysics_b200 import schema
    g = golden["ref_edge"]                                  # object 20 has no rows: must be omitted
    lc, meta, ids = _frames(g)
    lc_before = lc.copy()
    extract = getattr(importlib.import_module(modname), fn)
    request = list(ids[::-1])                               # rows come back in REQUEST order
    df = extract(lc, meta, request) if needs_meta else extract(lc, request)
    cols = schema.MODULE_COLUMNS[module]
    if module in ("statistical", "powerlaw"):
        assert list(df.columns) == ["object_id"] + cols     # statistical.py:166, train_v55_powerlaw.py:197
    else:
        assert list(df.columns) == cols + ["object_id"]     # e.g. colors.py:377
    if module == "powerlaw":
        # one row per REQUESTED id, NaN where nothing can be fitted: the reference loops over the ids (train_v55_powerlaw.py:196-202)
        assert list(df["object_id"]) == request and df[cols].iloc[request.index(ids[20])].isna().all()
    else:
        present = [i for i in request if i != ids[20]]
        assert list(df["object_id"]) == present             # absent ids are omitted (statistical.py:163-165)
    assert lc.equals(lc_before)                             # inputs are not mutated
    if module == "statistical":
        assert all(str(df[c].dtype).startswith("int") for c in schema.STAT_INT_COLUMNS)
    if module in ("statistical", "colors", "shape", "tde", "physics"):
        ref_rows = g[f"{module}_rows"]
        ref_present = g[f"{module}_present"]
        order = [int(np.flatnonzero(ids == i)[0]) for i in df["object_id"]]
        want = np.array([ref_rows[np.flatnonzero(ref_present == k)[0]] for k in order])
        got = df[cols].to_numpy(dtype=float)
        keep = np.array(order) != 15                        # the tie object (unstable sort upstream)
        assert np.array_equal(np.isnan(got[keep]), np.isnan(want[keep]))
        ok = ~np.isnan(want) & keep[:, None]
        np.testing.assert_allclose(got[ok], want[ok], rtol=1e-5, atol=1e-9)


def test_default_object_ids_and_metadata_merge(golden, features_pkg):
    g = golden["ref_sparse"]
    lc, meta, ids = _frames(g)
    st = importlib.import_module("features.statistical")
    df = st.extract_statistical_features(lc)                # object_ids=None -> unique() order of the light curves
    assert list(df["object_id"]) == list(lc["object_id"].unique())
    full = st.add_metadata_features(df, meta)               # statistical.py:229-253
    assert list(full.columns[-4:]) == ["Z", "EBV", "luminosity_distance", "time_dilation"]
    z = meta.set_index("object_id").loc[full["object_id"], "Z"].to_numpy()
    np.testing.assert_allclose(full["luminosity_distance"], 4280 * z)
    np.testing.assert_allclose(full["time_dilation"], 1 + z)
    assert importlib.import_module("features").extract_statistical_features is st.extract_statistical_features


def test_one_pack_pipeline_returns_the_per_module_frames(golden, features_pkg):
    """features.pipeline.extract_feature_tables (one K1 pack, all nine passes) == the per-module extractors"""
    g = golden["ref_sparse"]
    lc, meta, ids = _frames(g)
    pipe = importlib.import_module("features.pipeline")
    request = list(ids[5:30])
    tables = pipe.extract_feature_tables(lc, meta, request)
    assert set(tables) == set(pipe.ALL_MODULES)
    for module, modname, fn, needs_meta in CASES:
        extract = getattr(importlib.import_module(modname), fn)
        kw = {"verbose": False} if module in ("gp2d", "gp1d") else {}
        df = extract(lc, meta, request, **kw) if needs_meta else extract(lc, request)
        assert list(df.columns) == list(tables[module].columns)
        assert df.equals(tables[module]), module          # same kernels on the same packed rows: bit-identical
    with pytest.raises(KeyError):
        pipe.extract_feature_tables(lc, meta, request, modules=["nope"])


def test_store_from_csv_equals_store_from_the_pandas_frame(golden, tmp_path):
    """files -> native ingest -> K1 (no DataFrame) gives the rows pd.read_csv -> from_frame gives, bit for bit"""
    import pandas as pd
    from mallorn_astrophysics_b200.store import LightCurveStore
    g = golden["ref_sparse"]
    lc, meta, ids = _frames(g)
    lc = lc[lc["Filter"] != "X"]
    paths = []
    half = len(lc) // 2
    for i, part in enumerate((lc.iloc[:half], lc.iloc[half:])):
        p = tmp_path / f"part{i}.csv"
        part.to_csv(p, index=False)
        paths.append(p)
    frame = pd.concat([pd.read_csv(p) for p in paths], ignore_index=True)
    s1, ids1 = LightCurveStore.from_frame(frame)
    s2, ids2 = LightCurveStore.from_csv(paths)
    with s1, s2:
        assert list(ids1) == list(ids2)
        assert np.array_equal(s1.segments(), s2.segments())
        for m in ("statistical", "bazin"):
            assert np.array_equal(s1.features(m), s2.features(m), equal_nan=True), m


def test_overlapped_file_pipeline_equals_the_one_shot_path(golden, tmp_path):
    """ingest.featurise_files (parse of file i+1 overlapped with the staging and H2D of file i, one K1 pack at the end) gives
    the table the one-shot path (read all files, then pack) gives, bit for bit, with the same id numbering -- also when
    an object's rows are spread over several files."""
    import torch
    from mallorn_astrophysics_b200 import ingest
    from mallorn_astrophysics_b200.store import LightCurveStore
    g = golden["ref_sparse"]
    lc, meta, ids = _frames(g)
    lc = lc[lc["Filter"] != "X"]
    edges = np.linspace(0, len(lc), 4).astype(int)            # three files cutting through objects (rows are in shuffled file order)
    paths = []
    for i in range(3):
        p = tmp_path / f"split_{i}.csv"
        lc.iloc[edges[i]:edges[i + 1]].to_csv(p, index=False)
        paths.append(p)
    modules = ("statistical", "colors", "physics", "bazin")
    table, got_ids, info = ingest.featurise_files(paths, modules, metadata=meta, chunk_bytes=1)      # one group per file
    s, want_ids = LightCurveStore.from_csv(paths)
    with s:
        z = meta.set_index("object_id")["Z"].reindex(list(want_ids)).to_numpy()
        want = np.concatenate([s.features(m, z=z if m == "physics" else None) for m in modules], axis=1)
    assert list(got_ids) == list(want_ids) and info["files"] == 3 and info["groups"] == 3 and info["rows"] == len(lc)
    assert np.array_equal(table.cpu().numpy(), want, equal_nan=True)


def test_cache_driver_writes_what_the_trainers_read(golden, features_pkg, tmp_path):
    """features.caches.write_feature_caches on the GPU: every file loads with the reference's structure and holds the
    frames the per-module extractors return"""
    import pandas as pd
    g = golden["ref_sparse"]
    lc, meta, ids = _frames(g)
    train_meta, test_meta = meta.iloc[:30].reset_index(drop=True), meta.iloc[30:].reset_index(drop=True)
    train_lc = lc[lc["object_id"].isin(set(train_meta["object_id"]))]
    test_lc = lc[lc["object_id"].isin(set(test_meta["object_id"]))]
    caches = importlib.import_module("features.caches")
    written = caches.write_feature_caches(train_lc, train_meta, test_lc, test_meta, tmp_path)
    assert set(written) == set(caches.CACHE_FILES)
    bz = pd.read_pickle(tmp_path / "bazin_features_cache.pkl")
    direct = importlib.import_module("features.bazin_fitting").extract_bazin_features(test_lc, test_meta["object_id"].tolist())
    assert direct.equals(bz["test"])
    v4 = pd.read_pickle(tmp_path / "features_v4_cache.pkl")["train_features"]
    assert len(v4) == 30 and v4.columns[0] == "object_id" and "luminosity_distance" in v4.columns and "stetson_j_gr" in v4.columns
    assert len(pd.read_pickle(tmp_path / "powerlaw_features.pkl")) == 30
    assert caches.write_feature_caches(train_lc, train_meta, test_lc, test_meta, tmp_path) == {}      # present -> untouched
