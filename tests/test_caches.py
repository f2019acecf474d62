"""features.caches.write_feature_caches -- the reference's caching drivers as one call.  CPU test of the driver logic:
the GPU batch call is replaced by frames built from the CPU oracle (same schemas), and the files are compared with what
the reference's own scripts would write from those frames (structures of train_v4_physics.py:53-109,
cache_bazin_features.py:38-45, train_v19_multiband_gp.py:109-110, precompute_features.py:54-75)."""
import pickle
import warnings

import numpy as np
import pandas as pd
import pytest

from mallorn_astrophysics_b200 import schema
from mallorn_astrophysics_b200.features import caches, pipeline
from mallorn_astrophysics_b200.synth import make_table
from oracle.extract import run_table

FAST = ("statistical", "colors", "shape", "physics", "tde")       # oracle modules cheap enough for a CPU test


def _oracle_tables(table, modules):
    ids = np.asarray(table.object_ids, dtype=object)
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for m in modules:
            if m in FAST:
                present, rows = run_table(m, table)
            else:                                                    # fits / GPs: schema-shaped placeholder values
                present = np.arange(table.n_obj)
                rows = np.random.default_rng(len(m)).normal(size=(table.n_obj, len(schema.MODULE_COLUMNS[m])))
            df = pd.DataFrame(rows, columns=schema.MODULE_COLUMNS[m])
            if m == "statistical":
                for c in schema.STAT_INT_COLUMNS:
                    df[c] = df[c].astype(np.int64)
            if m in ("statistical", "powerlaw"):
                df.insert(0, "object_id", ids[present])
            else:
                df["object_id"] = ids[present]
            out[m] = df
    return out


@pytest.fixture()
def fake_gpu(monkeypatch):
    tables = {"train": make_table(12, seed=1), "test": make_table(9, seed=2)}
    calls = []

    def fake(lightcurves, metadata=None, object_ids=None, modules=pipeline.ALL_MODULES, device=0):
        split = "train" if len(object_ids) == 12 else "test"
        calls.append((split, tuple(modules)))
        return _oracle_tables(tables[split], modules)
    monkeypatch.setattr(pipeline, "extract_feature_tables", fake)
    return tables, calls


def test_cache_files_have_the_reference_structures(fake_gpu, tmp_path):
    tables, calls = fake_gpu
    tr, te = tables["train"], tables["test"]
    written = caches.write_feature_caches(tr.to_frame(), tr.metadata_frame(), te.to_frame(), te.metadata_frame(), tmp_path)
    assert set(written) == set(caches.CACHE_FILES) and len(calls) == 2          # one batch call per split
    assert "powerlaw" in dict(calls)["train"] and "powerlaw" not in dict(calls)["test"]   # the power-law pickle is train-only
    v4 = pd.read_pickle(tmp_path / "features_v4_cache.pkl")
    assert set(v4) == {"train_features", "test_features"}
    tf = v4["train_features"]
    assert tf.columns[0] == "object_id" and len(tf) == 12
    # exactly the reference's merge (train_v4_physics.py:65-80) of the per-module frames
    ref = _oracle_tables(tr, ("statistical", "colors", "shape", "physics"))
    want = ref["statistical"].merge(tr.metadata_frame()[["object_id", "Z", "EBV"]], on="object_id", how="left")
    want["luminosity_distance"] = want["Z"] * 4280
    want["time_dilation"] = 1 + want["Z"]
    for m in ("colors", "shape", "physics"):
        want = want.merge(ref[m], on="object_id", how="left")
    pd.testing.assert_frame_equal(tf, want)
    assert len(tf.columns) == 1 + 123 + 4 + 83 + 65 + 32
    for name, module in (("tde_physics_cache.pkl", "tde"), ("bazin_features_cache.pkl", "bazin"), ("multiband_gp_cache.pkl", "gp2d"),
                         ("enhanced_colors_cache.pkl", "colors"), ("gp_features_cache.pkl", "gp1d")):
        with open(tmp_path / name, "rb") as fh:
            d = pickle.load(fh)
        assert set(d) == {"train", "test"}
        assert list(d["train"].columns) == schema.MODULE_COLUMNS[module] + ["object_id"]
        assert len(d["train"]) == 12 and len(d["test"]) == 9
    pl = pd.read_pickle(tmp_path / "powerlaw_features.pkl")
    assert list(pl.columns) == ["object_id"] + schema.MODULE_COLUMNS["powerlaw"] and len(pl) == 12
    # what train_v34a_bazin.py:54-86 does with them
    base = v4["train_features"].merge(pd.read_pickle(tmp_path / "tde_physics_cache.pkl")["train"], on="object_id", how="left")
    assert len(base) == 12 and "temp_stability_x" in base.columns                 # colors and tde both have temp_stability


def test_existing_caches_are_kept_and_subsets_run_only_what_they_need(fake_gpu, tmp_path):
    tables, calls = fake_gpu
    tr, te = tables["train"], tables["test"]
    args = (tr.to_frame(), tr.metadata_frame(), te.to_frame(), te.metadata_frame(), tmp_path)
    w = caches.write_feature_caches(*args, files=["bazin_features_cache.pkl"])
    assert list(w) == ["bazin_features_cache.pkl"] and dict(calls) == {"train": ("bazin",), "test": ("bazin",)}
    calls.clear()
    assert caches.write_feature_caches(*args, files=["bazin_features_cache.pkl"]) == {} and not calls      # present -> untouched
    w = caches.write_feature_caches(*args, files=["bazin_features_cache.pkl", "powerlaw_features.pkl"])
    assert list(w) == ["powerlaw_features.pkl"] and calls == [("train", ("powerlaw",))]
    with pytest.raises(KeyError):
        caches.write_feature_caches(*args, files=["nope.pkl"])


def test_caches_from_raw_files_equal_caches_from_frames(tmp_path, monkeypatch):
    """write_feature_caches_from_files (train_log.csv / test_log.csv / split_XX files -> native reader -> K1) against
    write_feature_caches on the DataFrames pd.read_csv gives; the kernels are replaced by the oracle-backed store of
    tests/test_frames_cpu.py, so this runs on the CPU."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_frames_cpu import OracleStore
    from mallorn_astrophysics_b200.features import _frames
    import __graft_entry__ as entry
    entry.build()
    monkeypatch.setattr(_frames, "LightCurveStore", OracleStore)
    raw = tmp_path / "raw"
    sets = {"train": make_table(10, seed=11), "test": make_table(7, seed=12)}
    frames = {}
    for split, tb in sets.items():
        lc, meta = tb.to_frame(), tb.metadata_frame()
        meta.to_csv(raw.mkdir(exist_ok=True) or raw / f"{split}_log.csv", index=False)
        half = len(lc) // 2
        for i, part in enumerate((lc.iloc[:half], lc.iloc[half:])):
            d = raw / f"split_{i + 1:02d}"
            d.mkdir(exist_ok=True)
            part.to_csv(d / f"{split}_full_lightcurves.csv", index=False)
        frames[split] = (pd.concat([pd.read_csv(raw / f"split_{i:02d}" / f"{split}_full_lightcurves.csv") for i in (1, 2)], ignore_index=True),
                         pd.read_csv(raw / f"{split}_log.csv"))
    which = ["features_v4_cache.pkl", "tde_physics_cache.pkl"]
    a = caches.write_feature_caches_from_files(raw, tmp_path / "a", files=which)
    b = caches.write_feature_caches(frames["train"][0], frames["train"][1], frames["test"][0], frames["test"][1], tmp_path / "b", files=which)
    assert set(a) == set(b) == set(which)
    for name in which:
        da, db = pd.read_pickle(tmp_path / "a" / name), pd.read_pickle(tmp_path / "b" / name)
        assert set(da) == set(db)
        for k in da:
            pd.testing.assert_frame_equal(da[k], db[k])
    assert len(pd.read_pickle(tmp_path / "a" / "features_v4_cache.pkl")["train_features"]) == 10
