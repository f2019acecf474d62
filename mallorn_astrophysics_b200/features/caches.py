"""The reference's feature-caching drivers as one call.

Upstream, the tables the v34a / v92d trainers read are produced by separate scripts, each re-loading the data and
re-grouping the light curves per extractor:

=================================  ==========================================================  ======================
file under ``data/processed/``     written by                                                  content
=================================  ==========================================================  ======================
``features_v4_cache.pkl``          ``scripts/train_v4_physics.py:53-109``                      ``{'train_features', 'test_features'}``: statistical (+ metadata columns) left-merged with colors, shape, physics on ``object_id``
``tde_physics_cache.pkl``          ``scripts/train_v7_tde_physics.py:79-100``                  ``{'train', 'test'}``
``multiband_gp_cache.pkl``         ``scripts/train_v19_multiband_gp.py:92-111``                ``{'train', 'test'}``
``bazin_features_cache.pkl``       ``scripts/cache_bazin_features.py:29-45``                   ``{'train', 'test'}``
``enhanced_colors_cache.pkl``      ``scripts/precompute_features.py:40-56``                    ``{'train', 'test'}``
``gp_features_cache.pkl``          ``scripts/precompute_features.py:58-76``                    ``{'train', 'test'}``
``powerlaw_features.pkl``          ``scripts/visualize_and_powerlaw.py:374-375`` (train only)  DataFrame
=================================  ==========================================================  ======================

``write_feature_caches`` packs each split once (K1) and runs every pass on the device-resident store
(``pipeline.extract_feature_tables``), then writes the same files with the same structures, so
``train_v34a_bazin.py:54-86`` / ``train_v92_focal_adversarial.py`` load them unchanged.  Existing files are left alone
unless ``overwrite=True`` -- the reference's drivers also treat a present cache as final.
"""
from __future__ import annotations

import os
import pickle
from pathlib import Path
from typing import Dict, Optional, Sequence

import pandas as pd

from . import pipeline
from .statistical import add_metadata_features

CACHE_FILES = ("features_v4_cache.pkl", "tde_physics_cache.pkl", "multiband_gp_cache.pkl", "bazin_features_cache.pkl",
               "enhanced_colors_cache.pkl", "gp_features_cache.pkl", "powerlaw_features.pkl")


def merged_v4_table(tables: Dict[str, pd.DataFrame], metadata: pd.DataFrame) -> pd.DataFrame:
    """``extract_all_features_v4`` (train_v4_physics.py:64-80): stat + metadata, then colors, shape, physics, left-merged."""
    out = add_metadata_features(tables["statistical"], metadata)
    for m in ("colors", "shape", "physics"):
        out = out.merge(tables[m], on="object_id", how="left")
    return out


def _split_files(data_path: Path, split: str):
    files = [data_path / f"split_{i:02d}" / f"{split}_full_lightcurves.csv" for i in range(1, 21)]      # data_loader.py:49-57
    files = [f for f in files if f.exists()]
    if not files:
        raise FileNotFoundError(f"No {split} lightcurve files found")
    return files


def write_feature_caches_from_files(data_path, out_dir, overwrite: bool = False, files: Sequence[str] = CACHE_FILES,
                                    device: int = 0, threads: int = 0) -> Dict[str, Path]:
    """``write_feature_caches`` from the raw data directory (``train_log.csv``, ``test_log.csv``, ``split_XX/*_full_lightcurves.csv``;
    ``utils/data_loader.py:21-62``) without building the long DataFrames: the split files go through the native reader
    straight into K1.  Same files, same structures."""
    data_path = Path(data_path)
    train_meta = pd.read_csv(data_path / "train_log.csv")                # data_loader.py:31-32
    test_meta = pd.read_csv(data_path / "test_log.csv")
    return _write(lambda mods: pipeline.extract_feature_tables_from_csv(_split_files(data_path, "train"), train_meta,
                                                                        train_meta["object_id"].tolist(), modules=mods,
                                                                        device=device, threads=threads),
                  lambda mods: pipeline.extract_feature_tables_from_csv(_split_files(data_path, "test"), test_meta,
                                                                        test_meta["object_id"].tolist(), modules=mods,
                                                                        device=device, threads=threads),
                  train_meta, test_meta, out_dir, overwrite, files)


def write_feature_caches(train_lc: pd.DataFrame, train_meta: pd.DataFrame, test_lc: pd.DataFrame, test_meta: pd.DataFrame,
                         out_dir, overwrite: bool = False, files: Sequence[str] = CACHE_FILES, device: int = 0) -> Dict[str, Path]:
    """Write the cache files named in ``files`` (default: all seven) into ``out_dir``.  Returns ``{file name: path}`` of
    the files written by this call."""
    return _write(lambda mods: pipeline.extract_feature_tables(train_lc, train_meta, train_meta["object_id"].tolist(), modules=mods, device=device),
                  lambda mods: pipeline.extract_feature_tables(test_lc, test_meta, test_meta["object_id"].tolist(), modules=mods, device=device),
                  train_meta, test_meta, out_dir, overwrite, files)


def _write(run_train, run_test, train_meta, test_meta, out_dir, overwrite, files) -> Dict[str, Path]:
    unknown = [f for f in files if f not in CACHE_FILES]
    if unknown:
        raise KeyError(f"unknown cache file(s) {unknown}; known: {list(CACHE_FILES)}")
    out_dir = Path(out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    todo = [f for f in files if overwrite or not (out_dir / f).exists()]
    if not todo:
        return {}
    need = {"features_v4_cache.pkl": ("statistical", "colors", "shape", "physics"), "tde_physics_cache.pkl": ("tde",),
            "multiband_gp_cache.pkl": ("gp2d",), "bazin_features_cache.pkl": ("bazin",), "enhanced_colors_cache.pkl": ("colors",),
            "gp_features_cache.pkl": ("gp1d",), "powerlaw_features.pkl": ("powerlaw",)}
    test_modules = [m for m in pipeline.ALL_MODULES if any(m in need[f] for f in todo if f != "powerlaw_features.pkl")]
    train_modules = [m for m in pipeline.ALL_MODULES if any(m in need[f] for f in todo)]
    train = run_train(train_modules)
    test = run_test(test_modules) if test_modules else {}
    written = {}

    def dump(name, obj, how):
        path = out_dir / name
        tmp = out_dir / (name + ".tmp")
        if how == "pandas":
            pd.to_pickle(obj, tmp)
        else:
            with open(tmp, "wb") as fh:
                pickle.dump(obj, fh)
        os.replace(tmp, path)
        written[name] = path

    pair = {"tde_physics_cache.pkl": ("tde", "pandas"), "multiband_gp_cache.pkl": ("gp2d", "pickle"),
            "bazin_features_cache.pkl": ("bazin", "pandas"), "enhanced_colors_cache.pkl": ("colors", "pickle"),
            "gp_features_cache.pkl": ("gp1d", "pickle")}
    for name in todo:
        if name == "features_v4_cache.pkl":
            dump(name, {"train_features": merged_v4_table(train, train_meta), "test_features": merged_v4_table(test, test_meta)}, "pandas")
        elif name == "powerlaw_features.pkl":
            dump(name, train["powerlaw"], "pandas")
        else:
            module, how = pair[name]
            dump(name, {"train": train[module], "test": test[module]}, how)
    return written
