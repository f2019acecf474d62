"""One pack, every pass: the feature-caching drivers of the reference in a single call.

``scripts/precompute_features.py:31-76``, ``cache_bazin_features.py:44-60`` and the v19 / v55 drivers each call one
``extract_*_features`` after another on the same ``train_lc`` / ``test_lc`` frame, and every extractor rebuilds its
own ``groupby('object_id')`` dict.  The drop-in extractors of this package keep that interface (one K1 pack per
call); ``extract_feature_tables`` is the batch form for a caching driver: the light curves are factorised and packed
into HBM once and every requested pass runs on the same device-resident store.

    tables = extract_feature_tables(train_lc, train_meta, train_ids)
    tables["bazin"]      # == extract_bazin_features(train_lc, train_ids)
    tables["gp2d"]       # == extract_multiband_gp_features(train_lc, train_meta, train_ids)
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import pandas as pd

from ._frames import run_csv_modules, run_frame_modules

ALL_MODULES = ("statistical", "colors", "shape", "physics", "tde", "bazin", "powerlaw", "gp2d", "gp1d", "research")


def extract_feature_tables(lightcurves: pd.DataFrame, metadata: Optional[pd.DataFrame] = None,
                           object_ids: Optional[List[str]] = None,
                           modules: Sequence[str] = ALL_MODULES, device: int = 0) -> Dict[str, pd.DataFrame]:
    """``{module: DataFrame}`` with exactly the frames the per-module extractors return (same columns, same row
    order, ids without observations omitted).  ``metadata`` is needed only for ``physics`` (redshift)."""
    modules = list(modules)
    unknown = [m for m in modules if m not in ALL_MODULES]
    if unknown:
        raise KeyError(f"unknown feature module(s) {unknown}; known: {list(ALL_MODULES)}")
    if ("physics" in modules) and metadata is None:
        raise ValueError("the physics module needs the metadata frame (object_id, Z)")
    return run_frame_modules(modules, lightcurves, object_ids, metadata, device)


def extract_feature_tables_from_csv(paths, metadata: Optional[pd.DataFrame] = None, object_ids: Optional[List[str]] = None,
                                    modules: Sequence[str] = ALL_MODULES, device: int = 0, threads: int = 0) -> Dict[str, pd.DataFrame]:
    """The same tables straight from light-curve CSV files (``split_XX/{train,test}_full_lightcurves.csv``): the native
    reader (``ingest.read_lightcurve_csv``, bit-identical to ``pd.read_csv`` + ``pd.concat``) feeds K1 without building the
    long DataFrame.  ``object_ids`` default: every object of the files in order of first appearance."""
    modules = list(modules)
    unknown = [m for m in modules if m not in ALL_MODULES]
    if unknown:
        raise KeyError(f"unknown feature module(s) {unknown}; known: {list(ALL_MODULES)}")
    if ("physics" in modules) and metadata is None:
        raise ValueError("the physics module needs the metadata frame (object_id, Z)")
    return run_csv_modules(modules, paths, object_ids, metadata, device, threads)
