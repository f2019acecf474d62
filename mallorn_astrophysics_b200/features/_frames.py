"""Shared plumbing of the drop-in extractors: DataFrame in -> packed store -> CUDA pass -> DataFrame out.

Schema conventions reproduced from the reference: one row per id of
``object_ids`` in request order, ids without light-curve rows omitted
(``src/features/statistical.py:163-165``), ``object_id`` first for the statistical frame and last
for every other frame, float64 columns (``*_n_obs`` / ``peak_band`` integer), NaN for anything
the reference could not compute.  Inputs are never mutated.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from ._compat import LightCurveStore, schema


def _request_index(ids, object_ids):
    """positions in the unique-id table for every requested id (duplicates preserved)."""
    import pandas as pd
    return ids.get_indexer(pd.Series(list(object_ids), dtype=object))


def frames_from_store(store, ids, object_ids: Sequence, modules: Sequence[str], metadata=None) -> dict:
    """Run every pass of ``modules`` on a packed store whose dense object index follows ``ids`` and assemble the
    reference's per-module DataFrames for the requested ``object_ids`` (request order, ids without observations omitted).
    Closes the store."""
    import pandas as pd
    try:
        z = None
        if "physics" in modules or ("research" in modules and metadata is not None):
            # z_lookup = dict(zip(metadata['object_id'], metadata['Z'])): the LAST duplicate wins
            # (physics_based.py:481); ids missing from the metadata get NaN (:496)
            zmap = pd.Series(metadata["Z"].to_numpy(dtype=np.float64), index=metadata["object_id"].to_numpy())
            zmap = zmap[~zmap.index.duplicated(keep="last")]
            z = zmap.reindex(ids).to_numpy(dtype=np.float64)
        rows = {}
        if "bazin" in modules and "powerlaw" in modules and hasattr(store, "fit_passes"):     # both fit passes in one overlapped call (mlrn_fits)
            rows["bazin"] = np.empty((store.n_obj, len(schema.MODULE_COLUMNS["bazin"])), dtype=np.float64)
            rows["powerlaw"] = np.empty((store.n_obj, len(schema.MODULE_COLUMNS["powerlaw"])), dtype=np.float64)
            store.fit_passes(rows["bazin"], rows["powerlaw"])
        for m in modules:
            if m not in rows:
                rows[m] = store.features(m, z=z if m in ("physics", "research") else None)
        counts = store.segments()[:, 0:21:3].sum(axis=1)
    finally:
        store.close()
    req_all = _request_index(ids, object_ids)
    req = req_all[req_all >= 0]                      # ids that are not in the store at all
    req = req[counts[req] > 0]                       # drop ids with no observations
    id_col = np.asarray(ids, dtype=object)[req]
    frames = {}
    for module in modules:
        if module == "powerlaw":
            # the decline-model table has ONE ROW PER REQUESTED ID, NaN where a band cannot be fitted: the reference builds it
            # by looping over the ids (visualize_and_powerlaw.py:360-371, train_v55_powerlaw.py:196-202), not over a groupby
            full = np.full((len(req_all), rows[module].shape[1]), np.nan)
            full[req_all >= 0] = rows[module][req_all[req_all >= 0]]
            df = pd.DataFrame(full, columns=schema.MODULE_COLUMNS[module])
            df.insert(0, "object_id", np.asarray(list(object_ids), dtype=object))        # train_v55_powerlaw.py:197
            frames[module] = df
            continue
        df = pd.DataFrame(rows[module][req], columns=schema.MODULE_COLUMNS[module])
        if module == "statistical":
            for c in schema.STAT_INT_COLUMNS:
                df[c] = df[c].astype(np.int64)
            df.insert(0, "object_id", id_col)
        else:
            df["object_id"] = id_col
        frames[module] = df
    return frames


def run_frame_modules(modules: Sequence[str], lightcurves, object_ids: Optional[Sequence] = None, metadata=None,
                      device: int = 0) -> dict:
    """Pack once (K1), run every pass of ``modules`` on the same device-resident store, return ``{module: DataFrame}``
    with the reference's per-module schema."""
    if object_ids is None:
        object_ids = lightcurves["object_id"].unique()
    object_ids = list(object_ids)
    store, ids = LightCurveStore.from_frame(lightcurves, object_ids, device=device)
    return frames_from_store(store, ids, object_ids, modules, metadata)


def run_csv_modules(modules: Sequence[str], paths, object_ids: Optional[Sequence] = None, metadata=None, device: int = 0,
                    threads: int = 0) -> dict:
    """The same from light-curve CSV files (no DataFrame on the way in): native ingest -> K1 -> passes.  ``object_ids``
    default: every object of the files, in order of first appearance."""
    import pandas as pd
    try:
        from .. import ingest
    except ImportError:                              # imported as top-level `features`
        from mallorn_astrophysics_b200 import ingest
    cols = ingest.read_lightcurve_csv(paths, threads)
    if object_ids is None:
        ids, obj = pd.Index(cols.object_ids), cols.obj
        object_ids = list(cols.object_ids)
    else:
        object_ids = list(object_ids)
        ids = pd.Index(pd.unique(pd.Series(object_ids, dtype=object)))
        remap = np.append(ids.get_indexer(pd.Index(cols.object_ids)).astype(np.int32), np.int32(-1))   # file id -> request slot
        obj = remap[cols.obj]                        # rows without an id (-1) pick the appended -1
    store = LightCurveStore(obj, cols.band, cols.t, cols.flux, cols.err, len(ids), device=device)
    return frames_from_store(store, ids, object_ids, modules, metadata)


def run_frame_module(module: str, lightcurves, object_ids: Optional[Sequence] = None, metadata=None,
                     device: int = 0):
    return run_frame_modules([module], lightcurves, object_ids, metadata, device)[module]
