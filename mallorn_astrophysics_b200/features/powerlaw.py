"""Batch form of the reference's script-level decline-model features.

The reference has no module for these: ``extract_powerlaw_features(obj_id, lc_data)`` lives in
``scripts/train_v55_powerlaw.py:196-202`` (twin ``scripts/visualize_and_powerlaw.py:349-358``) and the
v55 driver reads the train rows from ``data/processed/powerlaw_features.pkl`` (:99).  The GPU path plugs in
by producing that frame: ``extract_powerlaw_features(lightcurves, object_ids)`` returns ``object_id`` plus
the 27 ``{g,r,i}_{model}_r2`` columns in the reference's order (27 bounded fits per object,
``csrc/fits.cuh``).  ``write_powerlaw_cache`` writes the pickle the driver expects.
"""
from __future__ import annotations

from typing import List, Optional

import pandas as pd

from ._frames import run_frame_module


def extract_powerlaw_features(lightcurves: pd.DataFrame, object_ids: Optional[List[str]] = None) -> pd.DataFrame:
    return run_frame_module("powerlaw", lightcurves, object_ids)


def write_powerlaw_cache(lightcurves: pd.DataFrame, object_ids, path) -> pd.DataFrame:
    """``data/processed/powerlaw_features.pkl`` as written by visualize_and_powerlaw.py:374-375."""
    df = extract_powerlaw_features(lightcurves, object_ids)
    df.to_pickle(path)
    return df
