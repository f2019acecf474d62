"""Drop-in for the reference's ``src/features/tde_physics.py`` served from the B200.

``extract_tde_physics_features(lightcurves, object_ids=None)`` keeps the reference signature and frame schema
(tde_physics.py:377); the per-object Python loop is replaced by the packed store (K1) and the
warp-per-object scalar pass in ``csrc/tde.cuh``.
"""
from __future__ import annotations

from typing import List, Optional

import pandas as pd

from ._compat import schema as _schema
_BANDS = _schema.LSST_BANDS
from ._frames import run_frame_module

LSST_BANDS = list(_BANDS)


def extract_tde_physics_features(lightcurves: pd.DataFrame, object_ids: Optional[List[str]] = None) -> pd.DataFrame:
    return run_frame_module("tde", lightcurves, object_ids)
