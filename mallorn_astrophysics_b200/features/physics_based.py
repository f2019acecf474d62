"""Drop-in for the reference's ``src/features/physics_based.py`` served from the B200.

``extract_physics_features(lightcurves, metadata, object_ids=None)`` keeps the reference signature
(physics_based.py:461); ``metadata`` supplies ``Z`` per ``object_id`` for the rest-frame
timescales (:481, :496).  Compute: ``csrc/physics.cuh``.
"""
from __future__ import annotations

from typing import List, Optional

import pandas as pd

from ._compat import schema as _schema
_BANDS = _schema.LSST_BANDS
BAND_WAVELENGTHS_A = _schema.BAND_WAVELENGTHS_A
from ._frames import run_frame_module

LSST_BANDS = list(_BANDS)


def extract_physics_features(lightcurves: pd.DataFrame, metadata: pd.DataFrame,
                             object_ids: Optional[List[str]] = None) -> pd.DataFrame:
    return run_frame_module("physics", lightcurves, object_ids, metadata=metadata)
