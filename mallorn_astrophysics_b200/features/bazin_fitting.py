"""Drop-in for the reference's ``src/features/bazin_fitting.py`` served from the B200.

Same public names: ``bazin_function`` (bazin_fitting.py:37-60, a plain formula kept on the host for
callers that plot fitted curves) and ``extract_bazin_features(lightcurves, object_ids=None)``
(:254-288).  The 6 bounded Trust-Region-Reflective fits per object run on the GPU
(``csrc/trf.cuh`` + ``csrc/fits.cuh``, one warp per object); ``scripts/cache_bazin_features.py`` and
``scripts/train_v34a_bazin.py:99-107`` call this unchanged.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import pandas as pd

from ._compat import schema as _schema
_BANDS = _schema.LSST_BANDS
from ._frames import run_frame_module

LSST_BANDS = list(_BANDS)


def bazin_function(t, A, t0, tau_rise, tau_fall, B):
    """f(t) = A exp(-(t-t0)/tau_fall) / (1 + exp(-(t-t0)/tau_rise)) + B."""
    t = np.asarray(t, dtype=float)
    return A * np.exp(-(t - t0) / tau_fall) / (1.0 + np.exp(-(t - t0) / tau_rise)) + B


def extract_bazin_features(lightcurves: pd.DataFrame, object_ids: Optional[List[str]] = None) -> pd.DataFrame:
    return run_frame_module("bazin", lightcurves, object_ids)
