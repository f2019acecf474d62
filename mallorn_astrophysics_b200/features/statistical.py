"""Drop-in for the reference's ``src/features/statistical.py`` served from the B200.

Same public names and signatures: ``extract_statistical_features`` (statistical.py:135-226),
``add_metadata_features`` (:229-253), ``LSST_BANDS`` (:38).  The per-object loop of the
reference is replaced by the packed store (K1) plus the warp-per-object moment pass (K2,
``csrc/stat.cuh``).
"""
from __future__ import annotations

from typing import List, Optional

import pandas as pd

from ._compat import schema as _schema
_BANDS = _schema.LSST_BANDS
from ._frames import run_frame_module

LSST_BANDS = list(_BANDS)


def extract_statistical_features(lightcurves: pd.DataFrame, object_ids: Optional[List[str]] = None,
                                 bands: List[str] = LSST_BANDS) -> pd.DataFrame:
    """One row per object: 17 statistics x (6 bands + all) + 3 flux ratios + peak_band.

    ``bands`` must be the six LSST bands in wavelength order (the only value the reference's
    callers ever pass); anything else is a structural error.
    """
    if list(bands) != LSST_BANDS:
        raise ValueError("the CUDA moment pass is compiled for bands ['u','g','r','i','z','y']")
    return run_frame_module("statistical", lightcurves, object_ids)


def add_metadata_features(features: pd.DataFrame, metadata: pd.DataFrame) -> pd.DataFrame:
    """Left-merge Z and EBV and derive ``luminosity_distance`` / ``time_dilation`` (host-side; four columns)."""
    result = features.merge(metadata[["object_id", "Z", "EBV"]], on="object_id", how="left")
    result["luminosity_distance"] = result["Z"] * 4280
    result["time_dilation"] = 1 + result["Z"]
    return result
