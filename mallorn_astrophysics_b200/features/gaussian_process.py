"""Drop-in for the reference's ``src/features/gaussian_process.py`` served from the B200.

``extract_gp_features(lightcurves, metadata, object_ids=None, verbose=True)`` keeps the reference signature
(gaussian_process.py:251); ``metadata`` is accepted and unused, as in the reference.  The per-band 1-D GPs
(g, r, i, z; ``ConstantKernel * RBF + WhiteKernel`` with the observation variances on the diagonal, four bounded
L-BFGS-B starts each, the best one kept -- scikit-learn's ``GaussianProcessRegressor`` in the reference,
:88-104) run one warp per object in ``csrc/gp1d.cu``; no scikit-learn dependency.
``scripts/precompute_features.py:58-76`` calls this unchanged.
"""
from __future__ import annotations

from typing import List, Optional

import pandas as pd

from ._compat import schema as _schema
LSST_BANDS = list(_schema.LSST_BANDS)                  # gaussian_process.py:26
from ._frames import run_frame_module


def extract_gp_features(lightcurves: pd.DataFrame, metadata: Optional[pd.DataFrame] = None,
                        object_ids: Optional[List[str]] = None, verbose: bool = True) -> pd.DataFrame:
    if verbose:
        n = len(object_ids) if object_ids is not None else lightcurves["object_id"].nunique()
        print(f"    GP features: {n} objects on the GPU", flush=True)
    return run_frame_module("gp1d", lightcurves, object_ids)
