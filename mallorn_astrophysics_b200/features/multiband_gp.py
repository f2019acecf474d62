"""Drop-in for the reference's ``src/features/multiband_gp.py`` served from the B200.

``extract_multiband_gp_features(lightcurves, metadata, object_ids=None, verbose=True)`` keeps the reference
signature (multiband_gp.py:347); ``metadata`` is accepted and unused, as in the reference.  The 2-D
(time x wavelength) Matern-3/2 GP -- kernel matrix, FP64 Cholesky, log-likelihood + gradient, L-BFGS
hyper-parameter search and the 12 predictions -- runs one CTA per object in ``csrc/gp2d.cu``; no ``george``
dependency.  ``scripts/train_v19_multiband_gp.py:92-111`` calls this unchanged.
"""
from __future__ import annotations

from typing import List, Optional

import pandas as pd

from ._compat import schema as _schema
BAND_WAVELENGTHS = _schema.BAND_WAVELENGTHS_A          # multiband_gp.py:26-29, Angstrom
from ._frames import run_frame_module


def extract_multiband_gp_features(lightcurves: pd.DataFrame, metadata: Optional[pd.DataFrame] = None,
                                  object_ids: Optional[List[str]] = None, verbose: bool = True) -> pd.DataFrame:
    if verbose:
        n = len(object_ids) if object_ids is not None else lightcurves["object_id"].nunique()
        print(f"    GP: {n} objects on the GPU", flush=True)
    return run_frame_module("gp2d", lightcurves, object_ids)
