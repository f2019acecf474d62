"""Drop-in for the reference's ``src/features/research_features.py`` (v115c block).

``extract_research_features(lightcurves, object_ids, metadata_df=None, verbose=True)`` keeps the reference signature
(research_features.py:566; note the argument order: ids before the metadata frame) and the 40 column names.  The
arithmetic (``csrc/research.cuh``) is held to the reference's own frames on the CPU (host instantiation) and on the B200
(``tests/test_gpu_research.py``); the pass is part of ``features.pipeline.ALL_MODULES``.
Redshifts come from ``metadata_df['Z']``; without a metadata frame (or for ``Z <= 0`` / missing ids) the five luminosity
columns are NaN, as upstream (:553-560).
"""
from __future__ import annotations

from typing import List, Optional

import pandas as pd

from ._frames import run_frame_module


def extract_research_features(lightcurves: pd.DataFrame, object_ids: List[str], metadata_df: Optional[pd.DataFrame] = None,
                              verbose: bool = True) -> pd.DataFrame:
    if verbose:
        print(f"    Research Features: {len(object_ids)} objects on the GPU", flush=True)
    return run_frame_module("research", lightcurves, object_ids, metadata=metadata_df)
