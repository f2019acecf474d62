"""Drop-in for the reference's ``src/features/interactions.py``: ``create_physics_interactions(df, feature_names)`` and
``select_top_interactions(X_train, y, original_feature_names, top_k)`` (interactions.py:23-196, :199-245).

The reference adds up to 41 element-wise products / ratios / powers of existing columns to the merged feature frame --
only those whose inputs are present, in a fixed order.  ``interaction_ops`` builds that list from the column names (host);
the arithmetic runs on the GPU over a device-resident table (``mlrn_interactions``, ``csrc/interactions.cu``):
``interactions_device`` is the form for a pipeline whose feature table already lives in HBM (``shard.featurise_shard``),
``create_physics_interactions`` keeps the reference's DataFrame-in / DataFrame-out signature."""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np
import pandas as pd

try:
    from .. import _native
except ImportError:                                  # imported as top-level `features`
    from mallorn_astrophysics_b200 import _native

MUL, DIV_OFFSET, DIFF_SCALED, CLIP_POW, SUB, SQUARE, ASYM_COLOR = range(7)


def interaction_ops(feature_names: Sequence[str]) -> List[Tuple[str, int, str, str, str, float]]:
    """``[(new column, kind, a, b, c, k)]`` in the order the reference inserts them (interactions.py:41-194); ``b`` / ``c``
    are ``None`` when unused.  A feature counts as present when it is in ``feature_names`` (the reference's ``get_feat``)."""
    have = set(feature_names)
    ops: List[Tuple[str, int, str, str, str, float]] = []

    def add(name, kind, a, b=None, c=None, k=0.0):
        if a in have and (b is None or b in have) and (c is None or c in have):
            ops.append((name, kind, a, b, c, float(k)))

    if "Z" in have:                                                                   # 1. colour x redshift (:45-59)
        for f in ("g_r_at_peak", "g_r_post_20d", "g_r_post_50d", "r_i_at_peak"):
            add(f"{f}_x_Z", MUL, f, "Z")
            add(f"{f}_div_Z", DIV_OFFSET, f, "Z", k=0.1)
        for f in ("gp_gr_color_20d", "gp_gr_color_50d", "gp_ri_color_20d"):
            add(f"{f}_x_Z", MUL, f, "Z")
    if "temp_at_peak" in have and "temp_post_50d" in have:                            # 2. temperature x time (:65-76)
        add("temp_cooling_ratio", DIV_OFFSET, "temp_post_50d", "temp_at_peak", k=100.0)
        add("temp_drop_rate", DIFF_SCALED, "temp_at_peak", "temp_post_50d", k=50.0)
        add("temp_peak_4th", CLIP_POW, "temp_at_peak", k=100000.0)
    for band in ("g", "r", "i"):                                                      # 3. amplitude x duration (:82-91)
        add(f"{band}_flux_duration", MUL, f"{band}_peak_flux", f"{band}_duration_50")
        add(f"{band}_flux_per_day", DIV_OFFSET, f"{band}_peak_flux", f"{band}_duration_50", k=1.0)
    for band in ("g", "r", "i"):                                                      # 4. GP scales x amplitude (:97-107)
        add(f"{band}_gp_amp_time", MUL, f"{band}_amplitude", "gp2d_time_scale")
        add(f"{band}_gp_amp_wave", MUL, f"{band}_amplitude", "gp2d_wave_scale")
    for band in ("g", "r"):                                                           # 5. asymmetry x colour (:113-122)
        add(f"{band}_asym_x_color", ASYM_COLOR, f"{band}_rise_time", f"{band}_fade_time_50",
            "g_r_at_peak" if band == "g" else "r_i_at_peak")
    add("gr_peak_x_slope50", MUL, "g_r_at_peak", "g_r_slope_50d")                       # 6. colour evolution (:128-139)
    add("gr_peak_x_slope100", MUL, "g_r_at_peak", "g_r_slope_100d")
    add("gr_color_accel", SUB, "g_r_slope_100d", "g_r_slope_50d")
    for band in ("g", "r", "i"):                                                      # 7. skewness x variability (:145-150)
        add(f"{band}_skew_x_std", MUL, f"{band}_skew", f"{band}_std")
    add("ug_x_gr_ratio", MUL, "u_g_peak_flux_ratio", "g_r_peak_flux_ratio")            # 8. multi-band (:156-161)
    for f in ("r_skew", "g_skew", "flux_p25"):                                        # 9. squares (:167-170)
        add(f"{f}_squared", SQUARE, f)
    add("rise_ratio_g_r", DIV_OFFSET, "g_rise_time", "r_rise_time", k=1.0)             # 10. ratios (:178-191)
    add("fade_ratio_g_r", DIV_OFFSET, "g_fade_time_50", "r_fade_time_50", k=1.0)
    add("gp_time_wave_ratio", DIV_OFFSET, "gp2d_time_scale", "gp2d_wave_scale", k=1e-6)
    return ops


class _Op(C.Structure):
    _fields_ = [("kind", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("c", C.c_int32), ("k", C.c_double)]


def interactions_device(table, column_names: Sequence[str], feature_names: Sequence[str] = None, stream: int = 0):
    """``table``: contiguous float64 CUDA tensor ``[n, len(column_names)]``.  Returns ``(new_table [n, n_new] CUDA tensor,
    new_names)`` -- the columns ``create_physics_interactions`` would append for these ``feature_names`` (default: all
    columns).  A name that occurs twice resolves to its last occurrence, as ``df[name]`` assignment order does upstream."""
    import torch
    L = _native.lib()
    _native.require_device()
    if not (table.is_cuda and table.dtype == torch.float64 and table.is_contiguous() and table.dim() == 2
            and table.shape[1] == len(column_names)):
        raise ValueError("table must be a contiguous float64 CUDA tensor with one column per name")
    pos = {name: i for i, name in enumerate(column_names)}
    ops = interaction_ops(column_names if feature_names is None else [f for f in feature_names if f in pos])
    arr = (_Op * max(len(ops), 1))()
    for i, (_, kind, a, b, c, k) in enumerate(ops):
        arr[i] = _Op(kind, pos[a], -1 if b is None else pos[b], -1 if c is None else pos[c], k)
    out = torch.empty((table.shape[0], len(ops)), dtype=torch.float64, device=table.device)
    if not hasattr(L.mlrn_interactions, "_bound"):
        L.mlrn_interactions.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
        L.mlrn_interactions._bound = True
    rc = L.mlrn_interactions(table.device.index, table.data_ptr(), int(table.shape[0]), int(table.shape[1]), C.addressof(arr),
                             len(ops), out.data_ptr(), max(len(ops), 1) if len(ops) == 0 else len(ops), int(stream))
    _native.check(rc, "mlrn_interactions")
    return out, [o[0] for o in ops]


def create_physics_interactions(df: pd.DataFrame, feature_names: List[str], device: int = 0) -> pd.DataFrame:
    """interactions.py:23-196: a copy of ``df`` with the interaction columns appended (same names, same order)."""
    import torch
    names = [n for n in dict.fromkeys(f for f in feature_names if f in df.columns)]
    ops = interaction_ops(names)
    used = list(dict.fromkeys(x for o in ops for x in (o[2], o[3], o[4]) if x is not None))
    out = df.copy()
    if not ops:
        return out
    host = np.ascontiguousarray(np.column_stack([df[c].to_numpy(dtype=np.float64) for c in used]))
    table = torch.from_numpy(host).to(f"cuda:{device}")
    new, new_names = interactions_device(table, used)
    new = new.cpu().numpy()
    for j, name in enumerate(new_names):
        out[name] = new[:, j]
    return out


def select_top_interactions(X_train: pd.DataFrame, y: np.ndarray, original_feature_names: List[str], top_k: int = 30) -> List[str]:
    """interactions.py:199-245: the ``top_k`` interaction columns by |point-biserial correlation| with the label among
    those significant at p < 0.05 (host-side selection, as upstream; columns with <= 100 finite values are skipped)."""
    from scipy.stats import pointbiserialr
    interaction_cols = [c for c in X_train.columns if c not in original_feature_names]
    rows = []
    for col in interaction_cols:
        vals = X_train[col].values
        valid = ~np.isnan(vals)
        if np.sum(valid) > 100:
            corr, pval = pointbiserialr(y[valid], vals[valid])
            rows.append({"feature": col, "abs_corr": abs(corr), "pval": pval})
    if not rows:
        print(f"   Created {len(interaction_cols)} interactions, selected 0 top features")
        return []
    corr_df = pd.DataFrame(rows).sort_values("abs_corr", ascending=False)
    top = corr_df[corr_df["pval"] < 0.05].head(top_k)["feature"].tolist()
    print(f"   Created {len(interaction_cols)} interactions, selected {len(top)} top features")
    return top
