"""Oracle for the feature interactions (test infrastructure; see ``oracle/__init__.py``).

Restates ``src/features/interactions.py:23-196`` (``create_physics_interactions``) on a dict of numpy columns: the same
expressions in the same insertion order; a feature is present when it is in ``feature_names``.  Pinned by
``tests/golden/interactions.npz`` (outputs of the unmodified reference, ``tests/golden/make_interactions_golden.py``)."""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np


def create(cols: Dict[str, np.ndarray], feature_names: Sequence[str]) -> Dict[str, np.ndarray]:
    """New columns only, in the reference's insertion order."""
    out: Dict[str, np.ndarray] = {}
    get = lambda name: cols[name] if name in feature_names else None
    with np.errstate(all="ignore"):
        Z = get("Z")
        if Z is not None:                                                             # :45-59
            for f in ("g_r_at_peak", "g_r_post_20d", "g_r_post_50d", "r_i_at_peak"):
                c = get(f)
                if c is not None:
                    out[f"{f}_x_Z"] = c * Z
                    out[f"{f}_div_Z"] = c / (Z + 0.1)
            for f in ("gp_gr_color_20d", "gp_gr_color_50d", "gp_ri_color_20d"):
                c = get(f)
                if c is not None:
                    out[f"{f}_x_Z"] = c * Z
        tp, tq = get("temp_at_peak"), get("temp_post_50d")                            # :65-76
        if tp is not None and tq is not None:
            out["temp_cooling_ratio"] = tq / (tp + 100)
            out["temp_drop_rate"] = (tp - tq) / 50.0
            out["temp_peak_4th"] = np.power(np.clip(tp, 0, 100000), 0.25)
        for b in ("g", "r", "i"):                                                     # :82-91
            p, d = get(f"{b}_peak_flux"), get(f"{b}_duration_50")
            if p is not None and d is not None:
                out[f"{b}_flux_duration"] = p * d
                out[f"{b}_flux_per_day"] = p / (d + 1)
        gt, gw = get("gp2d_time_scale"), get("gp2d_wave_scale")                       # :97-107
        for b in ("g", "r", "i"):
            a = get(f"{b}_amplitude")
            if gt is not None and a is not None:
                out[f"{b}_gp_amp_time"] = a * gt
            if gw is not None and a is not None:
                out[f"{b}_gp_amp_wave"] = a * gw
        for b in ("g", "r"):                                                          # :113-122
            rise, fade = get(f"{b}_rise_time"), get(f"{b}_fade_time_50")
            color = get("g_r_at_peak") if b == "g" else get("r_i_at_peak")
            if rise is not None and fade is not None and color is not None:
                out[f"{b}_asym_x_color"] = (rise / (fade + 1)) * color
        s50, s100, gp = get("g_r_slope_50d"), get("g_r_slope_100d"), get("g_r_at_peak")   # :128-139
        if s50 is not None and gp is not None:
            out["gr_peak_x_slope50"] = gp * s50
        if s100 is not None and gp is not None:
            out["gr_peak_x_slope100"] = gp * s100
        if s50 is not None and s100 is not None:
            out["gr_color_accel"] = s100 - s50
        for b in ("g", "r", "i"):                                                     # :145-150
            sk, sd = get(f"{b}_skew"), get(f"{b}_std")
            if sk is not None and sd is not None:
                out[f"{b}_skew_x_std"] = sk * sd
        ug, gr = get("u_g_peak_flux_ratio"), get("g_r_peak_flux_ratio")               # :156-161
        if ug is not None and gr is not None:
            out["ug_x_gr_ratio"] = ug * gr
        for f in ("r_skew", "g_skew", "flux_p25"):                                    # :167-170
            v = get(f)
            if v is not None:
                out[f"{f}_squared"] = v ** 2
        g_rise, r_rise = get("g_rise_time"), get("r_rise_time")                       # :178-191
        if g_rise is not None and r_rise is not None:
            out["rise_ratio_g_r"] = g_rise / (r_rise + 1)
        g_fade, r_fade = get("g_fade_time_50"), get("r_fade_time_50")
        if g_fade is not None and r_fade is not None:
            out["fade_ratio_g_r"] = g_fade / (r_fade + 1)
        if gt is not None and gw is not None:
            out["gp_time_wave_ratio"] = gt / (gw + 1e-6)
    return out
