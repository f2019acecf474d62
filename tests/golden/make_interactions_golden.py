"""Generates tests/golden/interactions.npz from the UNMODIFIED reference (run in the build container, where
/root/reference exists):  python tests/golden/make_interactions_golden.py

Two cases: every input of create_physics_interactions present (names as the reference spells them, including the ones no
extractor produces, e.g. g_r_at_peak), and the subset that the 434-column module table + metadata actually has.  Values
include NaN, +-inf, zeros and negative denominators."""
import os
import sys

import numpy as np
import pandas as pd

sys.path.insert(0, "/root/reference/src")
from features.interactions import create_physics_interactions  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ALL_INPUTS = ["Z", "g_r_at_peak", "g_r_post_20d", "g_r_post_50d", "r_i_at_peak", "gp_gr_color_20d", "gp_gr_color_50d", "gp_ri_color_20d",
              "temp_at_peak", "temp_post_50d", "g_peak_flux", "r_peak_flux", "i_peak_flux", "g_duration_50", "r_duration_50", "i_duration_50",
              "gp2d_time_scale", "gp2d_wave_scale", "g_amplitude", "r_amplitude", "i_amplitude", "g_rise_time", "r_rise_time",
              "g_fade_time_50", "r_fade_time_50", "g_r_slope_50d", "g_r_slope_100d", "g_skew", "r_skew", "i_skew", "g_std", "r_std", "i_std",
              "u_g_peak_flux_ratio", "g_r_peak_flux_ratio", "flux_p25"]
rng = np.random.default_rng(20260118)
n = 400
X = rng.normal(0, 3, size=(n, len(ALL_INPUTS)))
X[:, ALL_INPUTS.index("temp_at_peak")] = rng.uniform(-500, 150000, n)
X[:, ALL_INPUTS.index("temp_post_50d")] = rng.uniform(2000, 60000, n)
X[:, ALL_INPUTS.index("Z")] = rng.uniform(0.0, 1.0, n)
mask = rng.random(X.shape)
X[mask < 0.06] = np.nan
X[(mask > 0.06) & (mask < 0.07)] = np.inf
X[(mask > 0.07) & (mask < 0.08)] = -np.inf
X[(mask > 0.08) & (mask < 0.10)] = 0.0
X[(mask > 0.10) & (mask < 0.11)] = -1.0            # denominators of the form (x + 1)
df = pd.DataFrame(X, columns=ALL_INPUTS)
full = create_physics_interactions(df, ALL_INPUTS)
full_new = [c for c in full.columns if c not in ALL_INPUTS]
present = [c for c in ALL_INPUTS if c not in ("g_r_at_peak", "r_i_at_peak")]      # what the module table really has
part = create_physics_interactions(df[present], present)
part_new = [c for c in part.columns if c not in present]
np.savez_compressed(os.path.join(HERE, "interactions.npz"), inputs=np.array(ALL_INPUTS), X=X,
                    full_names=np.array(full_new), full=full[full_new].to_numpy(), present=np.array(present),
                    part_names=np.array(part_new), part=part[part_new].to_numpy())
print(len(full_new), "interactions with every input,", len(part_new), "with the module table's inputs")
