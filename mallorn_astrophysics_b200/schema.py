"""Column contracts of the per-module feature frames.

The names and their order are the drop-in boundary: the reference's caching
drivers and XGBoost trainers select columns by name and feed ``df.values`` by
position (reference ``scripts/train_v34a_bazin.py:118-123``), so every list
here reproduces the dict-insertion order of the corresponding reference
extractor:

* statistical  -- ``src/features/statistical.py:166-222``
* colors       -- ``src/features/colors.py:150-342``
* shape        -- ``src/features/lightcurve_shape.py:204-330``
* tde physics  -- ``src/features/tde_physics.py:355-374``
* physics      -- ``src/features/physics_based.py:316-456``
* bazin        -- ``src/features/bazin_fitting.py:194-249``
* power law    -- ``scripts/train_v55_powerlaw.py:135-145,196-202``
* 2-D GP       -- ``src/features/multiband_gp.py:182-188,224-277``

The same order is hard-wired in ``csrc/layout.h`` (the device writes rows in
this order); ``tests/test_schema.py`` checks the two agree.
"""

LSST_BANDS = ["u", "g", "r", "i", "z", "y"]
BAND_CODE = {b: i for i, b in enumerate(LSST_BANDS)}
UNKNOWN_BAND = 6  # rows whose Filter is not one of ugrizy
N_SEG = 7  # segments per object in the packed store: ugrizy + unknown

# Angstrom (multiband_gp.py:26-29, physics_based.py:25-28)
BAND_WAVELENGTHS_A = {"u": 3670, "g": 4825, "r": 6222, "i": 7545, "z": 8691, "y": 9710}
# nm (colors.py:25-28, data_loader.py:122-130)
BAND_WAVELENGTHS_NM = {"u": 367.0, "g": 482.5, "r": 622.2, "i": 754.5, "z": 869.1, "y": 971.0}

STAT_NAMES = ["n_obs", "mean", "std", "min", "max", "median", "skew", "kurtosis",
              "amplitude", "mad", "iqr", "beyond_1std", "beyond_2std", "max_slope",
              "mean_snr", "time_span", "cadence_mean"]
STAT_COLUMNS = [f"{p}_{s}" for p in LSST_BANDS + ["all"] for s in STAT_NAMES] + [
    "flux_ratio_g_r", "flux_ratio_r_i", "flux_ratio_i_z", "peak_band"]
STAT_INT_COLUMNS = [f"{p}_n_obs" for p in LSST_BANDS + ["all"]] + ["peak_band"]

COLOR_PAIRS = [("g", "r"), ("r", "i"), ("u", "g"), ("i", "z")]
COLOR_EPOCHS = [("peak", 0), ("post_10d", 10), ("post_20d", 20), ("post_30d", 30),
                ("post_50d", 50), ("post_75d", 75), ("post_100d", 100),
                ("post_150d", 150), ("pre_10d", -10), ("pre_20d", -20)]
_PAIR_NAMES = [f"{a}_{b}" for a, b in COLOR_PAIRS]
COLOR_COLUMNS = (
    ["peak_mjd"]
    + [f"{p}_{e}" for e, _ in COLOR_EPOCHS for p in _PAIR_NAMES]
    + [f"{p}_{s}" for p in _PAIR_NAMES for s in ("slope_50d", "slope_100d")]
    + [f"{p}_{s}" for p in _PAIR_NAMES for s in ("std", "range")]
    + [f"{b}_peak_flux" for b in LSST_BANDS]
    + [f"{p}_peak_flux_ratio" for p in _PAIR_NAMES]
    + ["g_r_peak_lag", "r_i_peak_lag", "g_r_curvature", "r_i_curvature",
       "g_r_late_stability", "g_r_late_mean", "r_i_late_stability", "r_i_late_mean",
       "temp_peak", "temp_post_30d", "temp_post_75d", "temp_post_150d",
       "temp_slope_early", "temp_slope_mid", "temp_slope_late", "temp_stability"]
)

SHAPE_BAND_NAMES = ["rise_time", "fade_time_50", "fade_time_25", "asymmetry",
                    "duration_50", "duration_25", "power_law_alpha", "power_law_residual"]
SHAPE_COLUMNS = [f"{b}_{s}" for b in LSST_BANDS for s in SHAPE_BAND_NAMES] + [
    "peak_time_spread", "peak_time_std", "optical_mean_rise_time",
    "optical_mean_fade_time", "optical_mean_power_alpha", "rise_time_consistency",
    "fade_time_consistency", "all_rise_time", "all_fade_time_50", "all_asymmetry",
    "all_power_law_alpha", "all_power_law_residual", "flux_p10", "flux_p25",
    "flux_p75", "flux_p90", "flux_concentration"]

TDE_COLUMNS = (
    [f"{p}_{s}" for p in ("g_r", "r_i") for s in ("color_var", "color_range", "color_trend")]
    + [f"{b}_{s}" for b in "gri" for s in ("late_slope", "late_flux_ratio", "rebrightening")]
    + ["g_rise_shape", "g_rise_rate", "r_rise_shape", "r_rise_rate",
       "temp_stability", "temp_trend", "temp_late_vs_peak",
       "r_decay_alpha", "r_decay_residual", "r_decay_alpha_late"]
)

PHYSICS_COLUMNS = (
    ["stetson_j_gr", "stetson_j_ri", "stetson_j_gi",
     "stetson_k_g", "stetson_k_r", "stetson_k_i",
     "r_sf_tau_1", "r_sf_tau_5", "r_sf_tau_10", "r_sf_tau_30", "r_sf_tau_100", "r_sf_slope"]
    + [f"{b}_{s}" for b in "gri" for s in ("rest_duration", "rest_rise", "rest_fade")]
    + ["temp_at_peak", "temp_post_50d", "temp_evolution",
       "r_bazin_amplitude", "r_bazin_t0", "r_bazin_rise_approx", "r_bazin_fall_approx",
       "r_bazin_plateau", "mean_snr", "median_snr", "excess_variance"]
)

BAZIN_BAND_NAMES = ["bazin_A", "bazin_t0", "bazin_tau_rise", "bazin_tau_fall", "bazin_B",
                    "bazin_fit_chi2", "bazin_rise_fall_ratio", "bazin_peak_flux"]
BAZIN_COLUMNS = [f"{b}_{s}" for b in LSST_BANDS for s in BAZIN_BAND_NAMES] + [
    "bazin_rise_consistency", "bazin_fall_consistency", "bazin_avg_fit_chi2",
    "bazin_fit_quality_dispersion"]

POWERLAW_MODELS = ["powerlaw_5_3", "powerlaw_1", "powerlaw_1_5", "powerlaw_2",
                   "powerlaw_2_5", "powerlaw_3", "powerlaw_0_5", "exponential", "linear"]
POWERLAW_COLUMNS = [f"{b}_{m}_r2" for b in "gri" for m in POWERLAW_MODELS]

GP2D_EPOCHS = [0, 20, 50, 100]
GP2D_COLUMNS = (
    ["gp2d_amplitude", "gp2d_time_scale", "gp2d_wave_scale", "gp2d_log_likelihood",
     "gp2d_time_wave_ratio"]
    + [c for e in GP2D_EPOCHS for c in (f"gp_flux_g_{e}d", f"gp_flux_r_{e}d", f"gp_flux_i_{e}d",
                                        f"gp_gr_color_{e}d", f"gp_ri_color_{e}d")]
    + ["gp_gr_slope_50d", "gp_gr_slope_100d"]
)

GP1D_BANDS = ["g", "r", "i", "z"]                  # gaussian_process.py:187
GP1D_COLUMNS = (
    [f"{b}_{s}" for b in GP1D_BANDS for s in ("gp_length_scale", "gp_amplitude", "gp_noise", "gp_log_likelihood")]
    + ["gp_ls_ratio_gr", "gp_ls_ratio_ri", "gp_mean_length_scale", "gp_std_length_scale", "gp_mean_amplitude"]
)

RESEARCH_POWERLAW_NAMES = ["powerlaw_alpha", "powerlaw_alpha_deviation_53", "powerlaw_alpha_deviation_512", "powerlaw_chi2",
                           "powerlaw_residual_std", "powerlaw_fit_success"]            # research_features.py:59-66
RESEARCH_COLUMNS = (
    [f"{b}_{s}" for b in "gri" for s in RESEARCH_POWERLAW_NAMES]
    + ["optical_mean_powerlaw_alpha", "optical_std_powerlaw_alpha", "optical_mean_deviation_53"]
    + ["nuclear_smoothness", "nuclear_concentration", "nuclear_variability_ratio", "nuclear_position_score"]
    + ["g_r_color_at_peak", "g_r_color_peak_to_late", "r_i_color_at_peak", "r_i_color_peak_to_late"]
    + ["mhps_10d", "mhps_30d", "mhps_100d", "mhps_10_100_ratio", "mhps_30_100_ratio", "mhps_dominant_scale"]
    + ["luminosity_distance_mpc", "peak_luminosity", "luminosity_amplitude", "mean_luminosity", "luminosity_decline_rate"]
)

MODULE_COLUMNS = {
    "statistical": STAT_COLUMNS, "colors": COLOR_COLUMNS, "shape": SHAPE_COLUMNS,
    "tde": TDE_COLUMNS, "physics": PHYSICS_COLUMNS, "bazin": BAZIN_COLUMNS,
    "powerlaw": POWERLAW_COLUMNS, "gp2d": GP2D_COLUMNS, "gp1d": GP1D_COLUMNS,
    "research": RESEARCH_COLUMNS,       # experimental pass (mlrn_research): not in the default pipeline yet
}
ORACLE_ONLY_COLUMNS = {}
assert len(RESEARCH_COLUMNS) == 40

assert len(STAT_COLUMNS) == 123 and len(COLOR_COLUMNS) == 83 and len(SHAPE_COLUMNS) == 65
assert len(TDE_COLUMNS) == 25 and len(PHYSICS_COLUMNS) == 32 and len(BAZIN_COLUMNS) == 52
assert len(POWERLAW_COLUMNS) == 27 and len(GP2D_COLUMNS) == 27 and len(GP1D_COLUMNS) == 21
