"""The committed downstream-acceptance record (profiles/r02_downstream_f1.json, written on a B200 by
tests/gpu_downstream_f1.py) and what it supports.

North star: OOF F1 on the GPU features within 0.005 of the CPU features.  Measured on the C3 set (3,054 labelled objects,
222-column v92d assembly, 10 seeds): that bound is NOT met per seed -- and it is not met by the reference against ITSELF
either: swapping numpy's exp for glibc's (<= 1 ulp) in the reference's own Bazin model moves F1 by 0.010 on average and by
up to 0.019 (0.010 / 0.031 with the K10 trainer).  The assertions below hold the record to exactly that statement: the GPU arm must stay within a small multiple
of the reference's own noise floor and show no significant systematic shift; they do not pretend the 0.005 bound holds."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


import pytest


@pytest.mark.parametrize("name", ["r02_downstream_f1.json", "r02_downstream_f1_k10.json"])
def test_downstream_record_is_complete_and_bounded_by_the_reference_noise_floor(name):
    """Two classifiers over the same three feature arms: scikit-learn's histogram GBDT (the stand-in for xgboost), and the
    product's own K10 trainer with the v92d parameters (row / column sub-sampling, L1 / L2, scale_pos_weight)."""
    rec = json.load(open(os.path.join(ROOT, "profiles", name)))
    assert rec["objects"] == 3054 and rec["columns"] == 222 and len(rec["seeds"]) >= 5
    assert len(rec["feature_names"]) == 222 and "all_rise_time" not in rec["feature_names"]
    d = np.array([v["delta"] for v in rec["seeds"].values()])
    assert np.isclose(rec["mean_abs_delta"], np.abs(d).mean()) and np.isclose(rec["mean_delta"], d.mean())
    floor = rec["noise_floor"]
    # the honest statement of the result
    assert rec["meets_north_star_per_seed"] == bool(np.abs(d).max() <= 0.005)
    # no implementation with another exp() can beat the reference's own reproducibility; stay within 2x of it
    assert rec["mean_abs_delta"] <= 2.0 * floor["mean_abs_delta"], (rec["mean_abs_delta"], floor["mean_abs_delta"])
    assert rec["max_abs_delta"] <= 2.0 * floor["max_abs_delta"]
    # no significant systematic shift: the mean signed difference is within two standard errors of zero and far inside
    # the seed-to-seed spread of the CPU arm itself
    assert abs(rec["mean_delta"]) <= 2.0 * rec["mean_delta_standard_error"] + 1e-12
    assert abs(rec["mean_delta"]) <= rec["f1_cpu_seed_std"]
