"""Acceptance rule for the iterative fits (Bazin, decline models), shared by the GPU tests.

North star: fitted parameters within rtol 1e-3 wherever the reference converged; elsewhere the final
chi^2 no worse than the reference's (+1e-6 relative).  The reference's optimiser differentiates by
forward differences with h ~ 1.5e-8 |x|, so a 1-ulp difference in exp()/pow() between numpy and any
other libm (CUDA's included) perturbs its Jacobian at the 1e-8 level; along the flat valleys typical of
sparse light curves the *reference's own* termination point moves by percents under such a perturbation.
Each fit is therefore classified as
  match     every compared value within rtol 1e-3 (+ atol)
  no_worse  values differ but our chi^2 (or 1 - R^2) <= reference * (1 + 1e-6)
  worse     a different local minimum with a higher chi^2
  nan_diff  one side NaN only (the reference raised / hit max_nfev, or we did)
and the tests bound the fraction in each class (tests/test_gpu_fits.py).
"""
import numpy as np


def classify_bazin(got, ref):
    """got, ref: [n, 52].  Returns dict of counts over (object, band) fits the reference attempted."""
    c = {"fits": 0, "match": 0, "no_worse": 0, "worse": 0, "nan_diff": 0, "worst": 0.0}
    for i in range(ref.shape[0]):
        for b in range(6):
            r, m = ref[i, 8 * b:8 * b + 8], got[i, 8 * b:8 * b + 8]
            if np.isnan(r[0]) and np.isnan(m[0]):
                continue
            c["fits"] += 1
            if np.isnan(r[0]) != np.isnan(m[0]):
                c["nan_diff"] += 1
            elif np.allclose(m[:5], r[:5], rtol=1e-3, atol=1e-9):
                c["match"] += 1
            elif m[5] <= r[5] * (1 + 1e-6):
                c["no_worse"] += 1
            else:
                c["worse"] += 1
                c["worst"] = max(c["worst"], m[5] / r[5] - 1)
    return c


def classify_r2(got, ref):
    """got, ref: [n, 27] R^2 tables."""
    c = {"fits": 0, "match": 0, "no_worse": 0, "worse": 0, "nan_diff": 0, "worst": 0.0}
    both_nan = np.isnan(got) & np.isnan(ref)
    one_nan = np.isnan(got) ^ np.isnan(ref)
    ok = ~both_nan & ~one_nan
    c["fits"] = int((~both_nan).sum())
    c["nan_diff"] = int(one_nan.sum())
    d = np.where(ok, got - ref, 0.0)
    match = ok & (np.abs(d) <= 1e-3 * np.abs(np.where(ok, ref, 0.0)) + 1e-6)
    no_worse = ok & ~match & (d >= -1e-6)
    worse = ok & ~match & ~no_worse
    c["match"], c["no_worse"], c["worse"] = int(match.sum()), int(no_worse.sum()), int(worse.sum())
    c["worst"] = float(-d[worse].min()) if worse.any() else 0.0
    return c


def classify_gp1d(got, ref):
    """got, ref: [n, 21] per-band GP tables.  One fit per (object, band in g r i z): {length scale, amplitude, noise,
    log-likelihood}.  ``match``: all four within rtol 1e-3 (atol 1e-6); ``no_worse``: hyper-parameters differ (a flat
    direction -- typically the white-noise level when the measurement variances dominate) but our log-likelihood
    is at least the reference's minus 1e-6 relative; ``worse`` otherwise.  ``cross_bad`` counts cross-band
    columns (ratios / means, 16..20) that disagree although every band they are built from matched."""
    c = {"fits": 0, "match": 0, "no_worse": 0, "worse": 0, "nan_diff": 0, "worst": 0.0, "cross_bad": 0}
    for i in range(ref.shape[0]):
        all_match = True
        for b in range(4):
            r, m = ref[i, 4 * b:4 * b + 4], got[i, 4 * b:4 * b + 4]
            if np.isnan(r).all() and np.isnan(m).all():
                continue
            c["fits"] += 1
            if not np.array_equal(np.isnan(r), np.isnan(m)):
                c["nan_diff"] += 1; all_match = False
            elif np.allclose(m, r, rtol=1e-3, atol=1e-6):
                c["match"] += 1
            elif m[3] >= r[3] - 1e-6 * abs(r[3]) - 1e-9:
                c["no_worse"] += 1; all_match = False
            else:
                c["worse"] += 1; all_match = False
                c["worst"] = max(c["worst"], float(r[3] - m[3]))
        if all_match:
            r, m = ref[i, 16:21], got[i, 16:21]
            if not (np.array_equal(np.isnan(r), np.isnan(m)) and np.allclose(m[~np.isnan(r)], r[~np.isnan(r)], rtol=2e-3, atol=1e-6)):
                c["cross_bad"] += 1
    return c
