"""Oracle for the v55 decline-model fits (test infrastructure; see ``oracle/__init__.py``).

Restates ``scripts/train_v55_powerlaw.py:108-202`` (twin: ``scripts/visualize_and_powerlaw.py:148-255``):
nine post-peak decline models per band g, r, i fitted with bounded, unweighted
``scipy.optimize.curve_fit`` (TRF, ``maxfev=1000``); the feature is R^2.
Row layout is ``schema.POWERLAW_COLUMNS`` (27 values).
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import curve_fit

from .common import NAN, ObjectView, G, R, I

POWER_INDICES = (-5 / 3, -1, -1.5, -2, -2.5, -3, -0.5)          # :108-127, in MODELS order


def _power(k):
    return lambda t, A, t0: A * np.power(np.maximum(t - t0, 0.1), k)


def _expo(t, A, tau, t0):
    return A * np.exp(-np.maximum(t - t0, 0) / tau)


def _line(t, A, b, t0):
    return A - b * np.maximum(t - t0, 0)


MODELS = [(_power(k), 2, None) for k in POWER_INDICES] + [(_expo, 3, "exp"), (_line, 3, "lin")]


def fit_band(t, f, want_state=False):
    """9 R^2 values for one time-sorted band (train_v55_powerlaw.py:147-194)."""
    out = np.full(9, NAN)
    states = [None] * 9
    if t.shape[0] < 5:
        return (out, states) if want_state else out
    k = int(np.argmax(f))
    pt, pf = t[k], f[k]
    post = t > pt
    if post.sum() < 3:
        return (out, states) if want_state else out
    x, y = t[post] - pt, f[post]
    for m, (fn, npar, kind) in enumerate(MODELS):
        try:
            if npar == 2:
                p0, bnd = [pf, 0], ([0, -10], [1e6, 10])
            elif kind == "exp":
                p0, bnd = [pf, 30, 0], ([0, 1, -10], [1e6, 500, 10])
            else:
                p0, bnd = [pf, 1, 0], ([0, 0, -10], [1e6, 100, 10])
            if want_state:
                popt, _, info, _, ier = curve_fit(fn, x, y, p0=p0, maxfev=1000, bounds=bnd, full_output=True)
                states[m] = {"status": ier, "nfev": info["nfev"], "x": popt}
            else:
                popt, _ = curve_fit(fn, x, y, p0=p0, maxfev=1000, bounds=bnd)
            ss_res = np.sum((y - fn(x, *popt)) ** 2)
            ss_tot = np.sum((y - np.mean(y)) ** 2)
            out[m] = 1 - ss_res / ss_tot if ss_tot > 0 else 0
            if want_state:
                states[m]["ss_res"] = ss_res
        except Exception:
            out[m] = NAN
    return (out, states) if want_state else out


def object_row(v: ObjectView) -> np.ndarray:
    row = np.full(27, NAN)
    for j, b in enumerate((G, R, I)):
        t, f, _ = v.band_sorted(b)
        row[9 * j:9 * j + 9] = fit_band(t, f)
    return row
