"""Oracle for the Bazin parametric fits (test infrastructure; see ``oracle/__init__.py``).

Restates ``src/features/bazin_fitting.py:63-251``; the optimiser is the very call the
reference makes -- ``scipy.optimize.curve_fit`` with bounds (=> ``least_squares(method='trf')``,
2-point finite-difference Jacobian, ftol=xtol=gtol=1e-8, ``max_nfev=2000``).
Row layout is ``schema.BAZIN_COLUMNS`` (52 values).  ``fit_band`` can also return the raw
optimiser state (status, cost, nfev) for the convergence-aware parity rule.
"""
from __future__ import annotations

import warnings

import numpy as np
from scipy.optimize import OptimizeWarning, curve_fit

from .common import NAN, ObjectView, G, R, I


def bazin(t, A, t0, tau_rise, tau_fall, B):
    """bazin_fitting.py:37-60."""
    return A * np.exp(-(t - t0) / tau_fall) / (1.0 + np.exp(-(t - t0) / tau_rise)) + B


def start_and_bounds(t, f):
    """p0 and box of bazin_fitting.py:97-118 for a time-sorted band."""
    k = int(np.argmax(f))
    med = np.median(f)
    dur = t[-1] - t[0]
    mx = np.max(f)
    p0 = [f[k] - med, t[k], dur * 0.2, dur * 0.3, med]
    lo = [0, t[0], 0.1, 0.1, -mx]
    hi = [3 * mx, t[-1], dur, dur, 2 * mx]
    return p0, lo, hi


def fit_band(t, f, e, want_state=False):
    """8 features of one band (bazin_fitting.py:63-179); ``t`` already time-sorted."""
    out = np.full(8, NAN)
    state = {"status": None, "chi2": NAN}
    n = t.shape[0]
    if n < 5:
        return (out, state) if want_state else out
    p0, lo, hi = start_and_bounds(t, f)
    sigma = np.where(e > 0, e, 1.0)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", OptimizeWarning)
            warnings.simplefilter("ignore", RuntimeWarning)
            if want_state:
                popt, _, info, _, ier = curve_fit(bazin, t, f, p0=p0, bounds=(lo, hi), sigma=sigma,
                                                  absolute_sigma=True, maxfev=2000, full_output=True)
                state["status"], state["nfev"] = ier, info["nfev"]
            else:
                popt, _ = curve_fit(bazin, t, f, p0=p0, bounds=(lo, hi), sigma=sigma,
                                    absolute_sigma=True, maxfev=2000)
            A, t0, tr, tf, B = popt
            A = np.clip(A, -1e6, 1e6)
            tr = np.clip(tr, 0.1, 1e4)
            tf = np.clip(tf, 0.1, 1e4)
            B = np.clip(B, -1e6, 1e6)
            chi2 = np.sum(((f - bazin(t, A, t0, tr, tf, B)) / sigma) ** 2)
            state["chi2"] = chi2
            out[:] = [A, t0, tr, tf, B, np.clip(chi2 / (n - 5), 0, 1e6),
                      np.clip(tr / (tf + 1e-6), 0, 100), np.clip(A + B, -1e6, 1e6)]
    except (RuntimeError, ValueError, OptimizeWarning):
        out[:] = NAN
        state["status"] = -100
    return (out, state) if want_state else out


def cross_band(row: np.ndarray) -> None:
    """bazin_fitting.py:217-249, in place on a 52-value row."""
    rise = [row[8 * b + 2] for b in (G, R, I) if not np.isnan(row[8 * b + 2])]
    fall = [row[8 * b + 3] for b in (G, R, I) if not np.isnan(row[8 * b + 3])]
    row[48] = np.std(rise) / np.mean(rise) if len(rise) >= 2 else NAN
    row[49] = np.std(fall) / np.mean(fall) if len(fall) >= 2 else NAN
    chi = [row[8 * b + 5] for b in range(6) if not np.isnan(row[8 * b + 5])]
    row[50] = np.mean(chi) if chi else NAN
    row[51] = np.std(chi) if chi else NAN


def object_row(v: ObjectView) -> np.ndarray:
    row = np.full(52, NAN)
    for b in range(6):
        t, f, e = v.band_sorted(b)
        row[8 * b:8 * b + 8] = fit_band(t, f, e)
    cross_band(row)
    return row
