"""Oracle for the light-curve shape features (test infrastructure; see ``oracle/__init__.py``).

Restates ``src/features/lightcurve_shape.py:177-332``.  Row layout is
``schema.SHAPE_COLUMNS`` (65 values).
"""
from __future__ import annotations

import numpy as np

from .common import NAN, ObjectView, G, R, I


def rise_time(t, f, pt, pf, frac=0.1):
    """lightcurve_shape.py:34-65 -- order-sensitive: 'first' is first in the given order."""
    if np.isnan(pt) or np.isnan(pf) or t.shape[0] < 2:
        return NAN
    pre = t < pt
    if not pre.any():
        return NAN
    tp, fp = t[pre], f[pre]
    above = fp > frac * pf
    if not above.any():
        return pt - tp[0]
    return pt - tp[int(np.argmax(above))]


def fade_time(t, f, pt, pf, frac):
    """lightcurve_shape.py:68-104 (sorts the post-peak part by time itself)."""
    if np.isnan(pt) or np.isnan(pf) or t.shape[0] < 2:
        return NAN
    post = t > pt
    if not post.any():
        return NAN
    tp, fp = t[post], f[post]
    o = np.argsort(tp, kind="stable")
    tp, fp = tp[o], fp[o]
    below = fp < frac * pf
    if not below.any():
        return tp[-1] - pt
    return tp[int(np.argmax(below))] - pt


def loglog_decay(t, f, pt, pf):
    """lightcurve_shape.py:107-144 -> (alpha, rms residual)."""
    if np.isnan(pt) or np.isnan(pf):
        return NAN, NAN
    m = (t > pt + 5) & (f > 0)
    if m.sum() < 5:
        return NAN, NAN
    x = np.log10(np.maximum(t[m] - pt, 1.0))
    y = np.log10(np.maximum(f[m], 1e-10))
    try:
        c = np.polyfit(x, y, 1)
    except Exception:
        return NAN, NAN
    return c[0], np.sqrt(np.mean((y - (c[0] * x + c[1])) ** 2))


def duration_above(t, f, frac):
    """lightcurve_shape.py:147-161."""
    if t.shape[0] < 2:
        return NAN
    above = f > frac * np.max(f)
    if not above.any():
        return 0.0
    ta = t[above]
    return np.max(ta) - np.min(ta)


def object_row(v: ObjectView) -> np.ndarray:
    row = np.full(65, NAN)
    peak_t = {}
    for b in range(6):                                   # :204-247
        t, f, _ = v.band_sorted(b)
        if t.shape[0] < 3:                               # :193
            continue
        k = int(np.argmax(f))
        pt, pf = t[k], f[k]
        peak_t[b] = pt
        o = 8 * b
        row[o] = rise_time(t, f, pt, pf)
        row[o + 1] = fade_time(t, f, pt, pf, 0.5)
        row[o + 2] = fade_time(t, f, pt, pf, 0.25)
        if not np.isnan(row[o]) and not np.isnan(row[o + 1]) and row[o + 1] > 0:
            row[o + 3] = row[o] / row[o + 1]
        row[o + 4] = duration_above(t, f, 0.5)
        row[o + 5] = duration_above(t, f, 0.25)
        row[o + 6], row[o + 7] = loglog_decay(t, f, pt, pf)
    k = 48
    pts = [x for x in peak_t.values() if not np.isnan(x)]    # :252-258
    if len(pts) >= 2:
        row[k] = np.max(pts) - np.min(pts)
        row[k + 1] = np.std(pts)
    rises = [row[8 * b] for b in (G, R, I) if not np.isnan(row[8 * b])]          # :261-284
    fades = [row[8 * b + 1] for b in (G, R, I) if not np.isnan(row[8 * b + 1])]
    alphas = [row[8 * b + 6] for b in (G, R, I) if not np.isnan(row[8 * b + 6])]
    row[k + 2] = np.mean(rises) if rises else NAN
    row[k + 3] = np.mean(fades) if fades else NAN
    row[k + 4] = np.mean(alphas) if alphas else NAN
    if len(rises) >= 2:
        row[k + 5] = np.std(rises) / (np.mean(rises) + 1e-6)
    if len(fades) >= 2:
        row[k + 6] = np.std(fades) / (np.mean(fades) + 1e-6)
    t, f = v.t, v.f                                      # file order, every row (:287-288)
    if t.shape[0] >= 5:                                  # :290-319
        j = int(np.argmax(f))
        pt, pf = t[j], f[j]
        row[k + 7] = rise_time(t, f, pt, pf)
        row[k + 8] = fade_time(t, f, pt, pf, 0.5)
        if not np.isnan(row[k + 7]) and not np.isnan(row[k + 8]) and row[k + 8] > 0:
            row[k + 9] = row[k + 7] / row[k + 8]
        row[k + 10], row[k + 11] = loglog_decay(t, f, pt, pf)
        row[k + 12:k + 16] = [np.percentile(f, q) for q in (10, 25, 75, 90)]
        if pf > 0:
            row[k + 16] = pf / (np.sum(f) + 1e-6)
    return row
