"""Oracle for the TDE-physics features (test infrastructure; see ``oracle/__init__.py``).

Restates ``src/features/tde_physics.py:25-374``.  Row layout is ``schema.TDE_COLUMNS``
(25 values).  Every band view is time-sorted, as the reference's ``sort_values``.
"""
from __future__ import annotations

import numpy as np

from .common import NAN, ObjectView, G, R, I


def _nearest_pairs(t1, f1, t2, f2, window):
    """For each band-1 point, the nearest band-2 point (first argmin) if |dt| < window."""
    for a, fa in zip(t1, f1):
        d = np.abs(t2 - a)
        j = int(np.argmin(d))
        if d[j] < window:
            yield a, fa, f2[j]


def color_variance(v: ObjectView, b1: int, b2: int):
    """tde_physics.py:37-88 -> (var, range, trend per 100 d)."""
    t1, f1, _ = v.band_sorted(b1)
    t2, f2, _ = v.band_sorted(b2)
    if t1.shape[0] < 3 or t2.shape[0] < 3:
        return NAN, NAN, NAN
    cs, ts = [], []
    for a, fa, fb in _nearest_pairs(t1, f1, t2, f2, 5.0):
        if fa > 0 and fb > 0:
            cs.append(-2.5 * np.log10(fa / fb))
            ts.append(a)
    if len(cs) < 3:
        return NAN, NAN, NAN
    cs, ts = np.array(cs), np.array(ts)
    return np.var(cs), np.max(cs) - np.min(cs), np.polyfit(ts - ts[0], cs, 1)[0] * 100.0


def late_time(v: ObjectView, b: int):
    """tde_physics.py:103-153 -> (late_slope, late_flux_ratio, rebrightening)."""
    t, f, _ = v.band_sorted(b)
    if t.shape[0] < 5:
        return NAN, NAN, NAN
    k = int(np.argmax(f))
    pt, pf = t[k], f[k]
    m = t > pt + 50
    lt, lf = t[m], f[m]
    if lt.shape[0] < 3 or not pf > 0:
        return NAN, NAN, NAN
    x = np.log10(lt - pt + 1)
    y = np.log10(np.maximum(lf, 1e-10))
    slope = np.polyfit(x, y, 1)[0] if np.std(x) > 0 else NAN
    lm = np.mean(lf)
    return slope, lm / pf, (np.max(lf) / lm if lm > 0 else 1.0)


def rise_shape(v: ObjectView, b: int):
    """tde_physics.py:167-204 -> (rise_shape, rise_rate)."""
    t, f, _ = v.band_sorted(b)
    if t.shape[0] < 5:
        return NAN, NAN
    k = int(np.argmax(f))
    pf = f[k]
    rt, rf = t[:k + 1], f[:k + 1]
    if rt.shape[0] < 3 or not pf > 0:
        return NAN, NAN
    nf = rf / pf
    nt = (rt - rt[0]) / (rt[-1] - rt[0] + 1e-6)
    shape = np.mean(nf) / np.mean(nt) if np.mean(nt) > 0 else 1.0
    rate = pf / (rt[-1] - rt[0]) if rt[-1] > rt[0] else NAN
    return shape, rate


def temperature(v: ObjectView):
    """tde_physics.py:221-282 -> (temp_stability, temp_trend, temp_late_vs_peak)."""
    tg, fg, _ = v.band_sorted(G)
    tr, fr, _ = v.band_sorted(R)
    if tg.shape[0] < 3 or tr.shape[0] < 3:
        return NAN, NAN, NAN
    temps, ts = [], []
    for a, fa, fb in _nearest_pairs(tg, fg, tr, fr, 3.0):
        if fa > 0 and fb > 0:
            c = -2.5 * np.log10(fa / fb)
            temps.append(40000.0 if c < -0.5 else (5000.0 if c > 1.5 else 7000.0 / (c + 0.5)))
            ts.append(a)
    n = len(temps)
    if n < 3:
        return NAN, NAN, NAN
    temps, ts = np.array(temps), np.array(ts)
    stab = np.std(temps) / np.mean(temps)
    trend = np.polyfit(ts - ts[0], temps, 1)[0] * 100.0
    lvp = NAN
    if n > 4:
        lvp = np.mean(temps[-3:]) / np.mean(temps[:max(2, n // 4)])
    return stab, trend, lvp


def decay_power_law(v: ObjectView, b: int):
    """tde_physics.py:296-350 -> (alpha, residual std, alpha_late)."""
    t, f, _ = v.band_sorted(b)
    if t.shape[0] < 5:
        return NAN, NAN, NAN
    k = int(np.argmax(f))
    pt, pf = t[k], f[k]
    m = t > pt
    tp, fp = t[m], f[m]
    if tp.shape[0] < 4 or not pf > 0:
        return NAN, NAN, NAN
    dt = np.maximum(tp - pt, 1)
    ok = fp > 0
    if ok.sum() < 3:
        return NAN, NAN, NAN
    x, y = np.log10(dt[ok]), np.log10(fp[ok])
    c = np.polyfit(x, y, 1)
    resid = np.std(y - (c[0] * x + c[1]))
    late = (dt > 50) & ok
    alpha_late = np.polyfit(np.log10(dt[late]), np.log10(fp[late]), 1)[0] if late.sum() >= 3 else NAN
    return c[0], resid, alpha_late


def object_row(v: ObjectView) -> np.ndarray:
    row = np.full(25, NAN)
    row[0:3] = color_variance(v, G, R)
    row[3:6] = color_variance(v, R, I)
    for j, b in enumerate((G, R, I)):
        row[6 + 3 * j:9 + 3 * j] = late_time(v, b)
    row[15:17] = rise_shape(v, G)
    row[17:19] = rise_shape(v, R)
    row[19:22] = temperature(v)
    a, res, al = decay_power_law(v, R)
    row[22], row[23], row[24] = a, res, al
    return row
