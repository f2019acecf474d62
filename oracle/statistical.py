"""Oracle for the per-band moment pass (test infrastructure; see ``oracle/__init__.py``).

Restates ``src/features/statistical.py``: ``compute_band_statistics`` (:41-132) and the
cross-band tail of ``extract_statistical_features`` (:185-222).  Output row layout is
``schema.STAT_COLUMNS`` (123 values; ``n_obs`` / ``peak_band`` are integer-valued).
"""
from __future__ import annotations

import numpy as np

from .common import NAN, ObjectView, G, R, I, Z

N_STAT = 17


def band_stats(f: np.ndarray, e: np.ndarray, t: np.ndarray) -> np.ndarray:
    """17 statistics of one segment (statistical.py:41-132)."""
    out = np.full(N_STAT, NAN)
    n = f.shape[0]
    out[0] = n
    if n == 0:                                      # :56-66
        return out
    mean = np.mean(f)
    std = np.std(f) if n > 1 else 0.0               # :71 (ddof=0)
    out[1], out[2], out[3], out[4] = mean, std, np.min(f), np.max(f)
    med = np.median(f)
    out[5] = med
    skew = kurt = 0.0                               # :77-82 with the gates of :14-35
    if n > 2 and std != 0:
        zs = (f - mean) / std
        skew = np.mean(zs ** 3)
        if n >= 4:
            kurt = np.mean(zs ** 4) - 3.0
    out[6], out[7] = skew, kurt
    out[8] = out[4] - out[3]
    out[9] = np.median(np.abs(f - med))
    out[10] = (np.percentile(f, 75) - np.percentile(f, 25)) if n > 1 else 0.0
    if std > 0:                                     # :90-96
        zabs = np.abs(f - mean) / std
        out[11], out[12] = np.mean(zabs > 1), np.mean(zabs > 2)
    else:
        out[11] = out[12] = 0.0
    ms = 0.0                                        # :99-113
    if n > 1:
        o = np.argsort(t, kind="stable")
        dt, df = np.diff(t[o]), np.diff(f[o])
        ok = dt > 0
        if ok.any():
            ms = np.max(np.abs(df[ok] / dt[ok]))
    out[13] = ms
    pos = e > 0                                     # :116-120
    out[14] = np.mean(np.abs(f[pos]) / e[pos]) if pos.any() else NAN
    if n > 1:                                       # :123-130
        out[15] = np.max(t) - np.min(t)
        out[16] = np.mean(np.diff(np.sort(t)))
    else:
        out[15] = out[16] = 0.0
    return out


def object_row(v: ObjectView) -> np.ndarray:
    """123-value row of ``extract_statistical_features`` for one object (:166-222)."""
    row = np.empty(7 * N_STAT + 4)
    for b in range(6):
        m = v.b == b                                # file order, :170
        row[b * N_STAT:(b + 1) * N_STAT] = band_stats(v.f[m], v.e[m], v.t[m])
    row[6 * N_STAT:7 * N_STAT] = band_stats(v.f, v.e, v.t)      # every row incl. unknown filters, :186-190
    mean = row[1:6 * N_STAT:N_STAT]
    mx = row[4:6 * N_STAT:N_STAT]
    k = 7 * N_STAT
    for j, (a, c) in enumerate(((G, R), (R, I), (I, Z))):       # :201-214
        row[k + j] = mean[a] / mean[c] if (not np.isnan(mean[a]) and mean[c] > 0) else NAN
    valid = ~np.isnan(mx)                                       # :217-222 (first band with the largest max)
    row[k + 3] = float(np.argmax(np.where(valid, mx, -np.inf))) if valid.any() else -1.0
    return row


def segment_meta(v: ObjectView) -> np.ndarray:
    """Integer K1 outputs for one object: for each of the 7 segments (ugrizy, unknown)
    ``[count, first-argmax position in file order, first-argmax position in time order]``
    (positions are within the band subsequence; -1 when empty), then ``peak_band``.

    These are the indices every reference extractor derives with ``np.argmax`` on the
    band's file-order values (colors.py:39-44) or time-sorted values
    (lightcurve_shape.py:30, bazin_fitting.py:97, tde_physics.py:116).
    """
    out = np.full(7 * 3 + 1, -1, dtype=np.int64)
    best, best_b = -np.inf, -1
    for b in range(7):
        m = v.b == b
        n = int(m.sum())
        out[3 * b] = n
        if n:
            fb, tb = v.f[m], v.t[m]
            out[3 * b + 1] = int(np.argmax(fb))
            o = np.argsort(tb, kind="stable")
            out[3 * b + 2] = int(np.argmax(fb[o]))
            if b < 6 and fb.max() > best:
                best, best_b = fb.max(), b
    out[21] = best_b
    return out
