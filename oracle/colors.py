"""Oracle for the colour-evolution features (test infrastructure; see ``oracle/__init__.py``).

Restates ``src/features/colors.py:108-344`` on flat arrays.  Row layout is
``schema.COLOR_COLUMNS`` (83 values, ``peak_mjd`` first).
"""
from __future__ import annotations

import numpy as np

from .common import NAN, ObjectView, mag_color, U, G, R, I, Z

PAIRS = ((G, R), (R, I), (U, G), (I, Z))
EPOCHS = (0.0, 10.0, 20.0, 30.0, 50.0, 75.0, 100.0, 150.0, -10.0, -20.0)
E_PEAK, E_P30, E_P50, E_P75, E_P100, E_P150 = 0, 3, 4, 5, 6, 7


def lerp_flux(ts: np.ndarray, fs: np.ndarray, target: float, max_gap: float) -> float:
    """Linear interpolation on a time-sorted band (colors.py:47-89)."""
    n = ts.shape[0]
    if n < 2:
        return NAN
    if target < ts[0] or target > ts[-1]:
        return NAN
    k = int(np.searchsorted(ts, target))           # left; NaN target lands at n
    if k == 0:
        return fs[0]
    if k == n:
        return fs[-1]
    t1, t2 = ts[k - 1], ts[k]
    if t2 - t1 > max_gap:
        return NAN
    w = (target - t1) / (t2 - t1)
    return fs[k - 1] + w * (fs[k] - fs[k - 1])


def temp_from_gr(c: float) -> float:
    """colors.py:297-308."""
    if np.isnan(c):
        return NAN
    if c < -0.5:
        return 50000.0
    if c > 2.0:
        return 3000.0
    return 7000.0 / (c + 0.6)


def object_row(v: ObjectView) -> np.ndarray:
    row = np.full(83, NAN)
    file_band = [v.band_file(b) for b in range(6)]        # file order (colors.py:123)
    sorted_band = [v.band_sorted(b) for b in range(6)]    # interpolate_flux sorts internally
    has = [fb[0].shape[0] > 0 for fb in file_band]

    peak_t = {}
    for b in (R, G, I):                                   # :133-138, first argmax in file order
        if has[b]:
            tb, fb, _ = file_band[b]
            peak_t[b] = tb[int(np.argmax(fb))]
    ref = NAN                                             # :141-148
    for b in (R, G, I):
        if b in peak_t and not np.isnan(peak_t[b]):
            ref = peak_t[b]
            break
    row[0] = ref

    col = np.full((len(EPOCHS), 4), NAN)
    for ei, d in enumerate(EPOCHS):                       # :168-189
        target = ref + d if not np.isnan(ref) else NAN
        fl = [lerp_flux(sorted_band[b][0], sorted_band[b][1], target, 50.0) if has[b] else NAN
              for b in range(6)]
        for pi, (a, c) in enumerate(PAIRS):
            col[ei, pi] = mag_color(fl[a], fl[c])
    row[1:41] = col.reshape(-1)

    k = 41
    for pi in range(4):                                   # :192-207
        c0, c50, c100 = col[E_PEAK, pi], col[E_P50, pi], col[E_P100, pi]
        row[k + 2 * pi] = (c50 - c0) / 50.0 if not (np.isnan(c0) or np.isnan(c50)) else NAN
        row[k + 2 * pi + 1] = (c100 - c0) / 100.0 if not (np.isnan(c0) or np.isnan(c100)) else NAN
    k = 49
    for pi, (a, c) in enumerate(PAIRS):                   # :211-232
        if has[a] and has[c]:
            ts2, fs2, _ = sorted_band[c]
            cs = []
            for t1, f1 in zip(file_band[a][0], file_band[a][1]):
                cc = mag_color(f1, lerp_flux(ts2, fs2, t1, 5.0))
                if not np.isnan(cc):
                    cs.append(cc)
            if len(cs) >= 3:
                row[k + 2 * pi] = np.std(cs)
                row[k + 2 * pi + 1] = np.max(cs) - np.min(cs)
    k = 57
    pk = np.full(6, NAN)
    for b in range(6):                                    # :235-239
        if has[b]:
            pk[b] = np.max(file_band[b][1])
    row[k:k + 6] = pk
    k = 63
    for pi, (a, c) in enumerate(PAIRS):                   # :242-248
        if not np.isnan(pk[a]) and not np.isnan(pk[c]) and pk[c] > 0:
            row[k + pi] = pk[a] / pk[c]
    k = 67
    for j, (a, c) in enumerate(((G, R), (R, I))):         # :252-257
        if a in peak_t and c in peak_t:
            row[k + j] = peak_t[a] - peak_t[c]
    k = 69
    for j in range(2):                                    # :263-275 (pairs g_r, r_i are PAIRS[0:2])
        c0, c30, c75 = col[E_PEAK, j], col[E_P30, j], col[E_P75, j]
        if not (np.isnan(c0) or np.isnan(c30) or np.isnan(c75)):
            row[k + j] = ((c75 - c30) / 45.0 - (c30 - c0) / 30.0) / 37.5
    k = 71
    for j in range(2):                                    # :279-293
        late = [col[ei, j] for ei in (E_P50, E_P75, E_P100, E_P150)]
        ok = [x for x in late if not np.isnan(x)]
        if len(ok) >= 2:
            row[k + 2 * j] = np.std(ok)
            row[k + 2 * j + 1] = np.mean(ok)
    k = 75
    temps = [temp_from_gr(col[ei, 0]) for ei in (E_PEAK, E_P30, E_P75, E_P150)]   # :310-313
    row[k:k + 4] = temps
    tp, t30, t75, t150 = temps
    if not (np.isnan(tp) or np.isnan(t30)):               # :321-334
        row[k + 4] = (t30 - tp) / 30.0
    if not (np.isnan(t30) or np.isnan(t75)):
        row[k + 5] = (t75 - t30) / 45.0
    if not (np.isnan(t75) or np.isnan(t150)):
        row[k + 6] = (t150 - t75) / 75.0
    ok = [x for x in temps if not np.isnan(x)]            # :337-342
    if len(ok) >= 2:
        row[k + 7] = np.std(ok) / np.mean(ok)
    return row
