"""Oracle for the physics-based features (test infrastructure; see ``oracle/__init__.py``).

Restates ``src/features/physics_based.py:31-458``.  Row layout is
``schema.PHYSICS_COLUMNS`` (32 values).  Band views are time-sorted and only bands with
at least 3 observations take part (:307-314).

Deliberate deviation: when the r band has 3 or 4 points the reference omits the
``r_sf_slope`` key from that object's dict (:128-129), which only perturbs the column
*position* pandas assigns; here the value is simply NaN in its usual place.
"""
from __future__ import annotations

import numpy as np

from .common import NAN, ObjectView, G, R, I

SF_TAUS = (1.0, 5.0, 10.0, 30.0, 100.0)


def stetson_j(t1, f1, e1, t2, f2, e2, max_dt=0.5):
    """physics_based.py:49-82."""
    if t1.shape[0] < 3 or t2.shape[0] < 3:
        return NAN
    m1, s1 = np.mean(f1), np.std(f1)
    m2, s2 = np.mean(f2), np.std(f2)
    if s1 == 0 or s2 == 0:
        return 0.0
    acc, npair = 0.0, 0
    for a, fa, ea in zip(t1, f1, e1):
        d = np.abs(t2 - a)
        j = int(np.argmin(d))
        if d[j] <= max_dt and ea > 0 and e2[j] > 0:
            p = ((fa - m1) / ea) * ((f2[j] - m2) / e2[j])
            acc += np.sign(p) * np.sqrt(np.abs(p))
            npair += 1
    return acc / npair if npair else NAN


def stetson_k(f, e):
    """physics_based.py:92-107 (sqrt(n) uses the full length, sums the err>0 subset)."""
    n = f.shape[0]
    if n < 4:
        return NAN
    ok = e > 0
    if ok.sum() < 4:
        return NAN
    d = np.abs(f[ok] - np.mean(f)) / e[ok]
    return np.sum(d) / np.sqrt(np.sum(d ** 2)) / np.sqrt(n)


def structure_function(t, f):
    """physics_based.py:128-168 -> 5 SF values + log-log slope."""
    out = np.full(6, NAN)
    n = t.shape[0]
    if n < 5:
        return out
    iu, ju = np.triu_indices(n, k=1)
    dt = np.abs(t[ju] - t[iu])
    d2 = (f[ju] - f[iu]) ** 2
    for k, tau in enumerate(SF_TAUS):
        sel = d2[(dt >= 0.5 * tau) & (dt <= 1.5 * tau)]
        if sel.shape[0] >= 3:
            out[k] = np.sqrt(np.mean(sel))
    xs = [np.log10(tau) for tau, s in zip(SF_TAUS, out[:5]) if not np.isnan(s) and s > 0]
    ys = [np.log10(s) for s in out[:5] if not np.isnan(s) and s > 0]
    if len(xs) >= 3:
        out[5] = np.polyfit(xs, ys, 1)[0]
    return out


def temperature(g, r, i):
    """physics_based.py:179-199."""
    for x in (g, r, i):
        if np.isnan(x) or x <= 0:
            return NAN
    c = -2.5 * np.log10(g / r)
    temp = 50000.0 if c < -0.5 else (3000.0 if c > 2.0 else 7000.0 / (c + 0.6))
    return float(np.clip(temp, 3000, 100000))


def bazin_heuristics(t, f):
    """physics_based.py:210-289 -> amplitude, t0, rise_approx, fall_approx, plateau."""
    out = np.full(5, NAN)
    n = t.shape[0]
    if n < 5:
        return out
    k = int(np.argmax(f))
    pt, pf = t[k], f[k]
    out[0], out[1] = pf, pt
    if k + 1 >= 2:                                      # :235-252
        t10, t90 = t[0], pt
        for a, fa in zip(t[:k + 1], f[:k + 1]):
            if fa >= 0.1 * pf and t10 == t[0]:
                t10 = a
            if fa >= 0.9 * pf:
                t90 = a
                break
        out[2] = t90 - t10
    tp, fp = t[k:], f[k:]
    if tp.shape[0] >= 3:                                # :258-272
        target = pf / np.e
        fall = NAN
        for a, fa in zip(tp, fp):
            if fa <= target:
                fall = a - pt
                break
        if np.isnan(fall):
            fall = (tp[-1] - pt) * pf / (pf - fp[-1] + 1e-6)
        out[3] = fall
    if fp.shape[0] >= 5:                                # :277-285
        mid = fp.shape[0] // 2
        early, late = np.mean(fp[:mid]), np.mean(fp[mid:])
        if early > 0:
            out[4] = late / early
    return out


def object_row(v: ObjectView, redshift: float) -> np.ndarray:
    row = np.full(32, NAN)
    bd = {}
    for b in (G, R, I):
        t, f, e = v.band_sorted(b)
        if t.shape[0] >= 3:
            bd[b] = (t, f, e)
    for j, (a, c) in enumerate(((G, R), (R, I), (G, I))):       # :318-326
        if a in bd and c in bd:
            row[j] = stetson_j(*bd[a], *bd[c])
    for j, b in enumerate((G, R, I)):                           # :329-334
        if b in bd:
            row[3 + j] = stetson_k(bd[b][1], bd[b][2])
    if R in bd:                                                 # :338-345
        row[6:12] = structure_function(bd[R][0], bd[R][1])
    z = redshift if not np.isnan(redshift) else 0.0             # :348
    for j, b in enumerate((G, R, I)):                           # :351-379
        if b in bd:
            t, f, _ = bd[b]
            k = int(np.argmax(f))
            row[12 + 3 * j] = (t[-1] - t[0]) / (1 + z)
            row[13 + 3 * j] = (t[k] - t[0]) / (1 + z) if k > 0 else NAN
            row[14 + 3 * j] = (t[-1] - t[k]) / (1 + z) if k < t.shape[0] - 1 else NAN
    if G in bd and R in bd and I in bd:                         # :383-423
        row[21] = temperature(np.max(bd[G][1]), np.max(bd[R][1]), np.max(bd[I][1]))
        target = bd[R][0][int(np.argmax(bd[R][1]))] + 50
        late = []
        for b in (G, R, I):
            d = np.abs(bd[b][0] - target)
            j = int(np.argmin(d))
            late.append(bd[b][1][j] if d[j] < 20 else NAN)
        row[22] = temperature(*late)
        if not np.isnan(row[21]) and not np.isnan(row[22]):
            row[23] = (row[22] - row[21]) / 50.0
    if R in bd:                                                 # :427-433
        row[24:29] = bazin_heuristics(bd[R][0], bd[R][1])
    ok = (v.e > 0) & (v.f > 0)                                  # :437-456, every row in file order
    if ok.sum() > 0:
        fo, eo = v.f[ok], v.e[ok]
        snr = fo / eo
        row[29], row[30] = np.mean(snr), np.median(snr)
        ev = (np.var(fo) - np.mean(eo ** 2)) / np.mean(fo) ** 2
        row[31] = max(0, ev)
    return row
