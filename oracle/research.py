"""Oracle for the "research" feature block (v115c).  TEST INFRASTRUCTURE ONLY.

Restates ``src/features/research_features.py:44-560`` on flat arrays: log-log power-law decay quality per band,
nuclear-position proxies, colour at peak, Mexican-hat power spectra of the r band on a 1-day grid, luminosity.
numpy / scipy calls are the reference's own (``np.polyfit``, ``np.percentile``, ``np.interp``,
``scipy.signal.convolve``).  Pinned: ``tests/golden/make_golden.py`` stores the reference's frames as
``research_rows`` and ``tests/test_oracle_golden.py`` holds this file to them.

Row layout is ``schema.RESEARCH_COLUMNS`` (40 values).
"""
from __future__ import annotations

import numpy as np
from scipy.signal import convolve

from .common import NAN, ObjectView, G, R, I

H0, C_KMS, OMEGA_M, OMEGA_L = 70.0, 299792.458, 0.3, 0.7       # research_features.py:27-31
MHPS_SCALES = (10, 30, 100)


def fit_power_law_decay(t, f, e):
    """research_features.py:44-117 -> [alpha, dev_53, dev_512, chi2, residual_std, fit_success]."""
    out = [NAN, NAN, NAN, NAN, NAN, 0.0]
    if len(t) < 5:
        return out
    k = int(np.argmax(f))
    post = (t > t[k] + 10) & (f > 0)
    pt, pf, pe = t[post], f[post], e[post]
    if len(pt) < 4:
        return out
    log_t, log_f = np.log10(pt - t[k]), np.log10(pf)
    try:
        coeffs, _ = np.polyfit(log_t, log_f, 1, cov=True)
        alpha = coeffs[0]
        out[0], out[1], out[2] = alpha, abs(alpha - (-5 / 3)), abs(alpha - (-5 / 12))
        res = log_f - (coeffs[0] * log_t + coeffs[1])
        out[4] = np.std(res)
        log_err = np.clip(pe / (pf * np.log(10) + 1e-10), 0.01, 1.0)
        out[3] = np.sum((res / log_err) ** 2) / max(len(res) - 2, 1)
        out[5] = 1.0
    except Exception:                      # the reference swallows everything here (:114)
        pass
    return out


def power_law_block(v: ObjectView):
    """compute_power_law_features (:120-156): 3 bands x 6 + 3 combined."""
    row, alphas = [], []
    for b in (G, R, I):
        t, f, e = v.band_sorted(b)
        r = [NAN] * 6 if len(t) < 5 else fit_power_law_decay(t, f, e)
        row += r
        if not np.isnan(r[0]):
            alphas.append(r[0])
    if len(alphas) >= 2:
        row += [np.mean(alphas), np.std(alphas), np.mean([abs(a + 5 / 3) for a in alphas])]
    else:
        row += [alphas[0] if alphas else NAN, NAN, abs(alphas[0] + 5 / 3) if alphas else NAN]
    return row


def nuclear_block(v: ObjectView):
    """compute_nuclear_position_proxy (:163-247)."""
    t, f, e = v.band_sorted(R)
    smooth = conc = ratio = score = NAN
    if len(t) < 10:
        return [smooth, conc, ratio, score]
    rate = np.abs(np.diff(f)) / (np.diff(t) + 0.1)
    med_err = np.median(e)
    if med_err > 0:
        smooth = 1.0 / (1.0 + np.median(rate) / med_err)
    peak, base = np.max(f), np.percentile(f, 10)
    if base > 0:
        conc = peak / base
    elif peak > 0:
        conc = peak / np.median(np.abs(f) + 1)
    if len(t) >= 20:
        short = [np.std(f[i:i + 5]) for i in range(len(t) - 5) if t[i + 5] - t[i] < 15]
        long_var = np.std(f)
        if short and long_var > 0:
            ratio = np.mean(short) / long_var
    scores = []
    if not np.isnan(smooth):
        scores.append(smooth)
    if not np.isnan(conc):
        scores.append(min(1.0, conc / 100))
    if not np.isnan(ratio):
        scores.append(1.0 - min(1.0, ratio))
    if scores:
        score = np.mean(scores)
    return [smooth, conc, ratio, score]


def color_at_peak_block(v: ObjectView):
    """compute_color_at_peak (:254-327); everything in FILE order (no sort upstream)."""
    out = [NAN] * 4
    rt, rf, _ = v.band_file(R)
    if len(rt) < 3:
        gt, gf, _ = v.band_file(G)
        if len(gt) < 3:
            return out
        peak_time = gt[int(np.nanargmax(gf))]
    else:
        peak_time = rt[int(np.nanargmax(rf))]
    for p, (b1, b2) in enumerate(((G, R), (R, I))):
        t1, f1, _ = v.band_file(b1)
        t2, f2, _ = v.band_file(b2)
        if len(t1) < 2 or len(t2) < 2:
            continue
        n1, n2 = np.abs(t1 - peak_time) < 10, np.abs(t2 - peak_time) < 10
        if not n1.any() or not n2.any():
            continue
        a = f1[n1][int(np.argmin(np.abs(t1[n1] - peak_time)))]
        b = f2[n2][int(np.argmin(np.abs(t2[n2] - peak_time)))]
        if a > 0 and b > 0:
            col = -2.5 * np.log10(a / b)
            out[2 * p] = col
            l1, l2 = t1 > peak_time + 50, t2 > peak_time + 50
            if l1.any() and l2.any():
                late = []
                for tt, ff in zip(t1[l1], f1[l1]):
                    dt = np.abs(t2[l2] - tt)
                    m = int(np.argmin(dt))
                    if dt[m] < 5:
                        f2l = f2[l2][m]
                        if ff > 0 and f2l > 0:
                            late.append(-2.5 * np.log10(ff / f2l))
                if late:
                    out[2 * p + 1] = np.mean(late) - col
    return out


def mexican_hat(scale: float, length: int) -> np.ndarray:
    """research_features.py:334-346 (note -length//2 is floor(-length/2))."""
    x = np.linspace(-length // 2, length // 2, length) / scale
    w = (1 - x ** 2) * np.exp(-x ** 2 / 2)
    return w / np.sqrt(np.sum(w ** 2))


def mhps_block(v: ObjectView):
    """compute_mhps_features (:349-425): r band on a 1-day grid, three wavelet scales."""
    out = [NAN] * 6
    t, f, _ = v.band_sorted(R)
    if len(t) < 20 or t[-1] - t[0] < 50:
        return out
    grid = np.arange(t[0], t[-1], 1.0)
    fr = np.interp(grid, t, f)
    fr = fr - np.mean(fr)
    power = {}
    for k, scale in enumerate(MHPS_SCALES):
        wl = int(min(5 * scale, len(fr) // 2))
        if wl < 5:
            continue
        conv = convolve(fr, mexican_hat(scale, wl), mode="same")
        power[scale] = np.sum(conv ** 2) / len(conv)
        out[k] = power[scale]
    if 10 in power and 100 in power and power[100] > 0:
        out[3] = power[10] / power[100]
    if 30 in power and 100 in power and power[100] > 0:
        out[4] = power[30] / power[100]
    if power:
        out[5] = float(max(power, key=power.get))
    return out


def luminosity_distance_mpc(z: float) -> float:
    """research_features.py:432-456."""
    if z <= 0 or np.isnan(z):
        return NAN
    if z < 0.1:
        return (C_KMS / H0) * z * (1 + z / 2)
    q0 = 0.5 * OMEGA_M - OMEGA_L
    return (C_KMS / H0) * z * (1 + 0.5 * (1 - q0) * z)


def luminosity_block(v: ObjectView, z: float):
    """compute_luminosity_features (:459-527), called only for Z > 0 (:556)."""
    out = [NAN] * 5
    if not (z > 0):
        return out
    d_l = luminosity_distance_mpc(z)
    if np.isnan(d_l):
        return out
    out[0] = d_l
    m = (v.b == G) | (v.b == R) | (v.b == I)
    if m.sum() < 5:
        return out
    o = np.argsort(v.t[m], kind="stable")
    t, lum = v.t[m][o], v.f[m][o] * d_l ** 2
    out[1] = np.max(lum)
    out[2] = np.max(lum) - np.percentile(lum, 10)
    out[3] = np.mean(lum)
    k = int(np.argmax(lum))
    if k < len(lum) - 5:
        pl, pt = lum[k:], t[k:]
        if len(pl) >= 3 and np.min(pl) > 0:
            dt = pt - pt[0]
            if np.std(dt) > 0:
                out[4] = np.polyfit(dt, np.log10(pl), 1)[0] * 100
    return out


def object_row(v: ObjectView, z: float):
    """extract_research_features_single (:534-563)."""
    return power_law_block(v) + nuclear_block(v) + color_at_peak_block(v) + mhps_block(v) + luminosity_block(v, z)
