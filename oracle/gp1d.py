"""Oracle for the per-band 1-D GP features (``extract_gp_features``).  TEST INFRASTRUCTURE ONLY.

Restates ``src/features/gaussian_process.py:29-248`` on flat arrays.  The reference delegates the
model to scikit-learn (``GaussianProcessRegressor(ConstantKernel * RBF + WhiteKernel, alpha=err^2,
n_restarts_optimizer=3, random_state=42)``, :88-104); what sklearn 1.9 does with that is restated
here in numpy so the optimiser trace (iterations / evaluations of each of the four starts) is
observable:

* theta = ln [constant_value, length_scale, noise_level], bounds ln [(0.01, 100), (0.01, 2), (1e-5, 10)];
* K = c exp(-d^2 / (2 l^2)) + noise I + diag(alpha); ln L = -1/2 y' K^-1 y - sum ln L_ii - n/2 ln 2 pi;
  d ln L / d theta_j = 1/2 tr((a a' - K^-1) dK/dtheta_j)  (``_gpr.py`` log_marginal_likelihood);
* four L-BFGS-B starts (``scipy.optimize.minimize(method='L-BFGS-B', jac=True, bounds=...)`` with SciPy's
  defaults): the kernel's own theta, then three draws ``RandomState(42).uniform(lo, hi)`` -- the generator
  is re-created from ``random_state=42`` in every ``fit``, so the three restart points are constants;
* the winner is the first start with the smallest objective.

**Pinned**: sklearn is installed, so ``tests/golden/make_golden.py`` runs the unmodified reference
module and ``tests/test_oracle_golden.py`` holds this file to its output.

Row layout is ``schema.GP1D_COLUMNS`` (21 values).
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import cho_solve, cholesky
from scipy.optimize import minimize

from .common import NAN, ObjectView, G, R, I, Z

BANDS = (G, R, I, Z)                                      # gaussian_process.py:187
LOG_BOUNDS = np.log(np.array([[0.01, 100.0], [0.01, 2.0], [1e-5, 10.0]]))
THETA0 = np.log(np.array([1.0, 0.2, 0.1]))                # :88-92
# RandomState(42).uniform(LOG_BOUNDS[:, 0], LOG_BOUNDS[:, 1]), three successive draws
RESTARTS = np.array([
    [float.fromhex("-0x1.27d0b258dec3ep+0"), float.fromhex("0x1.ba6262947e1f0p-2"), float.fromhex("-0x1.666a086b81fa0p+0")],
    [float.fromhex("0x1.d13e458b4d080p-1"), float.fromhex("-0x1.e3a6ffd15b375p+1"), float.fromhex("-0x1.2b72f23c29e62p+3")],
    [float.fromhex("-0x1.047e29999566cp+2"), float.fromhex("-0x1.04689013a2800p-6"), float.fromhex("-0x1.9aa6c74700fb8p+1")],
])
LOG_2PI = float(np.log(2.0 * np.pi))


def prepare(t, f, e):
    """gaussian_process.py:54-85 -> (x, y, alpha, t_range, f_std) or None."""
    if len(t) < 5:
        return None
    valid = ~(np.isnan(f) | np.isnan(e) | (e <= 0))
    if valid.sum() < 5:
        return None
    t, f, e = t[valid], f[valid], e[valid]
    t_min, t_max = t.min(), t.max()
    t_range = t_max - t_min
    if t_range == 0:
        return None
    x = (t - t_min) / t_range
    f_mean, f_std = np.mean(f), np.std(f)
    if f_std == 0:
        f_std = 1
    y = (f - f_mean) / f_std
    alpha = ((e / f_std) ** 2).clip(min=1e-10)
    return x, y, alpha, t_range, f_std


def neg_lml(theta, d2, y, alpha):
    """objective handed to L-BFGS-B: (-ln L, -grad) at theta (sklearn _gpr.py obj_func)."""
    c, l, s = np.exp(theta)
    q = d2 / (l * l)
    C = c * np.exp(-0.5 * q)
    K = C + np.diag(s + alpha)
    try:
        L = cholesky(K, lower=True, check_finite=False)
    except np.linalg.LinAlgError:
        return np.inf, np.zeros(3)
    a = cho_solve((L, True), y, check_finite=False)
    lml = -0.5 * y.dot(a) - np.log(np.diag(L)).sum() - 0.5 * len(y) * LOG_2PI
    W = np.outer(a, a) - cho_solve((L, True), np.eye(len(y)), check_finite=False)
    grad = 0.5 * np.array([(W * C).sum(), (W * C * q).sum(), s * np.trace(W)])
    return -lml, -grad


def fit_band(t, f, e, trace=None):
    """gaussian_process.py:29-146 -> [length_scale_days, amplitude, noise, log_likelihood]."""
    p = prepare(t, f, e)
    if p is None:
        return [NAN] * 4
    x, y, alpha, t_range, f_std = p
    d2 = (x[:, None] - x[None, :]) ** 2
    best = None
    for k, start in enumerate([THETA0, *RESTARTS]):
        res = minimize(neg_lml, start, args=(d2, y, alpha), method="L-BFGS-B", jac=True, bounds=LOG_BOUNDS)
        if trace is not None:
            trace.append((int(res.nit), int(res.nfev), float(res.fun), res.x.copy(), int(res.status)))
        if best is None or res.fun < best.fun:
            best = res
    c, l, s = np.exp(best.x)
    lml = -neg_lml(best.x, d2, y, alpha)[0]
    return [l * t_range, np.sqrt(c) * f_std, np.sqrt(s) * f_std, lml]


def object_row(v: ObjectView, trace=None):
    """extract_gp_features_single (gaussian_process.py:171-248)."""
    row, fitted = [], {}
    for b in BANDS:
        t, f, e = v.band_sorted(b)
        if len(t) >= 5:
            tr = [] if trace is not None else None
            r = fit_band(t, f, e, tr)
            if trace is not None:
                trace.append((b, tr))
            fitted[b] = r
        else:
            r = [NAN] * 4
        row += r
    for b1, b2 in ((G, R), (R, I)):
        if b1 in fitted and b2 in fitted:
            l1, l2 = fitted[b1][0], fitted[b2][0]
            row.append(l1 / l2 if (not np.isnan(l1) and not np.isnan(l2) and l2 > 0) else NAN)
        else:
            row.append(NAN)
    ls = [fitted[b][0] for b in (G, R, I) if b in fitted and not np.isnan(fitted[b][0])]
    row.append(float(np.mean(ls)) if ls else NAN)
    row.append((float(np.std(ls)) if len(ls) > 1 else 0.0) if ls else NAN)
    amp = [fitted[b][1] for b in (G, R, I) if b in fitted and not np.isnan(fitted[b][1])]
    row.append(float(np.mean(amp)) if amp else NAN)
    return row
