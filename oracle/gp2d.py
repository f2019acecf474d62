"""Oracle for the 2-D (time x wavelength) multi-band GP features.  TEST INFRASTRUCTURE ONLY.

** PARITY UNPINNED **  The reference (``src/features/multiband_gp.py:34-344``) delegates
the kernel, Cholesky, likelihood and gradient to ``george>=0.4`` (requirements.txt:16),
a third-party package that is not vendored under ``/root/reference`` and not installed
here, and the reference ships no test vector for this path.  This module restates
george 0.4's published semantics -- assumption **A-george** (SURVEY.md section 8c):

1. ``Matern32Kernel(metric=[a, b], ndim=2)``: axis-aligned metric with parameters
   ``log_M_0_0 = ln a``, ``log_M_1_1 = ln b``; ``r2 = sum_i dx_i^2 / M_ii``;
   ``k = (1 + sqrt(3 r2)) exp(-sqrt(3 r2))``; ``dk/dlogM_ii = 1.5 exp(-sqrt(3 r2)) dx_i^2 / M_ii``.
2. ``c * kernel`` = ``ConstantKernel(log_constant = ln(c / ndim), ndim) * kernel``; the constant
   kernel is summed over its ``ndim`` axes, so the amplitude is ``ndim * exp(log_constant)``.
3. ``GP(kernel, mean=m, fit_mean=True)``; white noise frozen at ``ln(1.25e-12)``;
   ``K = K_signal + diag(yerr^2 + 1.25e-12)``.
4. Parameter vector order ``[mean:value, log_constant, log_M_0_0, log_M_1_1]``.
5. ``ln L = -1/2 r^T K^-1 r - 1/2 ln|K| - N/2 ln(2 pi)``; gradient: mean -> ``sum(alpha)``,
   kernel -> ``1/2 tr((alpha alpha^T - K^-1) dK/dtheta)``; ``predict`` = ``k(x*, X) alpha + mean``.

The optimiser is the reference's own call: ``scipy.optimize.minimize(method='L-BFGS-B',
options={'maxiter': 100})`` (multiband_gp.py:158-164).  The reference then reads
parameter-vector positions 0, 1, 2 under the names amplitude / time metric / wave metric
(:171-188); ``FEATURE_SLOTS`` keeps that positional mapping in one place.

The mathematics of this file (likelihood, gradient, predictive mean) is cross-checked against scikit-learn's
independent implementation of the same model in ``tests/test_gp2d_math_crosscheck.py``; the assumption that
remains is george's parameter order and amplitude convention (items 2 and 4).

Row layout is ``schema.GP2D_COLUMNS`` (27 values).
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import cho_factor, cho_solve
from scipy.optimize import minimize

from .common import NAN, ObjectView, G, R, I

WAVE_A = np.array([3670.0, 4825.0, 6222.0, 7545.0, 8691.0, 9710.0])    # multiband_gp.py:26-29
TINY = 1.25e-12
NDIM = 2
FEATURE_SLOTS = (0, 1, 2)          # positions of "log_amp", "log_metric_time", "log_metric_wave" (:174-176)
EPOCHS = (0.0, 20.0, 50.0, 100.0)
LOG_2PI = float(np.log(2.0 * np.pi))


def prepare(v: ObjectView):
    """multiband_gp.py:34-87 -> (t, lam, y, yerr, scale) or None."""
    keep = (v.b < 6) & ~np.isnan(v.f) & ~np.isnan(v.e) & (v.e > 0)
    if keep.sum() < 10:
        return None
    t, f, e = v.t[keep], v.f[keep], v.e[keep]
    lam = WAVE_A[v.b[keep]]
    t = t - t.min()
    nz = f[f != 0]
    scale = np.median(np.abs(nz))       # NaN when every flux is exactly zero, as in the reference
    if scale == 0:
        scale = 1.0
    return t, lam, f / scale, e / scale, scale


class MaternGP:
    """numpy restatement of the george objects the reference drives."""

    def __init__(self, t, lam, y, yerr):
        self.y = y
        self.n = y.shape[0]
        self.dt2 = (t[:, None] - t[None, :]) ** 2
        self.dl2 = (lam[:, None] - lam[None, :]) ** 2
        self.noise = yerr ** 2 + TINY
        self.t, self.lam = t, lam

    def p0(self, init_time_scale=100.0, init_wave_scale=6000.0):
        return np.array([np.mean(self.y), np.log(np.var(self.y) / NDIM),
                         np.log(init_time_scale ** 2), np.log(init_wave_scale ** 2)])

    def _kernel(self, p):
        amp = NDIM * np.exp(p[1])
        m0, m1 = np.exp(p[2]), np.exp(p[3])
        a0, a1 = self.dt2 / m0, self.dl2 / m1
        s = np.sqrt(3.0 * (a0 + a1))
        ex = np.exp(-s)
        return amp, a0, a1, s, ex

    def nll_and_grad(self, p):
        """(-ln L, -grad ln L); (1e25, 0) where george would report a non-finite likelihood."""
        try:
            with np.errstate(all="ignore"):
                amp, a0, a1, s, ex = self._kernel(p)
                ks = amp * (1.0 + s) * ex
                K = ks + np.diag(self.noise)
                cf = cho_factor(K, lower=True)
                r = self.y - p[0]
                alpha = cho_solve(cf, r)
                ll = -0.5 * (r @ alpha) - np.sum(np.log(np.diag(cf[0]))) - 0.5 * self.n * LOG_2PI
                if not np.isfinite(ll):
                    return 1e25, np.zeros(4)
                A = np.outer(alpha, alpha) - cho_solve(cf, np.eye(self.n))
                dk = 1.5 * amp * ex
                g = np.array([np.sum(alpha), 0.5 * np.sum(A * ks),
                              0.5 * np.sum(A * dk * a0), 0.5 * np.sum(A * dk * a1)])
                return -ll, -g
        except Exception:
            return 1e25, np.zeros(4)

    def predict(self, p, tq, lq):
        amp = NDIM * np.exp(p[1])
        K = self.nll_matrix(p)
        alpha = cho_solve(cho_factor(K, lower=True), self.y - p[0])
        s = np.sqrt(3.0 * ((tq - self.t) ** 2 / np.exp(p[2]) + (lq - self.lam) ** 2 / np.exp(p[3])))
        return float((amp * (1.0 + s) * np.exp(-s)) @ alpha + p[0])

    def nll_matrix(self, p):
        amp, _, _, s, ex = self._kernel(p)
        return amp * (1.0 + s) * ex + np.diag(self.noise)


def fit(gp: MaternGP, want_state=False):
    """multiband_gp.py:141-188 -> (p_opt, 5 hyper-parameter features[, scipy result])."""
    res = minimize(lambda p: gp.nll_and_grad(p)[0], gp.p0(), jac=lambda p: gp.nll_and_grad(p)[1],
                   method="L-BFGS-B", options={"maxiter": 100})
    p = res.x
    amp = np.exp(p[FEATURE_SLOTS[0]])
    ts = np.sqrt(np.exp(p[FEATURE_SLOTS[1]]))
    ws = np.sqrt(np.exp(p[FEATURE_SLOTS[2]]))
    feats = np.array([amp, ts, ws, -res.fun, ts / (ws / 1000.0)])
    return (p, feats, res) if want_state else (p, feats)


def peak_time(v: ObjectView) -> float:
    """multiband_gp.py:331-338: r-band ``idxmax`` (first max, NaN skipped) else global; minus min over ALL rows."""
    m = v.b == R
    if m.sum() > 0:
        k = int(np.nanargmax(v.f[m]))
        return v.t[m][k] - v.t.min()
    return v.t[int(np.nanargmax(v.f))] - v.t.min()


def object_row(v: ObjectView, want_state=False) -> np.ndarray:
    row = np.full(27, NAN)
    prep = prepare(v)
    state = None
    if prep is not None:
        t, lam, y, yerr, scale = prep
        try:
            gp = MaternGP(t, lam, y, yerr)
            p, feats, state = fit(gp, want_state=True)
            row[0:5] = feats
            pt = peak_time(v)
            col = np.full((4, 2), NAN)
            for ei, ep in enumerate(EPOCHS):                          # :236-262
                fl = [gp.predict(p, pt + ep, WAVE_A[b]) * scale for b in (G, R, I)]
                row[5 + 5 * ei:8 + 5 * ei] = fl
                if fl[0] > 0 and fl[1] > 0:
                    col[ei, 0] = -2.5 * np.log10(fl[0] / fl[1])
                if fl[1] > 0 and fl[2] > 0:
                    col[ei, 1] = -2.5 * np.log10(fl[1] / fl[2])
                row[8 + 5 * ei:10 + 5 * ei] = col[ei]
            if not np.isnan(col[0, 0]) and not np.isnan(col[2, 0]):   # :265-277
                row[25] = (col[2, 0] - col[0, 0]) / 50.0
            if not np.isnan(col[0, 0]) and not np.isnan(col[3, 0]):
                row[26] = (col[3, 0] - col[0, 0]) / 100.0
        except Exception:
            row[:] = NAN
    return (row, state) if want_state else row
