"""The 2-D GP oracle is "parity unpinned" because george is absent (oracle/gp2d.py).  What CAN be pinned is its
mathematics: scikit-learn's GaussianProcessRegressor with ``ConstantKernel * Matern(nu=1.5, length_scale=[lt, lw])`` and
``alpha = yerr^2 + 1.25e-12`` is an independent implementation of the same model (a single Matern-3/2 of the combined
scaled distance; george's metric M_ii = l_i^2).  The oracle's log-likelihood, its gradient (chain rule: log M = 2 log l)
and its predictive mean must agree with sklearn's to rounding.  What stays an assumption (A-george) is only george's
parameter-vector ORDER and the ``c * kernel`` / ndim convention, i.e. which numbers the reference exports under which name."""
import warnings

import numpy as np
import pytest


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_gp2d_oracle_likelihood_gradient_and_prediction_match_sklearn(seed):
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern
    from mallorn_astrophysics_b200.synth import make_table
    from oracle import gp2d
    from oracle.common import group_objects
    tb = make_table(3, seed=100 + seed)
    rng = np.random.default_rng(seed)
    for _, v in group_objects(tb.obj, tb.band, tb.t, tb.flux, tb.err, tb.n_obj):
        t, lam, y, yerr, scale = gp2d.prepare(v)
        gp = gp2d.MaternGP(t, lam, y, yerr)
        p = gp.p0() + rng.normal(0, 0.5, 4)                      # [mean, log_constant, log M_t, log M_w] away from the start
        amp, lt, lw = gp2d.NDIM * np.exp(p[1]), np.exp(0.5 * p[2]), np.exp(0.5 * p[3])
        X = np.column_stack([t, lam])
        kern = ConstantKernel(amp, "fixed") * Matern(length_scale=[lt, lw], nu=1.5)
        sk = GaussianProcessRegressor(kernel=kern, alpha=yerr ** 2 + gp2d.TINY, optimizer=None)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sk.fit(X, y - p[0])
        nll, grad = gp.nll_and_grad(p)
        theta = np.log([lt, lw])                                 # the free hyper-parameters of `kern`
        ll_sk, g_sk = sk.log_marginal_likelihood(theta, eval_gradient=True)
        assert np.isclose(-nll, ll_sk, rtol=1e-10, atol=1e-8)
        # d/d log M = 1/2 d/d log l ; oracle returns -grad
        np.testing.assert_allclose(-grad[2:4], 0.5 * g_sk, rtol=1e-7, atol=1e-8)
        # amplitude derivative: d lnL / d log_constant = d lnL / d log amp (sklearn with the constant left free)
        kern2 = ConstantKernel(amp) * Matern(length_scale=[lt, lw], nu=1.5, length_scale_bounds="fixed")
        sk2 = GaussianProcessRegressor(kernel=kern2, alpha=yerr ** 2 + gp2d.TINY, optimizer=None).fit(X, y - p[0])
        _, g_amp = sk2.log_marginal_likelihood(np.log([amp]), eval_gradient=True)
        np.testing.assert_allclose(-grad[1], g_amp[0], rtol=1e-7, atol=1e-8)
        # mean derivative by central difference of the oracle itself (sklearn has no fitted mean)
        h = 1e-5
        fd = (gp.nll_and_grad(p + [h, 0, 0, 0])[0] - gp.nll_and_grad(p - [h, 0, 0, 0])[0]) / (2 * h)
        assert np.isclose(grad[0], fd, rtol=1e-5, atol=1e-7)
        # predictive mean at three (time, wavelength) points
        for tq, lq in ((t.mean(), 4825.0), (t.max() + 20.0, 6222.0), (t.min(), 7545.0)):
            want = sk.predict(np.array([[tq, lq]]))[0] + p[0]
            assert np.isclose(gp.predict(p, tq, lq), want, rtol=1e-8, atol=1e-9)
