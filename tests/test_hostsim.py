"""CPU coverage of the device arithmetic: the feature templates in csrc/*.cuh are instantiated with a
single-member HostTeam (tests/hostsim, g++ only, never shipped) and checked against the reference goldens /
the oracle.  This is what keeps trf.cuh, fits.cuh, lbfgs4.cuh and the scalar passes honest in a container
without a GPU; the CUDA instantiation of the same source is checked by the -m gpu tests."""
import ctypes
import os
import sys
import warnings

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from fitparity import classify_bazin, classify_gp1d, classify_r2   # noqa: E402
from packref import pack_reference                   # noqa: E402

P = ctypes.c_void_p
TIE_OBJECT = 15


@pytest.fixture(scope="module")
def hs():
    from hostsim.build import build
    return ctypes.CDLL(build())


def _run(lib, fn, ncol, pk, n_obj, z=None, status_cols=0):
    out = np.zeros((n_obj, ncol))
    st = np.zeros((n_obj, max(status_cols, 1)), np.int32)
    args = [ctypes.c_longlong(n_obj)] + [pk[k].ctypes.data_as(P) for k in ("seg_off", "t", "flux", "err", "rank")]
    if z is not None:
        args.append(np.ascontiguousarray(z).ctypes.data_as(P))
    args.append(out.ctypes.data_as(P))
    if status_cols:
        args.append(st.ctypes.data_as(P))
    getattr(lib, fn)(*args)
    return out, st


def _pack(g):
    n_obj = int(g["n_obj"])
    return pack_reference(g["obj"], g["band"], g["t"], g["flux"], g["err"], n_obj), n_obj


@pytest.mark.parametrize("module,fn,ncol", [("statistical", "hs_stat", 123), ("colors", "hs_color", 83), ("shape", "hs_shape", 65),
                                            ("tde", "hs_tde", 25), ("physics", "hs_physics", 32)])
@pytest.mark.parametrize("fixture", ["ref_sparse", "ref_edge"])
def test_scalar_templates_match_reference(golden, hs, fixture, module, fn, ncol):
    g = golden[fixture]
    pk, n_obj = _pack(g)
    out, _ = _run(hs, fn, ncol, pk, n_obj, z=None if module == "statistical" else g["z"])
    pres, ref = g[f"{module}_present"], g[f"{module}_rows"]
    keep = pres != (TIE_OBJECT if fixture == "ref_edge" else -1)
    got, ref = out[pres][keep], ref[keep]
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    np.testing.assert_allclose(got[ok], ref[ok], rtol=1e-7, atol=1e-9)


def test_trf_templates_follow_scipy(golden, hs):
    """Bazin / decline-model fits of the host instantiation vs the reference's own outputs (tests/fitparity.py)."""
    tb = {"fits": 0, "match": 0, "no_worse": 0, "worse": 0, "nan_diff": 0}
    tp = dict(tb)
    for fixture in ("ref_sparse", "ref_dense", "ref_edge"):
        g = golden[fixture]
        pk, n_obj = _pack(g)
        keep = g["bazin_present"] != (TIE_OBJECT if fixture == "ref_edge" else -1)
        ob, _ = _run(hs, "hs_bazin_t", 52, pk, n_obj, status_cols=12)
        op, _ = _run(hs, "hs_powerlaw_t", 27, pk, n_obj, status_cols=54)
        cb = classify_bazin(ob[g["bazin_present"]][keep], g["bazin_rows"][keep])
        cp = classify_r2(op[g["powerlaw_present"]][keep], g["powerlaw_rows"][keep])
        for k in tb:
            tb[k] += cb[k]; tp[k] += cp[k]
    assert tb["fits"] > 500 and tb["match"] >= 0.6 * tb["fits"] and tb["match"] + tb["no_worse"] >= 0.93 * tb["fits"], tb
    assert tb["nan_diff"] <= 0.01 * tb["fits"] + 1, tb
    assert tp["match"] >= 0.98 * tp["fits"] and tp["nan_diff"] == 0, tp


def test_lbfgs4_and_tile_sweep_follow_scipy(hs):
    """GP host model (same optimiser source as the GPU, same tile-sweep algebra) vs oracle/gp2d.py:
    identical iteration / evaluation counts and features to 1e-6."""
    from mallorn_astrophysics_b200.synth import make_table
    from oracle import gp2d
    from oracle.common import group_objects
    tb = make_table(8, config_index=2)
    pk = pack_reference(tb.obj, tb.band, tb.t, tb.flux, tb.err, tb.n_obj)
    out, st = _run(hs, "hs_gp2d", 27, pk, tb.n_obj, status_cols=4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for k, v in group_objects(tb.obj, tb.band, tb.t, tb.flux, tb.err, tb.n_obj):
            row, state = gp2d.object_row(v, want_state=True)
            assert (st[k, 0], st[k, 1]) == (state.nit, state.nfev), (k, st[k], state.nit, state.nfev)
            ok = ~np.isnan(row)
            assert np.array_equal(np.isnan(out[k]), np.isnan(row))
            np.testing.assert_allclose(out[k][ok], row[ok], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("which", [0, 1], ids=["product_dense_model", "exact_compact_form"])
def test_lbfgsb_follows_scipy_on_the_same_objective(hs, which):
    """The bounded optimisers driven through a callback with the oracle's own objective, which isolates them from the
    objective's rounding: csrc/lbfgsb.cuh (the product: explicit 3 x 3 quasi-Newton model) and the TEST-ONLY
    routine-by-routine restatement of the published compact-form code (tests/hostsim/lbfgsb_exact.cuh).
    Iteration and evaluation counts must equal SciPy's on (nearly) every start."""
    from scipy.optimize import minimize
    from mallorn_astrophysics_b200.synth import make_table
    from oracle import gp1d
    from oracle.common import group_objects
    CB = ctypes.CFUNCTYPE(ctypes.c_double, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double))
    lo = np.ascontiguousarray(gp1d.LOG_BOUNDS[:, 0]); hi = np.ascontiguousarray(gp1d.LOG_BOUNDS[:, 1])
    tb = make_table(10, config_index=1)
    total = same = at_bound = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _, v in group_objects(tb.obj, tb.band, tb.t, tb.flux, tb.err, tb.n_obj):
            for b in gp1d.BANDS:
                p = gp1d.prepare(*v.band_sorted(b))
                if p is None:
                    continue
                x, y, alpha = p[0], p[1], p[2]
                d2 = (x[:, None] - x[None, :]) ** 2

                def obj(xp, gp):
                    fv, g = gp1d.neg_lml(np.array([xp[0], xp[1], xp[2]]), d2, y, alpha)
                    gp[0], gp[1], gp[2] = g
                    return float(fv)
                cb = CB(obj)
                for start in (gp1d.THETA0, *gp1d.RESTARTS):
                    res = minimize(gp1d.neg_lml, start, args=(d2, y, alpha), method="L-BFGS-B", jac=True, bounds=gp1d.LOG_BOUNDS)
                    xo = np.zeros(3); fo = ctypes.c_double(); cnt = np.zeros(3, np.int32)
                    s0 = np.ascontiguousarray(start)
                    hs.hs_lbfgsb3(ctypes.c_int(which), cb, s0.ctypes.data_as(P), lo.ctypes.data_as(P), hi.ctypes.data_as(P), xo.ctypes.data_as(P),
                                  ctypes.byref(fo), cnt.ctypes.data_as(P))
                    total += 1
                    same += (cnt[0] == res.nit and cnt[1] == res.nfev)
                    at_bound += bool(np.any((res.x <= lo) | (res.x >= hi)))
                    assert fo.value <= res.fun + 1e-6 * abs(res.fun) + 1e-9, (fo.value, res.fun)
    assert total >= 120 and same >= 0.97 * total, (same, total)
    assert at_bound >= 0.3 * total          # the bound logic (Cauchy point, projected subspace step) is really exercised


@pytest.mark.parametrize("fixture", ["ref_sparse", "ref_edge"])
def test_per_band_gp_templates_match_reference(golden, hs, fixture):
    """gp1d.cuh (symmetric sweep + lbfgsb.cuh, four starts) host instantiation vs the reference's sklearn output."""
    g = golden[fixture]
    pk, n_obj = _pack(g)
    out, st = _run(hs, "hs_gp1d", 21, pk, n_obj, status_cols=12)
    pres, ref = g["gp1d_present"], g["gp1d_rows"]
    keep = pres != (TIE_OBJECT if fixture == "ref_edge" else -1)
    got, ref = out[pres][keep], ref[keep]
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    c = classify_gp1d(got, ref)                                    # tests/fitparity.py: the north star's fit rule
    assert c["worse"] == 0 and c["nan_diff"] == 0 and c["cross_bad"] == 0 and c["match"] >= 0.98 * c["fits"], c
    fitted = st[pres][keep][:, 2::3] >= 0
    assert fitted.sum() > 50 and (st[pres][keep][:, 0::3][fitted] >= 4).all()


def test_per_band_gp_crafted_edge_cases(hs):
    """constant flux, a single epoch, NaN rows, non-positive errors, duplicated times, exactly five points, huge / tiny
    errors: host instantiation and oracle vs the reference's own output (tests/golden/gp1d_crafted.npz)"""
    from oracle.extract import run_module
    g = np.load(os.path.join(HERE, "golden", "gp1d_crafted.npz"))
    pk, n_obj = _pack(g)
    out, st = _run(hs, "hs_gp1d", 21, pk, n_obj, status_cols=12)
    ref = g["gp1d_rows"]
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    assert np.isnan(ref[1]).all() and np.isnan(ref[2]).all() and np.isfinite(ref[0, 0:4]).all()
    ok = ~np.isnan(ref)
    np.testing.assert_allclose(out[ok], ref[ok], rtol=1e-5, atol=1e-9)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, rows = run_module("gp1d", g["obj"], g["band"], g["t"], g["flux"], g["err"], n_obj)
    np.testing.assert_allclose(rows[ok], ref[ok], rtol=1e-5, atol=1e-9)




@pytest.mark.parametrize("fixture", ["ref_sparse", "ref_dense", "ref_edge", "ref_c1_256"])
def test_research_template_matches_reference(golden, hs, fixture):
    """research.cuh (power-law quality, nuclear proxies, colour at peak, Mexican-hat power spectra,
    luminosity) host instantiation vs the reference's own frames"""
    g = golden[fixture]
    pk, n_obj = _pack(g)
    out, _ = _run(hs, "hs_research", 40, pk, n_obj, z=g["z"])
    pres, ref = g["research_present"], g["research_rows"]
    keep = pres != (TIE_OBJECT if fixture == "ref_edge" else -1)
    got, ref = out[pres][keep], ref[keep]
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    np.testing.assert_allclose(got[ok], ref[ok], rtol=1e-7, atol=1e-9)


def test_division_helpers_of_the_fit_models_equal_ieee_division(hs):
    """``div_r`` (a / d from a correctly rounded reciprocal: one multiplication and a residual correction) over the operand
    ranges of a fit pass, and ``div_curve`` (exact shortcuts for a zero numerator / an infinite denominator) on exactly
    those operands, against the IEEE quotient."""
    rng = np.random.default_rng(12)
    n = 2_000_000
    d = np.ldexp(rng.uniform(0.5, 1.0, n), rng.integers(-20, 20, n))                    # time scales, finite-difference steps
    a = np.ldexp(rng.uniform(0.5, 1.0, n), rng.integers(-30, 30, n)) * rng.choice([-1.0, 1.0], n)
    a[:1000] = 0.0
    with np.errstate(over="ignore", under="ignore"):
        den = 1.0 + np.exp(rng.uniform(-800.0, 800.0, n))                               # [1, +inf]
        num = np.exp(rng.uniform(-800.0, 700.0, n)) * rng.uniform(0.0, 1e4, n)          # [0, large], zeros by underflow
    P = ctypes.POINTER(ctypes.c_double)
    out = [np.empty(n) for _ in range(3)]
    with np.errstate(all="ignore"):
        hs.hs_div_helpers(ctypes.c_longlong(n), a.ctypes.data_as(P), d.ctypes.data_as(P), *[o.ctypes.data_as(P) for o in out])
        assert np.array_equal(out[0].view(np.uint64), out[2].view(np.uint64))
        hs.hs_div_helpers(ctypes.c_longlong(n), num.ctypes.data_as(P), den.ctypes.data_as(P), *[o.ctypes.data_as(P) for o in out])
        assert np.isinf(den).any() and (num == 0.0).any()
        assert np.array_equal(out[1].view(np.uint64), out[2].view(np.uint64))
