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

from fitparity import classify_bazin, classify_r2   # noqa: E402
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
