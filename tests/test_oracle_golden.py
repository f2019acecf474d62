"""The CPU oracle must reproduce the reference's own outputs (tests/golden/*.npz, written by
tests/golden/make_golden.py from the unmodified reference) before it is trusted as the checker."""
import warnings

import numpy as np
import pytest

from oracle.extract import run_module

TIE_OBJECT = 15   # ref_edge object with duplicated observation times: the reference's default
                  # (unstable) quicksort leaves the order of equal times unspecified


def _check(g, module, rtol, skip_obj=()):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        present, rows = run_module(module, g["obj"], g["band"], g["t"], g["flux"], g["err"],
                                   int(g["n_obj"]), g["z"])
    ref, ref_present = g[f"{module}_rows"], g[f"{module}_present"]
    assert np.array_equal(present, ref_present)          # absent ids are omitted, order kept
    keep = ~np.isin(present, skip_obj)
    rows, ref = rows[keep], ref[keep]
    assert np.array_equal(np.isnan(rows), np.isnan(ref))
    ok = ~np.isnan(ref)
    np.testing.assert_allclose(rows[ok], ref[ok], rtol=rtol, atol=0)


@pytest.mark.parametrize("module", ["statistical", "colors", "shape", "tde", "physics"])
@pytest.mark.parametrize("fixture", ["ref_sparse", "ref_dense", "ref_edge", "ref_c1_256"])
def test_scalar_modules_match_reference(golden, fixture, module):
    _check(golden[fixture], module, rtol=1e-13, skip_obj=(TIE_OBJECT,) if fixture == "ref_edge" else ())


@pytest.mark.parametrize("module", ["bazin", "powerlaw"])
@pytest.mark.parametrize("fixture", ["ref_sparse", "ref_edge"])
def test_fit_modules_match_reference(golden, fixture, module):
    # same scipy call on the same inputs: bit-for-bit in practice, 1e-12 leaves room for BLAS order
    _check(golden[fixture], module, rtol=1e-12, skip_obj=(TIE_OBJECT,) if fixture == "ref_edge" else ())


@pytest.mark.parametrize("fixture", ["ref_sparse", "ref_edge"])
def test_per_band_gp_matches_reference(golden, fixture):
    # numpy restatement of what sklearn's GaussianProcessRegressor does, same scipy L-BFGS-B: agrees to ~1e-9 in
    # practice (LAPACK Cholesky vs cho_solve ordering); 1e-6 leaves room for a different BLAS
    _check(golden[fixture], "gp1d", rtol=1e-6, skip_obj=(TIE_OBJECT,) if fixture == "ref_edge" else ())


@pytest.mark.parametrize("fixture", ["ref_sparse", "ref_dense", "ref_edge", "ref_c1_256"])
def test_research_block_matches_reference(golden, fixture):
    # SURVEY 8(f) row 3, oracle stage: research_features.py (v115c) restated with the reference's own numpy / scipy calls
    _check(golden[fixture], "research", rtol=1e-12, skip_obj=(TIE_OBJECT,) if fixture == "ref_edge" else ())


def test_absent_object_is_omitted(golden):
    g = golden["ref_edge"]
    assert 20 not in g["statistical_present"]            # object 20 has no rows (statistical.py:163-165)
    present, _ = run_module("statistical", g["obj"], g["band"], g["t"], g["flux"], g["err"], int(g["n_obj"]))
    assert 20 not in present and len(present) == int(g["n_obj"]) - 1
