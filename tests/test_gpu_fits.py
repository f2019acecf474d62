"""GPU parity of the K4 fit passes (Bazin, decline models) against the reference's own outputs
(tests/golden, produced by the unmodified reference + scipy) and the CPU oracle.  Rule: tests/fitparity.py."""
import warnings

import numpy as np
import pytest

from fitparity import classify_bazin, classify_r2

pytestmark = pytest.mark.gpu
TIE_OBJECT = 15


def _store(g):
    from mallorn_astrophysics_b200.store import LightCurveStore
    return LightCurveStore(g["obj"], g["band"], g["t"], g["flux"], g["err"], int(g["n_obj"]))


def _golden(golden, fixture, module, got):
    g = golden[fixture]
    present, want = g[f"{module}_present"], g[f"{module}_rows"]
    got = got[present]
    if fixture == "ref_edge":
        keep = present != TIE_OBJECT
        got, want = got[keep], want[keep]
    return got, want


@pytest.mark.parametrize("fixture", ["ref_sparse", "ref_dense", "ref_edge", "ref_c1_256"])
def test_bazin_matches_reference_golden(golden, fixture):
    with _store(golden[fixture]) as s:
        status = np.zeros((s.n_obj, 6, 2), np.int32)
        got = s.features("bazin", status=status)
    got, want = _golden(golden, fixture, "bazin", got)
    c = classify_bazin(got, want)
    print(fixture, c)
    assert c["fits"] > 50
    # per fixture (100-760 fits each, so a few points of sampling noise); the aggregate rule is in the next test
    assert c["match"] >= 0.62 * c["fits"], c
    assert c["worse"] <= 0.06 * c["fits"] + 1, c
    assert c["match"] + c["no_worse"] >= 0.93 * c["fits"], c
    assert c["nan_diff"] <= 0.02 * c["fits"] + 1, c
    # bands the reference never fits (n < 5) are NaN on both sides and flagged FIT_SKIPPED
    assert (status[..., 1][status[..., 0] == -200] == 0).all()


def test_bazin_classes_over_all_reference_fits(golden):
    """The measured level, not a loose floor: over the 2,047 fits the unmodified reference attempted on the four fixtures the
    B200 lands within rtol 1e-3 on 72 %, on a different point that is no worse on 25 %, on a worse one on 3 %
    (DESIGN.md section 5).  That split is the reference's OWN reproducibility under a 1-ulp change of exp()
    (70.4 / 27.1 / 2.4, tests/test_fit_ceiling.py), so the assertions sit just under it: a regression of a few points fails."""
    tot = {"fits": 0, "match": 0, "no_worse": 0, "worse": 0, "nan_diff": 0}
    for fixture in ("ref_sparse", "ref_dense", "ref_edge", "ref_c1_256"):
        with _store(golden[fixture]) as s:
            got = s.features("bazin")
        c = classify_bazin(*_golden(golden, fixture, "bazin", got))
        for k in tot:
            tot[k] += c[k]
    print(tot)
    assert tot["fits"] > 2000
    assert tot["match"] >= 0.68 * tot["fits"], tot
    assert tot["worse"] <= 0.04 * tot["fits"], tot
    assert tot["nan_diff"] <= 0.01 * tot["fits"], tot


@pytest.mark.parametrize("fixture", ["ref_sparse", "ref_dense", "ref_edge", "ref_c1_256"])
def test_powerlaw_matches_reference_golden(golden, fixture):
    with _store(golden[fixture]) as s:
        got = s.features("powerlaw")
    got, want = _golden(golden, fixture, "powerlaw", got)
    c = classify_r2(got, want)
    print(fixture, c)
    assert c["match"] >= 0.97 * c["fits"], c
    # a different local minimum in < 0.5 % of fits; the largest R^2 deficit seen on any set so far is 0.133
    assert c["worse"] <= 0.005 * c["fits"] + 1 and c["worst"] < 0.15, c
    assert c["nan_diff"] <= 0.005 * c["fits"] + 1, c


def test_bazin_cross_band_columns_follow_the_band_columns(golden):
    """bazin_fitting.py:217-249 recomputed from our own band columns must equal columns 48..51 exactly."""
    from oracle.bazin import cross_band
    with _store(golden["ref_sparse"]) as s:
        got = s.features("bazin")
    for row in got[golden["ref_sparse"]["bazin_present"]]:
        chk = row.copy()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            cross_band(chk)
        np.testing.assert_allclose(chk[48:], row[48:], rtol=1e-12, equal_nan=True)


def test_fits_match_oracle_at_scale():
    """1024 seeded objects: the same classification against the scipy oracle (about 6 CPU-minutes / 8 cores
    would be too slow single-threaded, so the oracle runs on a 192-object slice)."""
    from mallorn_astrophysics_b200.store import LightCurveStore
    from mallorn_astrophysics_b200.synth import make_table
    from oracle.extract import run_table
    table = make_table(192, config_index=4)
    with LightCurveStore.from_table(table) as s:
        gb, gp = s.features("bazin"), s.features("powerlaw")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        present, wb = run_table("bazin", table)
        _, wp = run_table("powerlaw", table)
    cb, cp = classify_bazin(gb[present], wb), classify_r2(gp[present], wp)
    print(cb, cp)
    assert cb["match"] >= 0.62 * cb["fits"] and cb["match"] + cb["no_worse"] >= 0.93 * cb["fits"], cb
    assert cb["worse"] <= 0.05 * cb["fits"] and cb["nan_diff"] <= 0.01 * cb["fits"] + 1, cb
    assert cp["match"] >= 0.97 * cp["fits"] and cp["worst"] < 0.15, cp
