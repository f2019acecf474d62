"""GPU parity of the K8 pass ("research" block, 40 columns; reference src/features/research_features.py) against the
reference's own frames (tests/golden/*.npz: research_rows).  The arithmetic is the template that tests/test_hostsim.py
holds to these goldens on the CPU."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TIE_OBJECT = 15


@pytest.mark.parametrize("fixture", ["ref_sparse", "ref_dense", "ref_edge", "ref_c1_256"])
def test_research_pass_matches_the_reference_goldens(golden, fixture):
    from mallorn_astrophysics_b200.store import LightCurveStore
    g = golden[fixture]
    with LightCurveStore(g["obj"], g["band"], g["t"], g["flux"], g["err"], int(g["n_obj"])) as s:
        rows = s.features("research", z=g["z"])
    pres, ref = g["research_present"], g["research_rows"]
    keep = pres != (TIE_OBJECT if fixture == "ref_edge" else -1)
    got, ref = rows[pres][keep], ref[keep]
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    np.testing.assert_allclose(got[ok], ref[ok], rtol=1e-5, atol=1e-9)


def test_research_frame_keeps_the_reference_schema(golden):
    import pandas as pd
    from mallorn_astrophysics_b200 import schema
    from mallorn_astrophysics_b200.features.research_features import extract_research_features
    g = golden["ref_edge"]
    n_obj = int(g["n_obj"])
    ids = np.array([f"Obj_{k:04d}" for k in range(n_obj)], dtype=object)
    bands = np.array(["u", "g", "r", "i", "z", "y", "X"], dtype=object)
    lc = pd.DataFrame({"object_id": ids[g["obj"]], "Time (MJD)": g["t"], "Flux": g["flux"], "Flux_err": g["err"],
                       "Filter": bands[np.minimum(g["band"], 6)]})
    meta = pd.DataFrame({"object_id": ids, "Z": g["z"], "EBV": g["ebv"]})
    df = extract_research_features(lc, list(ids), meta, verbose=False)
    assert list(df.columns) == schema.RESEARCH_COLUMNS + ["object_id"]
    assert list(df["object_id"]) == [i for i in ids if i != ids[20]]
    no_meta = extract_research_features(lc, list(ids), None, verbose=False)
    assert no_meta[schema.RESEARCH_COLUMNS[35:]].isna().all().all()          # no redshift -> luminosity block NaN
