"""Feature interactions (reference src/features/interactions.py:23-196): the oracle and the host-side op list against the
unmodified reference's outputs (tests/golden/interactions.npz), and -- on the GPU -- the element-wise kernel against both."""
import numpy as np
import pytest

CASES = (("full", "inputs"), ("part", "present"))


@pytest.fixture(scope="module")
def gold():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "interactions.npz"))


def _cols(gold, which):
    names = [str(x) for x in gold["inputs"]]
    keep = [str(x) for x in gold[which]]
    return {n: gold["X"][:, names.index(n)] for n in keep}, keep


@pytest.mark.parametrize("case,which", CASES)
def test_oracle_matches_the_reference(gold, case, which):
    from oracle import interactions
    cols, keep = _cols(gold, which)
    got = interactions.create(cols, keep)
    assert list(got) == [str(x) for x in gold[f"{case}_names"]]            # same columns, same insertion order
    want = gold[case]
    for j, name in enumerate(got):
        assert np.array_equal(got[name], want[:, j], equal_nan=True), name   # same numpy expressions: bit for bit


@pytest.mark.parametrize("case,which", CASES)
def test_op_list_follows_the_reference_order(gold, case, which):
    from mallorn_astrophysics_b200.features.interactions import interaction_ops
    _, keep = _cols(gold, which)
    assert [o[0] for o in interaction_ops(keep)] == [str(x) for x in gold[f"{case}_names"]]


@pytest.mark.gpu
@pytest.mark.parametrize("case,which", CASES)
def test_kernel_matches_the_reference(gold, case, which):
    import torch
    from mallorn_astrophysics_b200.features.interactions import interactions_device
    cols, keep = _cols(gold, which)
    table = torch.from_numpy(np.ascontiguousarray(np.column_stack([cols[n] for n in keep]))).cuda()
    new, names = interactions_device(table, keep)
    assert names == [str(x) for x in gold[f"{case}_names"]]
    got, want = new.cpu().numpy(), gold[case]
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    fin = ok & np.isfinite(want)
    assert np.array_equal(got[ok & ~fin], want[ok & ~fin])                  # infinities with the right sign
    np.testing.assert_allclose(got[fin], want[fin], rtol=1e-14, atol=0)    # products and quotients exact, pow within 2 ulp


@pytest.mark.gpu
def test_dataframe_drop_in_keeps_the_reference_frame(gold):
    import pandas as pd
    from mallorn_astrophysics_b200.features.interactions import create_physics_interactions
    cols, keep = _cols(gold, "inputs")
    df = pd.DataFrame(cols)
    df["object_id"] = [f"o{i}" for i in range(len(df))]
    out = create_physics_interactions(df, keep)
    assert list(out.columns) == list(df.columns) + [str(x) for x in gold["full_names"]]
    assert out["object_id"].equals(df["object_id"]) and "gr_color_accel" not in df.columns          # input untouched
    np.testing.assert_allclose(out[[str(x) for x in gold["full_names"]]].to_numpy(), gold["full"], rtol=1e-14, equal_nan=True)
