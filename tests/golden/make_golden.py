#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs ``/root/reference``; the GPU box has no copy):

    python tests/golden/make_golden.py

For each seeded synthetic set it imports the reference's own extractors from
``/root/reference/src/features`` (and exec's ``scripts/train_v55_powerlaw.py:106-202`` for the
script-level decline-model fits), runs them on the long-format DataFrame, and stores
inputs + outputs as ``<name>.npz``.  ``multiband_gp`` cannot be run (it imports ``george``,
absent here) so it has no golden -- see ``oracle/gp2d.py``.  The per-band GP module
(``gaussian_process.py``, scikit-learn) runs and is recorded as ``gp1d_*``;
``python tests/golden/make_golden.py --add gp1d`` adds it to existing fixtures without re-running the rest.

Nothing is copied from the reference: only its outputs are recorded.
"""
import io
import contextlib
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(REF, "src"))

from mallorn_astrophysics_b200 import schema                      # noqa: E402
from mallorn_astrophysics_b200.synth import make_edge_table, make_table   # noqa: E402


def reference_powerlaw_namespace():
    """exec the self-contained function block of the v55 driver (its body needs Kaggle data)."""
    with open(os.path.join(REF, "scripts", "train_v55_powerlaw.py")) as fh:
        lines = fh.readlines()
    block = "".join(lines[105:202])          # lines 106-202: models, fit_decline_models, extract_powerlaw_features
    ns = {"np": np}
    exec(compile(block, "train_v55_powerlaw.py[106:202]", "exec"), ns)
    return ns


def run_reference(table):
    import pandas as pd
    from features.statistical import extract_statistical_features
    from features.colors import extract_color_features
    from features.lightcurve_shape import extract_shape_features
    from features.tde_physics import extract_tde_physics_features
    from features.physics_based import extract_physics_features
    from features.bazin_fitting import extract_bazin_features

    lc = table.to_frame()
    meta = table.metadata_frame()
    ids = list(table.object_ids)
    out = {}
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        frames = {
            "statistical": extract_statistical_features(lc, ids),
            "colors": extract_color_features(lc, ids),
            "shape": extract_shape_features(lc, ids),
            "tde": extract_tde_physics_features(lc, ids),
            "physics": extract_physics_features(lc, meta, ids),
            "bazin": extract_bazin_features(lc, ids),
        }
        ns = reference_powerlaw_namespace()
        present = [i for i in ids if (lc["object_id"] == i).any()]
        frames["powerlaw"] = pd.DataFrame([ns["extract_powerlaw_features"](i, lc) for i in present])
    id_to_idx = {s: k for k, s in enumerate(ids)}
    for name, df in frames.items():
        cols = [c for c in df.columns if c != "object_id"]
        want = schema.MODULE_COLUMNS[name]
        # Column *sets* must match exactly.  The order matches too except for two reference
        # quirks that depend on which object comes first (tde_physics.py:299-303 inserts
        # alpha_late before residual on its short-band path; physics_based.py:128-129 omits
        # r_sf_slope for a 3-4 point r band) -- rows are stored in schema order.
        assert sorted(cols) == sorted(want), (name, set(cols) ^ set(want))
        if name not in ("tde", "physics"):
            assert cols == want, name
        cols = want
        if name == "statistical":
            assert df.columns[0] == "object_id"
        else:
            assert df.columns[-1] == "object_id" or name == "powerlaw"
        out[f"{name}_rows"] = df[cols].to_numpy(dtype=np.float64)
        out[f"{name}_present"] = np.array([id_to_idx[s] for s in df["object_id"]], dtype=np.int64)
    return out


def run_reference_gp1d(table):
    """src/features/gaussian_process.py (sklearn GaussianProcessRegressor) -> gp1d_rows / gp1d_present."""
    from features.gaussian_process import extract_gp_features
    lc, meta, ids = table.to_frame(), table.metadata_frame(), list(table.object_ids)
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        df = extract_gp_features(lc, meta, ids, verbose=False)
    cols = [c for c in df.columns if c != "object_id"]
    assert cols == schema.MODULE_COLUMNS["gp1d"] and df.columns[-1] == "object_id", cols
    id_to_idx = {s: k for k, s in enumerate(ids)}
    return {"gp1d_rows": df[cols].to_numpy(dtype=np.float64),
            "gp1d_present": np.array([id_to_idx[s] for s in df["object_id"]], dtype=np.int64)}


def run_reference_research(table):
    """src/features/research_features.py (v115c block) -> research_rows / research_present."""
    from features.research_features import extract_research_features
    lc, meta, ids = table.to_frame(), table.metadata_frame(), list(table.object_ids)
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        df = extract_research_features(lc, ids, meta, verbose=False)
    cols = [c for c in df.columns if c != "object_id"]
    assert cols == schema.RESEARCH_COLUMNS and df.columns[-1] == "object_id", cols
    id_to_idx = {s: k for k, s in enumerate(ids)}
    return {"research_rows": df[cols].to_numpy(dtype=np.float64),
            "research_present": np.array([id_to_idx[s] for s in df["object_id"]], dtype=np.int64)}


def add_module(name, table, runner, label):
    path = os.path.join(HERE, f"{name}.npz")
    old = dict(np.load(path))
    assert np.array_equal(old["t"], table.t) and np.array_equal(old["flux"], table.flux) and np.array_equal(old["obj"], table.obj)
    old.update(runner(table))
    np.savez_compressed(path, **old)
    print(name, label, "added:", os.path.getsize(path) // 1024, "KiB")


def add_gp1d(name, table):
    """Append the gp1d golden to an existing fixture; the stored inputs must be the table's."""
    path = os.path.join(HERE, f"{name}.npz")
    old = dict(np.load(path))
    assert np.array_equal(old["t"], table.t) and np.array_equal(old["flux"], table.flux) and np.array_equal(old["obj"], table.obj)
    old.update(run_reference_gp1d(table))
    np.savez_compressed(path, **old)
    print(name, "gp1d added:", old["gp1d_rows"].shape, os.path.getsize(path) // 1024, "KiB")


def crafted_gp1d_cases():
    """Ten hand-made objects for the per-band GP: constant flux, one epoch only, NaN rows leaving < 5 / >= 5 valid
    points, zero and negative errors, duplicated times, exactly five points, huge and tiny errors, two clusters."""
    rng = np.random.default_rng(1)
    rows = []

    def add(o, band, t, f, e):
        rows.extend(zip([o] * len(t), [band] * len(t), t, f, e))
    t = np.sort(60000 + rng.uniform(0, 500, 12))
    add(0, 1, t, np.full(12, 3.0), np.full(12, 0.5))
    add(1, 2, np.full(8, 60010.0), rng.normal(0, 1, 8), np.full(8, 0.3))
    f = rng.normal(5, 2, 8); f[:4] = np.nan; add(2, 3, t[:8], f, np.full(8, 0.4))
    f = rng.normal(5, 2, 12); f[[1, 5]] = np.nan; add(3, 1, t, f, np.full(12, 0.4))
    e = np.full(12, 0.5); e[[0, 3]] = 0.0; e[7] = -1.0; add(4, 2, t, rng.normal(2, 3, 12), e)
    tt = t.copy(); tt[3] = tt[2]; tt[9] = tt[8]; add(5, 3, tt, rng.normal(2, 3, 12), np.full(12, 0.6))
    add(6, 4, t[:5], rng.normal(1, 1, 5), np.full(5, 0.2))
    add(7, 1, t, rng.normal(0, 1, 12), np.full(12, 50.0))
    add(8, 2, t, np.sin(t / 50) * 5, np.full(12, 1e-7))
    tc = np.sort(np.concatenate([60000 + rng.uniform(0, 5, 10), 60400 + rng.uniform(0, 5, 10)]))
    add(9, 3, tc, rng.normal(0, 2, 20), np.full(20, 0.5))
    a = np.array(rows)
    return (a[:, 0].astype(np.int32), a[:, 1].astype(np.uint8), a[:, 2].copy(), a[:, 3].copy(), a[:, 4].copy(), 10)


def save_gp1d_crafted():
    """gp1d_crafted.npz: the reference's extract_gp_features_single on the crafted objects (row per object)."""
    import pandas as pd
    from features.gaussian_process import extract_gp_features_single
    obj, band, t, f, e, n_obj = crafted_gp1d_cases()
    names = np.array(list("ugrizy"))
    out = np.full((n_obj, 21), np.nan)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for k in range(n_obj):
            m = obj == k
            df = pd.DataFrame({"Time (MJD)": t[m], "Flux": f[m], "Flux_err": e[m], "Filter": names[band[m]]})
            r = extract_gp_features_single(df)
            out[k] = [r[c] for c in schema.MODULE_COLUMNS["gp1d"]]
    path = os.path.join(HERE, "gp1d_crafted.npz")
    np.savez_compressed(path, obj=obj, band=band, t=t, flux=f, err=e, n_obj=np.int64(n_obj), gp1d_rows=out)
    print("gp1d_crafted", out.shape, "->", path)


def save(name, table):
    ref = run_reference(table)
    ref.update(run_reference_gp1d(table))
    ref.update(run_reference_research(table))
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, obj=table.obj, band=table.band, t=table.t, flux=table.flux, err=table.err,
                        n_obj=np.int64(table.n_obj), z=table.z, ebv=table.ebv, **ref)
    print(name, table.n_obj, "objects", table.n_rows, "rows ->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    tables = [("ref_sparse", lambda: make_table(48, config_index=0)),
              ("ref_dense", lambda: make_table(24, seed=424242, count_scale=2.5)),
              ("ref_edge", make_edge_table),
              ("ref_c1_256", lambda: make_table(256, config_index=1))]   # a slice of the C1-shaped set: more objects, same generator
    only_gp1d = sys.argv[1:3] == ["--add", "gp1d"]
    if sys.argv[1:3] == ["--add", "gp1d_crafted"]:
        save_gp1d_crafted()
        sys.exit(0)
    if sys.argv[1:3] == ["--add", "research"]:
        for name, make in tables:
            add_module(name, make(), run_reference_research, "research")
        sys.exit(0)
    for name, make in tables:
        (add_gp1d if only_gp1d else save)(name, make())
    save_gp1d_crafted()
