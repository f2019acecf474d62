"""Developer script (CPU, needs /root/reference): fuzz the device feature templates (host instantiation) against the
UNMODIFIED reference on small, deliberately degenerate objects -- few points per band, repeated fluxes (ties in
argmax / median), quantised times (equal epochs across bands, zero gaps), zero and negative errors, all-negative or
all-zero bands.  Reports every column where NaN patterns differ or values differ by more than 1e-7 relative.

    python tests/fuzz_hostsim_vs_reference.py [objects] [seed]
"""
import contextlib, ctypes, io, os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, "/root/reference/src")
import numpy as np
import pandas as pd
from packref import pack_reference
from hostsim.build import build
from mallorn_astrophysics_b200 import schema

P = ctypes.c_void_p
ALL_COLUMNS = dict(schema.MODULE_COLUMNS, **schema.ORACLE_ONLY_COLUMNS)


def make_fuzz(n_obj, seed):
    rng = np.random.default_rng(seed)
    obj, band, t, f, e = [], [], [], [], []
    for k in range(n_obj):
        mode = rng.integers(0, 6)
        t0 = 60000.0 + rng.integers(0, 300)
        for b in range(6):
            n = int(rng.choice([0, 1, 2, 3, 4, 5, 6, 8, 12, 20], p=[.08, .06, .06, .1, .1, .1, .1, .15, .15, .1]))
            if n == 0:
                continue
            if mode in (0, 1):      # quantised epochs shared across bands (equal times, exact gaps)
                tt = t0 + rng.choice(np.arange(0, 400, 2.5 if mode == 0 else 0.5), size=n, replace=True)
            else:
                tt = t0 + rng.uniform(0, rng.choice([3.0, 40.0, 400.0]), n)
            if mode == 2:           # quantised fluxes: ties in argmax / median / percentiles
                ff = rng.integers(-3, 6, n).astype(float)
            elif mode == 3:
                ff = -np.abs(rng.normal(2, 1, n))
            elif mode == 4 and b == 2:
                ff = np.zeros(n)
            else:
                ff = rng.normal(3, 4, n) * (1 + 5 * np.exp(-0.5 * ((tt - t0 - 100) / 30) ** 2))
            ee = rng.uniform(0.2, 2.0, n)
            if mode == 5:
                ee[rng.random(n) < 0.3] = 0.0
                ee[rng.random(n) < 0.1] = -1.0
            obj += [k] * n; band += [b] * n; t += list(tt); f += list(ff); e += list(ee)
    obj = np.array(obj, np.int32); band = np.array(band, np.uint8)
    t, f, e = np.array(t), np.array(f), np.array(e)
    perm = rng.permutation(len(obj))                     # file order is not time order
    return obj[perm], band[perm], t[perm], f[perm], e[perm], rng.uniform(0.01, 1.0, n_obj)


def reference_frames(obj, band, t, f, e, z, n_obj):
    """Per-object calls of the reference's *_single functions, so that an object on which the reference itself raises
    (e.g. np.polyfit -> LinAlgError inside tde_physics.compute_decay_power_law) is reported instead of ending the run."""
    from features.statistical import extract_statistical_features
    from features.colors import extract_color_features_single
    from features.lightcurve_shape import extract_shape_features_single
    from features.tde_physics import extract_tde_physics_features_single
    from features.physics_based import extract_physics_features_single
    from features.research_features import extract_research_features_single
    names = np.array(list("ugrizy"))
    out = {m: (np.arange(n_obj), np.full((n_obj, len(ALL_COLUMNS[m])), np.nan)) for m in ("statistical", "colors", "shape", "tde", "physics", "research")}
    raised = {}
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for k in range(n_obj):
            mk = obj == k
            lc = pd.DataFrame({"object_id": f"o{k}", "Time (MJD)": t[mk], "Flux": f[mk], "Flux_err": e[mk], "Filter": names[band[mk]]})
            calls = {"statistical": lambda: extract_statistical_features(lc, [f"o{k}"]).iloc[0].to_dict(),
                     "colors": lambda: extract_color_features_single(lc), "shape": lambda: extract_shape_features_single(lc),
                     "tde": lambda: extract_tde_physics_features_single(lc), "physics": lambda: extract_physics_features_single(lc, z[k]),
                     "research": lambda: extract_research_features_single(lc, {"Z": z[k]})}
            for m, call in calls.items():
                try:
                    d = call()
                except Exception as exc:                      # noqa: BLE001 -- the point is to see what upstream does
                    raised.setdefault(m, []).append((k, type(exc).__name__))
                    continue
                for c_i, c in enumerate(ALL_COLUMNS[m]):
                    if c in d:
                        out[m][1][k, c_i] = float(d[c])
    out["_raised"] = raised
    return out


def main(n_obj=None, seed=None):
    if n_obj is None:
        n_obj = int(sys.argv[1]) if len(sys.argv) > 1 else 200
        seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    obj, band, t, f, e, z = make_fuzz(n_obj, seed)
    lib = ctypes.CDLL(build())
    pk = pack_reference(obj, band, t, f, e, n_obj)
    ref = reference_frames(obj, band, t, f, e, z, n_obj)
    fns = {"statistical": ("hs_stat", False), "colors": ("hs_color", True), "shape": ("hs_shape", True), "tde": ("hs_tde", True), "physics": ("hs_physics", True),
           "research": ("hs_research", True)}
    total_bad = 0
    total_defined = [0]
    # objects whose answer is not defined upstream: equal times inside a band (or anywhere, for the all-band columns)
    # are ordered by an unstable quicksort in the reference
    tie_band = np.zeros(n_obj, bool); tie_any = np.zeros(n_obj, bool)
    for k in range(n_obj):
        mk = obj == k
        tk, bk = t[mk], band[mk]
        tie_any[k] = len(np.unique(tk)) < len(tk)
        tie_band[k] = any(len(np.unique(tk[bk == b])) < (bk == b).sum() for b in range(6))
    print(f"{n_obj} objects: {int(tie_band.sum())} with equal times inside a band, {int(tie_any.sum())} with equal times anywhere")
    for m, (fn, takes_z) in fns.items():
        ncol = len(ALL_COLUMNS[m])
        out = np.zeros((n_obj, ncol))
        args = [ctypes.c_longlong(n_obj)] + [pk[k].ctypes.data_as(P) for k in ("seg_off", "t", "flux", "err", "rank")]
        if takes_z:
            args.append(np.ascontiguousarray(z).ctypes.data_as(P))
        getattr(lib, fn)(*args, out.ctypes.data_as(P))
        present, rows = ref[m]
        crashed = [k for k, _ in ref["_raised"].get(m, [])]
        if crashed:
            print(f"{m}: the REFERENCE raised on {len(crashed)} objects {ref['_raised'][m][:4]}; our rows there: "
                  f"{[('obj %d' % k, int(np.isnan(out[k]).sum()), 'NaN of', ncol) for k in crashed[:3]]}")
        present = np.array([k for k in present if k not in set(crashed)])
        got = out[present]; want = rows[present]
        nan_bad = np.isnan(got) != np.isnan(want)
        both = ~np.isnan(got) & ~np.isnan(want)
        rel = np.where(both, np.abs(got - want) / np.maximum(np.abs(want), 1e-9), 0.0)
        val_bad = both & (rel > 1e-7) & (np.abs(got - want) > 1e-8)     # slopes that are 0 up to rounding come out as ~1e-9 upstream
        bad = nan_bad | val_bad
        total_bad += int(bad.sum())
        print(f"{m}: {len(present)} objects, NaN-pattern mismatches {int(nan_bad.sum())}, value mismatches {int(val_bad.sum())}")
        cols = ALL_COLUMNS[m]
        seen = {}
        for i, j in np.argwhere(bad):
            seen.setdefault(cols[j], []).append((int(present[i]), got[i, j], want[i, j]))
        for c, items in list(seen.items())[:12]:
            defined = [it for it in items if not (tie_any[it[0]] if (c.startswith("all_") or m != "statistical") else tie_band[it[0]])]
            print(f"    {c}: {len(items)} objects ({len(defined)} without equal times)" +
                  (f", e.g. obj {defined[0][0]}: got {defined[0][1]!r} reference {defined[0][2]!r}" if defined else ""))
            total_defined[0] += len(defined)
    print("total mismatching cells:", total_bad, "of which on objects without equal times:", total_defined[0])
    return total_bad, total_defined[0], {m: len(v) for m, v in ref["_raised"].items()}


if __name__ == "__main__":
    main()
