"""Downstream acceptance check: does the v34a/v92d-style GBDT score the same
out-of-fold F1 on the GPU features as on the CPU (oracle) features?

Test infrastructure.  What is reproduced from the reference's drivers, and how:

* feature assembly -- ``scripts/train_v34a_bazin.py:55-123``: base table = statistical (+ ``add_metadata_features``) +
  colors + shape + physics (``train_v4_physics.py:53-109``), ``selected_120`` = the 120 most important base features after
  dropping the second member of every |corr| >= 0.95 pair, then + 25 tde + 27 gp2d + 52 bazin columns; the v92d list is
  the v34a list minus the two "shift" features (``train_v92_focal_adversarial.py:87-97``) -> 222 columns.
* the artefacts the drivers read from disk (``selected_features.pkl``: ``importance_df``, ``high_corr_df``;
  ``v34a_artifacts.pkl``: ``feature_names``) do not exist offline, so deterministic SURROGATES are built from the CPU
  arm's base table and fed identically to both arms -- exactly as the real files would be: importance from a gradient
  boosted model (sklearn; xgboost is not installed), correlation pairs from the CPU table.  The real v34a list is known to
  contain ``all_rise_time`` / ``all_asymmetry`` (v92d drops them), so the surrogate list keeps them in: 118 + 2 = 120.
* classifier -- 5-fold ``StratifiedKFold(shuffle=True, random_state=seed)``, early stopping on the validation fold with
  patience 50 out of 500 rounds at learning rate 0.025, depth 5, column sub-sampling 0.8, L2 1.5, positive-class weight
  n_neg / n_pos (``train_v34a_bazin.py:131-170``), best-F1 threshold over ``linspace(0.05, 0.5, 100)`` on the out-of-fold
  probabilities (:199-206).  Stand-in for xgboost: sklearn ``HistGradientBoostingClassifier`` (same histogram-tree
  family); row sub-sampling 0.8 has no equivalent there.
"""
from __future__ import annotations

import os
import warnings

import numpy as np

BASE_MODULES = ("statistical", "colors", "shape", "physics")
EXTRA_MODULES = ("tde", "gp2d", "bazin")
ALL = BASE_MODULES + EXTRA_MODULES
SHIFT_FEATURES = ["all_rise_time", "all_asymmetry"]            # train_v92_focal_adversarial.py:95


def c3_train_table():
    """The labelled part of the C3 set: 10,178 objects (seed 20260118 + 2), the first 3,054 are 'train' (SURVEY 8d)."""
    from mallorn_astrophysics_b200.synth import make_table
    return make_table(10178, config_index=2).subset(3054)


# ---------------------------------------------------------------------------------------------- feature arms
def _oracle_chunk(args):
    from oracle.extract import run_module
    cols, lo, hi, modules = args
    keep = (cols["obj"] >= lo) & (cols["obj"] < hi)
    obj = (cols["obj"][keep] - lo).astype(np.int32)
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for m in modules:
            present, rows = run_module(m, obj, cols["band"][keep], cols["t"][keep], cols["flux"][keep], cols["err"][keep],
                                       hi - lo, cols["z"][lo:hi])
            full = np.full((hi - lo, rows.shape[1]), np.nan)
            full[present] = rows
            out[m] = full
    return lo, out


def oracle_rows(table, modules=ALL, pool=None, cores=None):
    """{module: [n_obj, cols]} from the CPU oracle (the reference's arithmetic), object chunks over worker processes."""
    cols = {"obj": table.obj, "band": table.band, "t": table.t, "flux": table.flux, "err": table.err, "z": table.z}
    cores = cores or os.cpu_count() or 1
    per = max(1, table.n_obj // (cores * 6))
    jobs = [(cols, lo, min(lo + per, table.n_obj), tuple(modules)) for lo in range(0, table.n_obj, per)]
    parts = dict((pool.map(_oracle_chunk, jobs) if pool is not None else map(_oracle_chunk, jobs)))
    return {m: np.concatenate([parts[lo][m] for lo in sorted(parts)], axis=0) for m in modules}


def _libm_bazin_chunk(args):
    """Bazin rows of the oracle with glibc's exp (math.exp) in place of numpy's SIMD exp: the two differ by at most one unit
    in the last place on a few per cent of arguments.  Both are "the reference"; the arm built from this one is the noise
    floor of the downstream check (how far F1 moves under the reference's own libm-level irreproducibility)."""
    import math
    from oracle import bazin as ob
    from oracle.common import group_objects
    cols, lo, hi = args

    def _exp1(v):
        try:
            return math.exp(v)
        except OverflowError:
            return math.inf
    vexp = np.frompyfunc(_exp1, 1, 1)
    exp_libm = lambda x: vexp(np.asarray(x, dtype=np.float64)).astype(np.float64)
    orig = ob.bazin
    ob.bazin = lambda t, A, t0, tau_rise, tau_fall, B: A * exp_libm(-(t - t0) / tau_fall) / (1.0 + exp_libm(-(t - t0) / tau_rise)) + B
    keep = (cols["obj"] >= lo) & (cols["obj"] < hi)
    obj = (cols["obj"][keep] - lo).astype(np.int32)
    full = np.full((hi - lo, 52), np.nan)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for k, v in group_objects(obj, cols["band"][keep], cols["t"][keep], cols["flux"][keep], cols["err"][keep], hi - lo):
                full[k] = ob.object_row(v)
    finally:
        ob.bazin = orig
    return lo, full


def libm_bazin_rows(table, pool=None, cores=None):
    cols = {"obj": table.obj, "band": table.band, "t": table.t, "flux": table.flux, "err": table.err}
    cores = cores or os.cpu_count() or 1
    per = max(1, table.n_obj // (cores * 6))
    jobs = [(cols, lo, min(lo + per, table.n_obj)) for lo in range(0, table.n_obj, per)]
    parts = dict((pool.map(_libm_bazin_chunk, jobs) if pool is not None else map(_libm_bazin_chunk, jobs)))
    return np.concatenate([parts[lo] for lo in sorted(parts)], axis=0)


def gpu_rows(table, modules=ALL, device=0):
    from mallorn_astrophysics_b200.store import LightCurveStore
    with LightCurveStore.from_table(table, device=device) as s:
        return {m: s.features(m, z=table.z if m == "physics" else None) for m in modules}


def hostsim_rows(table, modules=ALL):
    """CPU-only stand-in for the GPU arm while developing without a GPU: the host instantiation of the device templates."""
    import ctypes
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    from hostsim.build import build
    from packref import pack_reference
    lib = ctypes.CDLL(build())
    P = ctypes.c_void_p
    pk = pack_reference(table.obj, table.band, table.t, table.flux, table.err, table.n_obj)
    fn = {"statistical": ("hs_stat", 123, False, 0), "colors": ("hs_color", 83, True, 0), "shape": ("hs_shape", 65, True, 0),
          "tde": ("hs_tde", 25, True, 0), "physics": ("hs_physics", 32, True, 0), "bazin": ("hs_bazin_t", 52, False, 12),
          "powerlaw": ("hs_powerlaw_t", 27, False, 54), "gp2d": ("hs_gp2d", 27, False, 4)}
    out = {}
    for m in modules:
        name, nc, takes_z, ns = fn[m]
        rows = np.zeros((table.n_obj, nc))
        st = np.zeros((table.n_obj, max(ns, 1)), np.int32)
        args = [ctypes.c_longlong(table.n_obj)] + [pk[k].ctypes.data_as(P) for k in ("seg_off", "t", "flux", "err", "rank")]
        if takes_z:
            args.append(np.ascontiguousarray(table.z).ctypes.data_as(P))
        args.append(rows.ctypes.data_as(P))
        if ns:
            args.append(st.ctypes.data_as(P))
        getattr(lib, name)(*args)
        out[m] = rows
    return out


# ---------------------------------------------------------------------------------------------- assembly
def assemble(rows: dict, table):
    """The frames the reference's drivers read from their caches, from per-module rows: ``base`` = statistical
    (+ ``add_metadata_features``, statistical.py:229) merged with colors, shape, physics (train_v4_physics.py:63-79), and
    the tde / gp2d / bazin frames; one row per object, reference column names, ``object_id`` included."""
    import pandas as pd
    from mallorn_astrophysics_b200 import schema
    ids = table.object_ids
    frames = {}
    for m in ALL:
        f = pd.DataFrame(rows[m], columns=schema.MODULE_COLUMNS[m])
        f.insert(0, "object_id", ids)
        frames[m] = f
    meta = table.metadata_frame()
    base = frames["statistical"].merge(meta[["object_id", "Z", "EBV"]], on="object_id", how="left")
    base["luminosity_distance"] = base["Z"] * 4280
    base["time_dilation"] = 1 + base["Z"]
    for m in ("colors", "shape", "physics"):
        base = base.merge(frames[m], on="object_id", how="left")
    return base, {m: frames[m] for m in EXTRA_MODULES}


def surrogate_artifacts(frame, base_cols, y):
    """``importance_df`` / ``high_corr_df`` of selected_features.pkl, rebuilt deterministically from the CPU arm."""
    import pandas as pd
    from sklearn.ensemble import GradientBoostingClassifier
    X = frame[base_cols].to_numpy(dtype=np.float64)
    X = np.where(np.isfinite(X), X, np.nan)
    med = np.nanmedian(X, axis=0)
    med = np.where(np.isfinite(med), med, 0.0)
    Xi = np.where(np.isnan(X), med, X)
    gb = GradientBoostingClassifier(n_estimators=150, max_depth=3, subsample=0.8, learning_rate=0.05, random_state=42)
    gb.fit(Xi, y)
    importance_df = pd.DataFrame({"feature": base_cols, "importance": gb.feature_importances_}) \
        .sort_values("importance", ascending=False, kind="stable").reset_index(drop=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        corr = np.corrcoef(Xi, rowvar=False)
    pairs = [(base_cols[i], base_cols[j], corr[i, j]) for i in range(len(base_cols)) for j in range(i + 1, len(base_cols))
             if np.isfinite(corr[i, j]) and abs(corr[i, j]) >= 0.95]
    high_corr_df = pd.DataFrame(pairs, columns=["feature_1", "feature_2", "correlation"])
    return importance_df, high_corr_df


def selected_120(importance_df, high_corr_df):
    """train_v34a_bazin.py:62-67 with the surrogate artefacts (the two shift features kept in, see the module docstring)."""
    corr_to_drop = set()
    for _, row in high_corr_df.iterrows():
        if row["feature_1"] not in corr_to_drop:
            corr_to_drop.add(row["feature_2"])
    corr_to_drop -= set(SHIFT_FEATURES)
    clean = importance_df[~importance_df["feature"].isin(corr_to_drop)]
    ranked = [f for f in clean["feature"].tolist() if f not in SHIFT_FEATURES]
    return ranked[:118] + SHIFT_FEATURES


def v34a_frame(base, extra, selected):
    """train_v34a_bazin.py:84-86,118: selected_120 -> + tde -> + gp2d -> + bazin, pandas' default merge suffixes."""
    f = base[["object_id"] + selected].copy()
    f = f.merge(extra["tde"], on="object_id", how="left")
    f = f.merge(extra["gp2d"], on="object_id", how="left")
    return f.merge(extra["bazin"], on="object_id", how="left")


def v92d_matrix(frame):
    """The 222 columns: the v34a list without the shift features, infinities clipped (train_v92_focal_adversarial.py:95-103)."""
    names = [c for c in frame.columns if c != "object_id" and c not in SHIFT_FEATURES]
    X = frame[names].to_numpy(dtype=np.float64)
    return np.nan_to_num(X, nan=np.nan, posinf=1e10, neginf=-1e10), names


# ---------------------------------------------------------------------------------------------- classifier
def _fit_fold(args):
    from sklearn.ensemble import HistGradientBoostingClassifier
    X, y, tr, va, seed = args
    pos = float((y[tr] == 1).sum())
    w = np.where(y[tr] == 1, (len(tr) - pos) / max(pos, 1.0), 1.0)                     # scale_pos_weight
    clf = HistGradientBoostingClassifier(max_iter=500, learning_rate=0.025, max_depth=5, l2_regularization=1.5,
                                         max_features=0.8, min_samples_leaf=3, early_stopping=False, random_state=seed)
    clf.fit(X[tr], y[tr], sample_weight=w)
    # early stopping on the validation fold, patience 50 (xgb.train(..., early_stopping_rounds=50) predicts at the best round)
    best_loss, best_raw, since = np.inf, None, 0
    yv = y[va]
    for raw in clf.staged_decision_function(X[va]):
        p = 1.0 / (1.0 + np.exp(-raw))
        loss = -np.mean(yv * np.log(np.clip(p, 1e-15, 1)) + (1 - yv) * np.log(np.clip(1 - p, 1e-15, 1)))
        if loss < best_loss:
            best_loss, best_raw, since = loss, raw, 0
        else:
            since += 1
            if since >= 50:
                break
    return va, 1.0 / (1.0 + np.exp(-best_raw))


def oof_f1(X, y, seed, pool=None):
    from sklearn.model_selection import StratifiedKFold
    folds = list(StratifiedKFold(5, shuffle=True, random_state=seed).split(X, y))
    jobs = [(X, y, tr, va, seed) for tr, va in folds]
    oof = np.zeros(len(y))
    for va, p in (pool.map(_fit_fold, jobs) if pool is not None else map(_fit_fold, jobs)):
        oof[va] = p
    best, best_t = 0.0, 0.5
    for thr in np.linspace(0.05, 0.5, 100):
        pred = oof > thr
        tp = np.sum(pred & (y == 1)); fp = np.sum(pred & (y == 0)); fn = np.sum(~pred & (y == 1))
        f1 = 2 * tp / max(2 * tp + fp + fn, 1)
        if f1 > best:
            best, best_t = f1, thr
    return float(best), float(best_t), oof


def oof_f1_k10(X, y, seed, weight=None):
    """The same protocol with the product's own trainer (K10, csrc/gbdt.cu: xgboost-hist restated with the v92d parameters,
    row and column sub-sampling, L1/L2 regularisation, scale_pos_weight, optional adversarial weights) on the GPU."""
    import torch
    from mallorn_astrophysics_b200 import gbdt
    table = torch.from_numpy(np.ascontiguousarray(X)).cuda()
    res = gbdt.cross_validate(table, y, weight, gbdt.GBDTParams(seed=seed), 5, 500, 50, use_scale_pos_weight=True, fold_seed=seed)
    f1, thr = gbdt.best_f1(y, res["oof"], np.linspace(0.05, 0.5, 100))
    return f1, thr, res["oof"]


def compare_arms(table, rows_cpu, rows_gpu, seeds=(42, 43, 44, 45, 46), pool=None, classifier="sklearn", weight=None):
    """The C3 acceptance numbers: OOF F1 of both arms per seed on the 222-column v92d table and on the v34a table."""
    y = (table.metadata_frame()["target"].to_numpy()).astype(int)
    base_c, extra_c = assemble(rows_cpu, table)
    base_g, extra_g = assemble(rows_gpu, table)
    base_cols = [c for c in base_c.columns if c != "object_id"]
    importance_df, high_corr_df = surrogate_artifacts(base_c, base_cols, y)
    sel = selected_120(importance_df, high_corr_df)
    Xc, names = v92d_matrix(v34a_frame(base_c, extra_c, sel))
    Xg, names_g = v92d_matrix(v34a_frame(base_g, extra_g, sel))
    assert names == names_g
    res = {"objects": int(table.n_obj), "positives": int(y.sum()), "columns": len(names), "base_columns": len(base_cols),
           "high_corr_pairs": int(len(high_corr_df)), "seeds": {}}
    for seed in seeds:
        if classifier == "k10":
            fc, tc, _ = oof_f1_k10(Xc, y, seed, weight)
            fg, tg, _ = oof_f1_k10(Xg, y, seed, weight)
        else:
            fc, tc, _ = oof_f1(Xc, y, seed, pool)
            fg, tg, _ = oof_f1(Xg, y, seed, pool)
        res["seeds"][int(seed)] = {"f1_cpu": fc, "f1_gpu": fg, "delta": fg - fc, "threshold_cpu": tc, "threshold_gpu": tg}
    d = np.array([v["delta"] for v in res["seeds"].values()])
    res["mean_delta"] = float(d.mean())
    res["mean_abs_delta"] = float(np.abs(d).mean())
    res["max_abs_delta"] = float(np.abs(d).max())
    res["std_delta"] = float(d.std(ddof=1)) if len(d) > 1 else 0.0
    res["f1_cpu_seed_std"] = float(np.std([v["f1_cpu"] for v in res["seeds"].values()], ddof=1)) if len(d) > 1 else 0.0
    differing = int(np.sum(~np.isclose(Xc, Xg, rtol=1e-3, atol=1e-9, equal_nan=True)))
    res["cells_differing_beyond_1e-3"] = differing
    res["cells"] = int(Xc.size)
    res["classifier"] = ("K10: this library's histogram GBDT on the GPU (csrc/gbdt.cu), v92d parameters and protocol" if classifier == "k10" else
                         "sklearn HistGradientBoostingClassifier, v34a protocol (xgboost is not installed in this image)")
    res["feature_names"] = names
    return res
