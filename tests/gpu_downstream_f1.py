"""Downstream check: does a gradient-boosted classifier trained on the GPU
features score the same out-of-fold F1 as one trained on the CPU (oracle) features?

The reference's v34a/v92d drivers use xgboost (absent from this image); the stand-in is scikit-learn's
HistGradientBoostingClassifier with the same protocol -- 5-fold stratified CV, TDE-vs-rest labels, best-F1
threshold on the out-of-fold probabilities -- run on the merged 434-column table of a labelled synthetic set.
North star: |F1_gpu - F1_cpu| <= 0.005.   Usage (GPU box):  python tests/gpu_downstream_f1.py [n_objects]
"""
import os
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[_v] = "1"          # one thread per worker process: the parallelism is across processes
import json
import multiprocessing as mp
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
MODULES = ("statistical", "colors", "shape", "tde", "physics", "bazin", "powerlaw", "gp2d")


def _cpu_chunk(args):
    from oracle.extract import run_module
    cols, lo, hi = args
    keep = (cols["obj"] >= lo) & (cols["obj"] < hi)
    obj = (cols["obj"][keep] - lo).astype(np.int32)
    out = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for m in MODULES:
            present, rows = run_module(m, obj, cols["band"][keep], cols["t"][keep], cols["flux"][keep], cols["err"][keep],
                                       hi - lo, cols["z"][lo:hi])
            full = np.full((hi - lo, rows.shape[1]), np.nan)
            full[present] = rows
            out.append(full)
    return lo, np.concatenate(out, axis=1)


def _fit_fold(args):
    from sklearn.ensemble import HistGradientBoostingClassifier
    X, y, tr, va, seed = args
    clf = HistGradientBoostingClassifier(max_iter=60, learning_rate=0.1, max_depth=5, random_state=seed,
                                         class_weight="balanced")
    clf.fit(X[tr], y[tr])
    return va, clf.predict_proba(X[va])[:, 1]


def oof_f1(pool, X, y, seed=0):
    from sklearn.model_selection import StratifiedKFold
    folds = list(StratifiedKFold(5, shuffle=True, random_state=seed).split(X, y))
    oof = np.zeros(len(y))
    for va, p in pool.map(_fit_fold, [(X, y, tr, va, seed) for tr, va in folds]):
        oof[va] = p
    best = 0.0
    for thr in np.linspace(0.05, 0.95, 91):
        pred = oof >= thr
        tp = np.sum(pred & (y == 1)); fp = np.sum(pred & (y == 0)); fn = np.sum(~pred & (y == 1))
        best = max(best, 2 * tp / max(2 * tp + fp + fn, 1))
    return best


def main():
    from mallorn_astrophysics_b200.store import LightCurveStore
    from mallorn_astrophysics_b200.synth import CLASS_TDE, make_table
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    table = make_table(n, config_index=2)
    y = (table.label == CLASS_TDE).astype(int)
    # CPU oracle first (worker processes are forked before this process touches CUDA)
    cols = {"obj": table.obj, "band": table.band, "t": table.t, "flux": table.flux, "err": table.err, "z": table.z}
    cores = os.cpu_count() or 1
    per = max(1, n // (cores * 4))
    jobs = [(cols, lo, min(lo + per, n)) for lo in range(0, n, per)]
    t0 = time.time()
    pool = mp.get_context("fork").Pool(cores)          # forked before this process touches CUDA
    parts = dict(pool.map(_cpu_chunk, jobs))
    Xc = np.concatenate([parts[lo] for lo in sorted(parts)], axis=0)
    t_cpu = time.time() - t0
    print(f"cpu oracle: {t_cpu:.1f} s on {cores} processes", file=sys.stderr, flush=True)
    t0 = time.time()
    with LightCurveStore.from_table(table) as s:
        Xg = np.concatenate([s.features(m, z=table.z if m == "physics" else None) for m in MODULES], axis=1)
    t_gpu = time.time() - t0
    print(f"gpu: {t_gpu:.1f} s", file=sys.stderr, flush=True)
    # the GBDT stand-in cannot bin columns that are entirely NaN / constant, nor infinities: clean BOTH tables the same way
    Xg = np.where(np.isfinite(Xg), Xg, np.nan); Xc = np.where(np.isfinite(Xc), Xc, np.nan)
    usable = np.array([min((~np.isnan(Xc[:, j])).sum(), (~np.isnan(Xg[:, j])).sum()) >= 0.05 * n
                       and (np.unique(Xc[~np.isnan(Xc[:, j]), j]).size > 1) and (np.unique(Xg[~np.isnan(Xg[:, j]), j]).size > 1)
                       for j in range(Xc.shape[1])])
    Xg, Xc = Xg[:, usable], Xc[:, usable]
    res = {"objects": n, "positives": int(y.sum()), "columns_used": int(Xg.shape[1]), "gpu_seconds": t_gpu, "cpu_seconds": t_cpu,
           "cpu_cores": cores, "seeds": {}}
    for seed in (0, 1):
        t0 = time.time()
        fg, fc = oof_f1(pool, Xg, y, seed), oof_f1(pool, Xc, y, seed)
        print(f"seed {seed}: f1 gpu {fg:.4f} cpu {fc:.4f} ({time.time() - t0:.1f} s)", file=sys.stderr, flush=True)
        res["seeds"][seed] = {"f1_gpu": fg, "f1_cpu": fc, "delta": fg - fc}
    d = [v["delta"] for v in res["seeds"].values()]
    res["mean_abs_delta"] = float(np.mean(np.abs(d)))
    res["max_abs_delta"] = float(np.max(np.abs(d)))
    res["classifier"] = "sklearn HistGradientBoostingClassifier (xgboost is not installed in this image)"
    pool.close()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
