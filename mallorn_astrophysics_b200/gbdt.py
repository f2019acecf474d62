"""Histogram gradient-boosted trees trained on the device-resident feature table.

Host side of ``csrc/gbdt.cu`` (C-ABI ``mlrn_gbdt_*``): the downstream classifier of the reference's drivers --
XGBoost ``tree_method='hist'``, binary:logistic, with the v92d parameters, adversarial sample weights and 5-fold
early-stopped cross-validation (``non_successful_tests/scripts/train_v92_focal_adversarial.py:157-169,232-272``,
``adversarial_validation.py:120-160,236-246``) -- fed straight from the table the feature passes leave in HBM
(``shard.featurise_shard``): no host round trip of the features.  Only the cut points are computed on the host, from a
bounded sample of rows.  xgboost is not installed here, so agreement with it is not pinned; the CPU restatement of the same
algorithm (``oracle/gbdt.py``) grows bit-identical trees (``tests/test_gbdt.py``).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _native

MISSING = 255
MASK64 = (1 << 64) - 1


@dataclass
class GBDTParams:
    """train_v92_focal_adversarial.py:157-169 (``base_params``)."""
    max_depth: int = 5
    learning_rate: float = 0.025
    subsample: float = 0.8
    colsample_bytree: float = 0.8
    min_child_weight: float = 3.0
    reg_alpha: float = 0.2
    reg_lambda: float = 1.5
    scale_pos_weight: float = 1.0
    seed: int = 42


class _CParams(C.Structure):
    _fields_ = [("max_depth", C.c_int32), ("learning_rate", C.c_double), ("subsample", C.c_double), ("min_child_weight", C.c_double),
                ("reg_alpha", C.c_double), ("reg_lambda", C.c_double), ("scale_pos_weight", C.c_double), ("seed", C.c_uint64)]


class _CTree(C.Structure):
    _fields_ = [("feature", C.c_int32 * 63), ("bin", C.c_int32 * 63), ("default_left", C.c_int32 * 63), ("value", C.c_double * 63)]


def _bind():
    L = _native.lib()
    if not getattr(L, "_gbdt_bound", False):
        vp = C.c_void_p
        L.mlrn_gbdt_bin.argtypes = [C.c_int, vp, C.c_int64, C.c_int32, C.c_int32, vp, vp, vp, vp]
        L.mlrn_gbdt_create.argtypes = [C.c_int, vp, C.c_int64, C.c_int32, vp, vp, C.POINTER(_CParams), C.c_int32, vp, C.POINTER(vp)]
        L.mlrn_gbdt_destroy.argtypes = [vp]
        L.mlrn_gbdt_round.argtypes = [vp, vp, vp]
        L.mlrn_gbdt_predict.argtypes = [vp, vp, C.c_int64, C.c_int32, C.c_int32, vp, vp]
        L.mlrn_gbdt_export.argtypes = [vp, vp, vp]
        L.mlrn_gbdt_rounds.argtypes = [vp]
        L._gbdt_bound = True
    return L


def _hash(seed: int, a: int, b: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        b = np.asarray(b, dtype=np.uint64)
        z = np.uint64(seed & MASK64) + np.uint64(0x9E3779B97F4A7C15) * np.uint64(a + 1) + np.uint64(0xBF58476D1CE4E5B9) * (b + np.uint64(1))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def column_mask(seed: int, rnd: int, F: int, frac: float) -> np.ndarray:
    """colsample_bytree: the ``max(1, floor(frac * F))`` columns with the smallest hash of (seed, round, column)."""
    k = max(1, int(np.floor(frac * F)))
    if k >= F:
        return np.ones(F, np.uint8)
    h = _hash(seed ^ 0xC01C01C01, rnd, np.arange(F))
    order = sorted(range(F), key=lambda j: (int(h[j]), j))
    m = np.zeros(F, np.uint8)
    m[order[:k]] = 1
    return m


def quantile_cuts(X: np.ndarray, max_cuts: int = 254):
    """Cut points per feature from a HOST sample ``X`` [rows, F]: midpoints between distinct values, thinned to at most
    ``max_cuts`` evenly spaced ones (xgboost's ``max_bin`` = 256 sketch plays this role upstream)."""
    F = X.shape[1]
    cuts = np.full((F, 255), np.inf)
    ncuts = np.zeros(F, np.int32)
    for f in range(F):
        col = X[:, f]
        v = np.unique(col[~np.isnan(col)])
        if v.size <= 1:
            continue
        mids = 0.5 * (v[:-1] + v[1:])
        if mids.size > max_cuts:
            mids = mids[np.unique(np.round(np.linspace(0, mids.size - 1, max_cuts)).astype(int))]
        cuts[f, :mids.size] = mids
        ncuts[f] = mids.size
    return cuts, ncuts


class Binner:
    """Feature quantisation shared by every fold / model trained on one table."""

    def __init__(self, table, sample_rows: int = 65536, stream: int = 0):
        import torch
        n = table.shape[0]
        step = max(1, n // sample_rows)
        sample = table[::step].cpu().numpy()                        # the only host copy: a bounded sample for the cut points
        sample = np.where(np.isfinite(sample), sample, np.where(np.isnan(sample), np.nan, np.sign(sample) * 1e10))
        self.cuts, self.ncuts = quantile_cuts(sample)
        self.device = table.device
        self.d_cuts = torch.from_numpy(self.cuts).to(self.device)
        self.d_ncuts = torch.from_numpy(self.ncuts).to(self.device)
        self.stream = stream

    def bins(self, table):
        """[n, F] uint8 CUDA tensor; +-inf are clipped to +-1e10 as train_v92_focal_adversarial.py:102-103 does."""
        import torch
        L = _bind()
        X = torch.nan_to_num(table, nan=float("nan"), posinf=1e10, neginf=-1e10).contiguous()
        out = torch.empty(X.shape, dtype=torch.uint8, device=X.device)
        _native.check(L.mlrn_gbdt_bin(X.device.index, X.data_ptr(), X.shape[0], X.shape[1], X.shape[1], self.d_cuts.data_ptr(),
                                      self.d_ncuts.data_ptr(), out.data_ptr(), self.stream), "mlrn_gbdt_bin")
        return out


class HistGBDT:
    """One boosted model on device-resident binned features."""

    def __init__(self, params: GBDTParams = GBDTParams()):
        self.p = params
        self._h = C.c_void_p()
        self._keep = None
        self.best_round = 0

    def fit(self, bins, y, weight=None, eval_bins=None, eval_y=None, rounds: int = 500, early_stopping_rounds: int = 50, stream: int = 0):
        """``bins`` [n, F] uint8 CUDA, ``y`` [n] float64 CUDA (0/1), ``weight`` [n] float64 CUDA or None.  With an eval set:
        stop when its log-loss has not improved for ``early_stopping_rounds`` rounds; predictions then use the best round,
        as ``xgb.train(..., early_stopping_rounds=50)`` + ``predict`` does (train_v92_focal_adversarial.py:262-271)."""
        import torch
        L = _bind()
        _native.require_device()
        self.close()
        n, F = bins.shape
        cp = _CParams(self.p.max_depth, self.p.learning_rate, self.p.subsample, self.p.min_child_weight, self.p.reg_alpha,
                      self.p.reg_lambda, self.p.scale_pos_weight, self.p.seed & MASK64)
        self._keep = (bins, y, weight)
        _native.check(L.mlrn_gbdt_create(bins.device.index, bins.data_ptr(), n, F, y.data_ptr(), weight.data_ptr() if weight is not None else None,
                                         C.byref(cp), rounds, stream, C.byref(self._h)), "mlrn_gbdt_create")
        self.F = F
        val_margin = torch.zeros(eval_bins.shape[0], dtype=torch.float64, device=bins.device) if eval_bins is not None else None
        best_loss, best_margin, since = float("inf"), None, 0
        self.best_round = 0
        for r in range(rounds):
            mask = column_mask(self.p.seed, r, F, self.p.colsample_bytree)
            _native.check(L.mlrn_gbdt_round(self._h, mask.ctypes.data, stream), "mlrn_gbdt_round")
            self.best_round = r + 1
            if eval_bins is None:
                continue
            _native.check(L.mlrn_gbdt_predict(self._h, eval_bins.data_ptr(), eval_bins.shape[0], r, r + 1, val_margin.data_ptr(), stream),
                          "mlrn_gbdt_predict")
            loss = float(torch.nn.functional.binary_cross_entropy_with_logits(val_margin, eval_y))       # metric only: one scalar per round
            if loss < best_loss:
                best_loss, best_margin, since, best_r = loss, val_margin.clone(), 0, r + 1
            else:
                since += 1
                if since >= early_stopping_rounds:
                    break
        if eval_bins is not None:
            self.best_round = best_r
            self.eval_margin = best_margin
        return self

    def predict_margin(self, bins, stream: int = 0):
        import torch
        L = _bind()
        m = torch.zeros(bins.shape[0], dtype=torch.float64, device=bins.device)
        _native.check(L.mlrn_gbdt_predict(self._h, bins.data_ptr(), bins.shape[0], 0, self.best_round, m.data_ptr(), stream), "mlrn_gbdt_predict")
        return m

    def predict_proba(self, bins, stream: int = 0):
        import torch
        return torch.sigmoid(self.predict_margin(bins, stream))

    def trees(self):
        """Host copy of the trees: dict of arrays ``feature, bin, default_left`` [rounds, 63] int32 and ``value`` float64."""
        L = _bind()
        k = L.mlrn_gbdt_rounds(self._h)
        arr = (_CTree * max(k, 1))()
        _native.check(L.mlrn_gbdt_export(self._h, C.addressof(arr), None), "mlrn_gbdt_export")
        return {"feature": np.array([list(arr[i].feature) for i in range(k)], np.int32).reshape(k, 63),
                "bin": np.array([list(arr[i].bin) for i in range(k)], np.int32).reshape(k, 63),
                "default_left": np.array([list(arr[i].default_left) for i in range(k)], np.int32).reshape(k, 63),
                "value": np.array([list(arr[i].value) for i in range(k)], np.float64).reshape(k, 63)}

    def train_margin(self) -> np.ndarray:
        L = _bind()
        out = np.empty(self._keep[0].shape[0])
        _native.check(L.mlrn_gbdt_export(self._h, None, out.ctypes.data), "mlrn_gbdt_export")
        return out

    def close(self):
        if self._h:
            _native.lib().mlrn_gbdt_destroy(self._h)
            self._h = C.c_void_p()
        self._keep = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def stratified_folds(y: np.ndarray, k: int = 5, seed: int = 42):
    """``StratifiedKFold(n_splits=k, shuffle=True, random_state=seed)`` (train_v92_focal_adversarial.py:214)."""
    from sklearn.model_selection import StratifiedKFold
    return list(StratifiedKFold(n_splits=k, shuffle=True, random_state=seed).split(np.zeros(len(y)), y))


def best_f1(y: np.ndarray, p: np.ndarray, grid=None):
    """Best-threshold F1 of train_v92_focal_adversarial.py:283-289 (``linspace(0.05, 0.5, 200)``)."""
    best, best_t = 0.0, 0.3
    for t in (np.linspace(0.05, 0.5, 200) if grid is None else grid):
        pred = p > t
        tp = np.sum(pred & (y == 1)); fp = np.sum(pred & (y == 0)); fn = np.sum(~pred & (y == 1))
        f1 = 2 * tp / max(2 * tp + fp + fn, 1)
        if f1 > best:
            best, best_t = f1, t
    return float(best), float(best_t)


def cross_validate(table, y: np.ndarray, weight: Optional[np.ndarray] = None, params: GBDTParams = GBDTParams(), n_folds: int = 5,
                   rounds: int = 500, early_stopping_rounds: int = 50, use_scale_pos_weight: bool = True, fold_seed: int = 42):
    """The v92d loop (train_v92_focal_adversarial.py:214-281) on a device-resident table: out-of-fold probabilities, the
    best-threshold OOF F1 and the rounds each fold kept.  ``table`` [n, F] float64 CUDA; ``y`` / ``weight`` host arrays."""
    import torch
    from dataclasses import replace
    binner = Binner(table)
    bins = binner.bins(table)
    dev = table.device
    yd = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float64)).to(dev)
    wd = torch.from_numpy(np.ascontiguousarray(weight, dtype=np.float64)).to(dev) if weight is not None else None
    oof = np.zeros(len(y))
    kept = []
    for tr, va in stratified_folds(y, n_folds, fold_seed):
        tr_d, va_d = torch.from_numpy(tr).to(dev), torch.from_numpy(va).to(dev)
        p = params
        if use_scale_pos_weight:
            p = replace(params, scale_pos_weight=float((y[tr] == 0).sum()) / float(max((y[tr] == 1).sum(), 1)))
        model = HistGBDT(p).fit(bins[tr_d].contiguous(), yd[tr_d].contiguous(), wd[tr_d].contiguous() if wd is not None else None,
                                bins[va_d].contiguous(), yd[va_d].contiguous(), rounds, early_stopping_rounds)
        oof[va] = torch.sigmoid(model.eval_margin).cpu().numpy()
        kept.append(model.best_round)
        model.close()
    f1, thr = best_f1(y, oof)
    return {"oof": oof, "f1": f1, "threshold": thr, "rounds": kept}


def adversarial_weights(train_table, test_table, seed: int = 42):
    """adversarial_validation.py:120-160,236-246: a train-vs-test classifier (depth 3, learning rate 0.05, 100 rounds, early
    stopping 20) scored out of fold; train rows that look like test rows get up to 4x the weight of those that do not:
    ``0.5 + 1.5 p`` when the AUC signals a shift (>= 0.55), otherwise all ones.  Returns ``(weights, auc)``."""
    import torch
    from sklearn.metrics import roc_auc_score
    X = torch.cat([train_table, test_table], dim=0)
    y = np.concatenate([np.zeros(train_table.shape[0]), np.ones(test_table.shape[0])])
    p = GBDTParams(max_depth=3, learning_rate=0.05, subsample=0.8, colsample_bytree=0.8, min_child_weight=1.0, reg_alpha=0.0,
                   reg_lambda=1.0, seed=seed)
    res = cross_validate(X, y, None, p, 5, 100, 20, use_scale_pos_weight=False, fold_seed=seed)
    auc = float(roc_auc_score(y, res["oof"]))
    p_train = res["oof"][: train_table.shape[0]]
    w = 0.5 + 1.5 * p_train if auc >= 0.55 else np.ones(train_table.shape[0])
    return w, auc
