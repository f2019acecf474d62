"""Oracle for the histogram GBDT (test infrastructure; see ``oracle/__init__.py``): a numpy restatement of the algorithm
of ``csrc/gbdt.cu`` -- xgboost's published ``tree_method='hist'`` for binary:logistic as the reference's drivers configure
it (``non_successful_tests/scripts/train_v92_focal_adversarial.py:157-169``) -- operation for operation, so that the GPU
trees can be compared bit for bit: gradients quantised to 2^-30 and summed in int64, the fixed-polynomial ``gbdt_exp``,
the gain evaluated with individually rounded operations, counter-based sub-sampling.

xgboost itself is not installed (parity unpinned against it); what IS checked is that this restatement
behaves like a gradient-boosted classifier should (``tests/test_gbdt.py``: against scikit-learn's histogram GBDT on the
same data the out-of-fold AUC / F1 agree within the spread of seeds)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

Q = 1073741824.0
MISSING = 255
NODES = 63
MASK64 = (1 << 64) - 1


@dataclass
class Params:
    max_depth: int = 5
    learning_rate: float = 0.025
    subsample: float = 0.8
    colsample_bytree: float = 0.8
    min_child_weight: float = 3.0
    reg_alpha: float = 0.2
    reg_lambda: float = 1.5
    scale_pos_weight: float = 1.0
    seed: int = 42


def gb_hash(seed: int, a: int, b) -> np.ndarray:
    """splitmix64-style hash of (seed, a, b) in uint64 arithmetic (wraps mod 2^64); ``b``: array of non-negative ints."""
    with np.errstate(over="ignore"):
        b = np.asarray(b, dtype=np.uint64)
        z = np.uint64(seed & MASK64) + np.uint64(0x9E3779B97F4A7C15) * np.uint64(a + 1) + np.uint64(0xBF58476D1CE4E5B9) * (b + np.uint64(1))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform(seed: int, a: int, idx) -> np.ndarray:
    return (gb_hash(seed, a, idx) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def column_mask(seed: int, rnd: int, F: int, frac: float) -> np.ndarray:
    """colsample_bytree: the ``max(1, floor(frac * F))`` columns with the smallest hash of (seed ^ COL, round, column)."""
    k = max(1, int(np.floor(frac * F)))
    if k >= F:
        return np.ones(F, np.uint8)
    h = gb_hash(seed ^ 0xC01C01C01, rnd, np.arange(F))
    order = sorted(range(F), key=lambda j: (int(h[j]), j))
    m = np.zeros(F, np.uint8)
    m[order[:k]] = 1
    return m


_C = [1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0,
      1.0 / 6.0, 0.5, 1.0, 1.0]


def gbdt_exp(x: np.ndarray) -> np.ndarray:
    x = np.clip(np.asarray(x, dtype=np.float64), -700.0, 700.0)
    kf = np.rint(x * 1.4426950408889634)
    r = (x - kf * 0.693147180369123816490) - kf * 1.90821492927058770002e-10
    p = np.full_like(x, 1.0 / 479001600.0)
    for c in _C:
        p = p * r + c
    return np.ldexp(p, kf.astype(np.int64))


def quantile_cuts(X: np.ndarray, max_cuts: int = 254):
    """Per feature: cut points (at most 254) -- midpoints between distinct values when there are few, otherwise midpoints
    at evenly spaced quantiles of the distinct values.  Host-side in the product too (``gbdt.py`` calls this very
    function's twin on a sample)."""
    F = X.shape[1]
    cuts = np.full((F, 255), np.inf)
    ncuts = np.zeros(F, np.int32)
    for f in range(F):
        v = np.unique(X[~np.isnan(X[:, f]), f])
        if v.size <= 1:
            continue
        mids = 0.5 * (v[:-1] + v[1:])
        if mids.size > max_cuts:
            pick = np.unique(np.round(np.linspace(0, mids.size - 1, max_cuts)).astype(int))
            mids = mids[pick]
        cuts[f, :mids.size] = mids
        ncuts[f] = mids.size
    return cuts, ncuts


def apply_bins(X: np.ndarray, cuts: np.ndarray, ncuts: np.ndarray) -> np.ndarray:
    n, F = X.shape
    bins = np.full((n, F), MISSING, np.uint8)
    for f in range(F):
        ok = ~np.isnan(X[:, f])
        b = np.searchsorted(cuts[f, :ncuts[f]], X[ok, f], side="left")
        bins[ok, f] = np.minimum(b, MISSING - 1)
    return bins


def _score(G, H, alpha, lam):
    a = np.abs(G) - alpha
    T = np.where(a > 0.0, a, 0.0)
    return (T * T) / (H + lam)


@dataclass
class Tree:
    feature: np.ndarray = field(default_factory=lambda: np.full(NODES, -1, np.int32))
    bin: np.ndarray = field(default_factory=lambda: np.zeros(NODES, np.int32))
    default_left: np.ndarray = field(default_factory=lambda: np.zeros(NODES, np.int32))
    value: np.ndarray = field(default_factory=lambda: np.zeros(NODES))


class Booster:
    def __init__(self, bins: np.ndarray, y: np.ndarray, weight: Optional[np.ndarray], params: Params):
        self.bins, self.y, self.w, self.p = bins, y.astype(np.float64), weight, params
        self.n, self.F = bins.shape
        self.margin = np.zeros(self.n)
        self.trees: List[Tree] = []

    def round(self):
        p, rnd = self.p, len(self.trees)
        colmask = column_mask(p.seed, rnd, self.F, p.colsample_bytree)
        u = uniform(p.seed, rnd, np.arange(self.n))
        sampled = u < p.subsample
        prob = 1.0 / (1.0 + gbdt_exp(-self.margin))
        wt = np.ones(self.n) if self.w is None else self.w.astype(np.float64).copy()
        wt = np.where(self.y > 0.5, wt * p.scale_pos_weight, wt)
        g = (prob - self.y) * wt
        h = np.maximum((prob * (1.0 - prob)) * wt, 1e-16)
        gq = np.where(sampled, np.rint(g * Q), 0).astype(np.int64)
        hq = np.where(sampled, np.rint(h * Q), 0).astype(np.int64)
        node = np.zeros(self.n, np.int64)
        sums = np.zeros((NODES, 2), np.int64)
        sums[0] = gq.sum(), hq.sum()
        tree = Tree()
        cols = np.flatnonzero(colmask)
        for depth in range(p.max_depth + 1):
            first, n_nodes = (1 << depth) - 1, 1 << depth
            for nd in range(n_nodes):
                nid = first + nd
                Gt, Ht = int(sums[nid, 0]), int(sums[nid, 1])
                best, bf, bb, bl = 0.0, -1, -1, 0
                hist_f = None
                if depth < p.max_depth and Ht != 0:
                    rows = np.flatnonzero((node == nid) & ((gq != 0) | (hq != 0)))
                    parent = float(_score(np.float64(Gt) / Q, np.float64(Ht) / Q, p.reg_alpha, p.reg_lambda))
                    hists = {}
                    for f in cols:
                        hg = np.zeros(256, np.int64); hh = np.zeros(256, np.int64)
                        np.add.at(hg, self.bins[rows, f], gq[rows]); np.add.at(hh, self.bins[rows, f], hq[rows])
                        hists[f] = (hg, hh)
                        cg, ch = np.cumsum(hg[:MISSING - 1]), np.cumsum(hh[:MISSING - 1])        # left = value bins 0..b, b <= 253
                        fb, fg, fl = -1, 0.0, 0
                        for miss_left in (0, 1):
                            GL = cg + (hg[MISSING] if miss_left else 0)
                            HL = ch + (hh[MISSING] if miss_left else 0)
                            GR, HR = Gt - GL, Ht - HL
                            hl, hr = HL.astype(np.float64) / Q, HR.astype(np.float64) / Q
                            ok = (hl >= p.min_child_weight) & (hr >= p.min_child_weight)
                            gain = (_score(GL.astype(np.float64) / Q, hl, p.reg_alpha, p.reg_lambda)
                                    + _score(GR.astype(np.float64) / Q, hr, p.reg_alpha, p.reg_lambda)) - parent
                            gain = np.where(ok, gain, -np.inf)
                            for b in np.flatnonzero(gain > 0.0):
                                gb = float(gain[b])
                                # larger gain, then lower bin, then missing-right
                                if fb < 0 or gb > fg or (gb == fg and (b < fb or (b == fb and miss_left < fl))):
                                    fb, fg, fl = int(b), gb, miss_left
                        if fb >= 0 and (bf < 0 or fg > best):          # ascending f: the first of equal gains stays
                            best, bf, bb, bl = fg, int(f), fb, fl
                    hist_f = hists.get(bf)
                split = depth < p.max_depth and bf >= 0 and best > 1e-6 and Ht != 0
                if not split:
                    val = 0.0
                    if Ht != 0:
                        G, H = np.float64(Gt) / Q, np.float64(Ht) / Q
                        a = abs(G) - p.reg_alpha
                        T = (a if G > 0.0 else -a) if a > 0.0 else 0.0
                        val = float(-(np.float64(T) / (H + p.reg_lambda)) * p.learning_rate)
                    tree.feature[nid] = -1; tree.value[nid] = val
                else:
                    tree.feature[nid], tree.bin[nid], tree.default_left[nid] = bf, bb, bl
                    hg, hh = hist_f
                    GL = int(hg[:bb + 1].sum()) + (int(hg[MISSING]) if bl else 0)
                    HL = int(hh[:bb + 1].sum()) + (int(hh[MISSING]) if bl else 0)
                    sums[2 * nid + 1] = GL, HL
                    sums[2 * nid + 2] = Gt - GL, Ht - HL
            if depth < p.max_depth:
                for nd in range(n_nodes):
                    nid = first + nd
                    f = tree.feature[nid]
                    if f < 0:
                        continue
                    rows = np.flatnonzero(node == nid)
                    b = self.bins[rows, f]
                    left = np.where(b == MISSING, tree.default_left[nid] != 0, b <= tree.bin[nid])
                    node[rows] = 2 * nid + np.where(left, 1, 2)
        self.margin = self.margin + tree.value[node]
        self.trees.append(tree)
        return tree

    def predict_margin(self, bins: np.ndarray, tree_from: int = 0, tree_to: Optional[int] = None, margin=None) -> np.ndarray:
        m = np.zeros(bins.shape[0]) if margin is None else margin.copy()
        for tr in self.trees[tree_from:tree_to]:
            node = np.zeros(bins.shape[0], np.int64)
            for _ in range(6):
                f = tr.feature[node]
                live = f >= 0
                if not live.any():
                    break
                b = bins[np.arange(len(node)), np.maximum(f, 0)]
                left = np.where(b == MISSING, tr.default_left[node] != 0, b <= tr.bin[node])
                node = np.where(live, 2 * node + np.where(left, 1, 2), node)
            m = m + tr.value[node]
        return m
