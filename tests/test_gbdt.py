"""Histogram GBDT: the numpy oracle behaves like a boosted classifier and -- on the GPU -- csrc/gbdt.cu
grows the same trees bit for bit."""
import numpy as np
import pytest


def _problem(n=1200, F=24, seed=3):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, F))
    X[:, 3] = np.round(X[:, 3], 1)                                  # few distinct values
    X[:, 5] = (X[:, 5] > 0.3).astype(float)                         # binary column
    logit = 1.5 * X[:, 0] - 1.0 * X[:, 1] * X[:, 2] + 0.8 * (X[:, 3] > 0.5) + 0.5 * X[:, 5] - 2.2
    y = (rng.random(n) < 1 / (1 + np.exp(-logit))).astype(np.float64)
    X[rng.random(X.shape) < 0.08] = np.nan                          # missing values
    X[rng.random(X.shape) < 0.002] = np.inf
    w = rng.uniform(0.5, 2.0, n)
    return X, y, w


def test_fixed_polynomial_exp_is_accurate():
    from oracle.gbdt import gbdt_exp
    x = np.concatenate([np.linspace(-40, 40, 20001), [-700, 700, -745, 800, 0.0, 1e-300]])
    rel = np.abs(gbdt_exp(x) - np.exp(np.clip(x, -700, 700))) / np.exp(np.clip(x, -700, 700))
    assert rel.max() < 4e-16


def test_oracle_boosts_like_a_histogram_gbdt():
    """Against scikit-learn's histogram GBDT on the same data: held-out AUC within 0.03, and the training loss falls."""
    from sklearn.ensemble import HistGradientBoostingClassifier
    from sklearn.metrics import roc_auc_score
    from oracle import gbdt
    X, y, _ = _problem(1600, 12, seed=11)
    tr, te = np.arange(1200), np.arange(1200, 1600)
    Xc = np.where(np.isinf(X), np.sign(X) * 1e10, X)
    cuts, ncuts = gbdt.quantile_cuts(Xc[tr])
    b_tr, b_te = gbdt.apply_bins(Xc[tr], cuts, ncuts), gbdt.apply_bins(Xc[te], cuts, ncuts)
    bo = gbdt.Booster(b_tr, y[tr], None, gbdt.Params(max_depth=3, learning_rate=0.2, subsample=0.8, colsample_bytree=0.8, min_child_weight=1.0,
                                                       reg_alpha=0.0, reg_lambda=1.0, seed=1))
    loss = []
    for _ in range(25):
        bo.round()
        p = 1 / (1 + np.exp(-bo.margin))
        loss.append(-np.mean(y[tr] * np.log(p) + (1 - y[tr]) * np.log(1 - p)))
    assert loss[-1] < 0.8 * loss[0] and all(b <= a + 1e-3 for a, b in zip(loss, loss[1:]))
    auc = roc_auc_score(y[te], bo.predict_margin(b_te))
    ref = HistGradientBoostingClassifier(max_depth=3, learning_rate=0.2, max_iter=25, random_state=0).fit(Xc[tr], y[tr])
    auc_ref = roc_auc_score(y[te], ref.decision_function(Xc[te]))
    assert auc > 0.75 and abs(auc - auc_ref) < 0.03, (auc, auc_ref)
    assert np.allclose(bo.predict_margin(b_tr), bo.margin, rtol=0, atol=1e-12)       # traversal == training assignment


def test_sub_sampling_is_a_function_of_seed_round_and_index():
    from oracle import gbdt
    from mallorn_astrophysics_b200 import gbdt as prod
    u = gbdt.uniform(42, 7, np.arange(100000))
    assert 0.0 <= u.min() and u.max() < 1.0 and abs((u < 0.8).mean() - 0.8) < 0.01
    assert np.array_equal(u[:10], gbdt.uniform(42, 7, np.arange(10)))
    for F, frac in ((222, 0.8), (5, 0.8), (3, 0.1)):
        m = gbdt.column_mask(42, 3, F, frac)
        assert m.sum() == max(1, int(np.floor(frac * F))) and np.array_equal(m, prod.column_mask(42, 3, F, frac))


@pytest.mark.gpu
def test_gpu_trees_are_bit_identical_to_the_oracle():
    import torch
    from oracle import gbdt as og
    from mallorn_astrophysics_b200 import gbdt as pg
    X, y, w = _problem()
    table = torch.from_numpy(X).cuda()
    binner = pg.Binner(table)
    bins = binner.bins(table)
    Xc = np.where(np.isinf(X), np.sign(X) * 1e10, X)
    cuts, ncuts = og.quantile_cuts(Xc)
    assert np.array_equal(cuts, binner.cuts) and np.array_equal(ncuts, binner.ncuts)
    assert np.array_equal(bins.cpu().numpy(), og.apply_bins(Xc, cuts, ncuts))
    params = dict(max_depth=5, learning_rate=0.1, subsample=0.8, colsample_bytree=0.8, min_child_weight=3.0, reg_alpha=0.2, reg_lambda=1.5,
                  scale_pos_weight=float((y == 0).sum() / (y == 1).sum()), seed=42)
    rounds = 12
    model = pg.HistGBDT(pg.GBDTParams(**params)).fit(bins, torch.from_numpy(y).cuda(), torch.from_numpy(w).cuda(), rounds=rounds)
    got = model.trees()
    bo = og.Booster(bins.cpu().numpy(), y, w, og.Params(**params))
    for r in range(rounds):
        t = bo.round()
        live = np.zeros(63, bool)                                    # nodes reachable from the root
        stack = [0]
        while stack:
            k = stack.pop(); live[k] = True
            if t.feature[k] >= 0:
                stack += [2 * k + 1, 2 * k + 2]
        assert np.array_equal(got["feature"][r][live], t.feature[live]), r
        inner = live & (t.feature >= 0)
        assert np.array_equal(got["bin"][r][inner], t.bin[inner]) and np.array_equal(got["default_left"][r][inner], t.default_left[inner]), r
        assert np.array_equal(got["value"][r][live & ~inner], t.value[live & ~inner]), r          # bit for bit
    assert np.array_equal(model.train_margin(), bo.margin)
    assert np.array_equal(model.predict_margin(bins).cpu().numpy(), bo.predict_margin(bins.cpu().numpy()))
    model.close()


@pytest.mark.gpu
def test_cross_validation_and_adversarial_weights_on_a_device_table():
    import torch
    from mallorn_astrophysics_b200 import gbdt as pg
    X, y, w = _problem(1500, 16, seed=8)
    table = torch.from_numpy(X).cuda()
    res = pg.cross_validate(table, y, w, pg.GBDTParams(learning_rate=0.1), rounds=80, early_stopping_rounds=15)
    assert res["oof"].shape == (1500,) and 0.3 < res["f1"] <= 1.0 and all(1 <= r <= 80 for r in res["rounds"])
    shifted = table[:400].clone(); shifted[:, 0] += 1.5                     # a "test set" drawn from a shifted distribution
    wts, auc = pg.adversarial_weights(table[400:], shifted)
    assert auc > 0.6 and wts.shape == (1100,) and 0.5 <= wts.min() and wts.max() <= 2.0
    same, auc0 = pg.adversarial_weights(table[400:], table[:400])
    assert auc0 < 0.6
