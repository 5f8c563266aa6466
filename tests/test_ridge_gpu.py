"""GPU parity of the batched RidgeCV datamodel estimator (csrc/ridge.cuh) with sklearn's per-behaviour fits.

Tolerance: the selected alpha (grid index) must be identical; coefficients and intercepts agree to 1e-8 relative
(fp64 everywhere; the only differences are summation order and Jacobi vs LAPACK eigenvectors).
"""
import os

import numpy as np
import pytest

from oracle import aggregation as oagg

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ridge_golden.npz")
ALPHAS = np.linspace(0.01, 10, 100)


@pytest.mark.parametrize("tag", ["long", "wide", "square"])
def test_golden(tag):
    import gadm_b200 as G

    g = np.load(GOLDEN)
    X, Y = g[f"{tag}_X"].astype(np.float64), g[f"{tag}_Y"]
    res = G.ridge_cv_batched(X, Y, ALPHAS)
    np.testing.assert_array_equal(res["alpha"], g[f"{tag}_alpha"])
    np.testing.assert_allclose(res["coef"], g[f"{tag}_coef"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(res["intercept"], g[f"{tag}_intercept"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(G.datamodel_ridge_batched(X, Y), g[f"{tag}_coef"], rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("n,d,K", [(500, 100, 33), (257, 41, 5), (30, 64, 3), (1000, 100, 64)])
def test_seeded_vs_oracle(n, d, K):
    import gadm_b200 as G

    rng = np.random.RandomState(n + d)
    X = oagg.datamodel_masks(d, list(range(n)), alpha=0.5)
    w = rng.normal(size=(d, K))
    Y = X @ w + rng.uniform(0.05, 2.0, size=K)[None, :] * rng.normal(size=(n, K)) - 1.5
    coef, alpha, ic = oagg.datamodel_ridge(X, Y)
    res = G.ridge_cv_batched(X, Y)
    np.testing.assert_array_equal(res["alpha"], alpha)
    np.testing.assert_allclose(res["coef"], coef, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(res["intercept"], ic, rtol=1e-8, atol=1e-10)
    # the whole LOO score surface, not only its argmax
    _, _, _, scores = oagg.ridge_gcv_closed_form(X, Y, ALPHAS)
    np.testing.assert_allclose(res["cv_scores"], scores, rtol=1e-9)


def test_sklearn_class_surface():
    """lds.py:413-421 reads .alpha_ and .coef_ after fit(X, y_1d); also the multi-target default (one alpha)."""
    from sklearn.linear_model import RidgeCV as SkRidgeCV

    import gadm_b200 as G

    rng = np.random.RandomState(0)
    X = oagg.datamodel_masks(30, list(range(200)), alpha=0.5)
    Y = X @ rng.normal(size=(30, 4)) + 0.5 * rng.normal(size=(200, 4))
    ours = G.RidgeCV(alphas=ALPHAS).fit(X, Y[:, 1])
    ref = SkRidgeCV(alphas=ALPHAS).fit(X, Y[:, 1])
    assert ours.alpha_ == ref.alpha_ and ours.coef_.shape == ref.coef_.shape == (30,)
    np.testing.assert_allclose(ours.coef_, ref.coef_, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(ours.intercept_, ref.intercept_, rtol=1e-8)
    np.testing.assert_allclose(ours.best_score_, ref.best_score_, rtol=1e-9)
    np.testing.assert_allclose(ours.predict(X[:5]), ref.predict(X[:5]), rtol=1e-8)
    ours2 = G.RidgeCV(alphas=ALPHAS).fit(X, Y)
    ref2 = SkRidgeCV(alphas=ALPHAS).fit(X, Y)
    assert ours2.alpha_ == ref2.alpha_ and ours2.coef_.shape == ref2.coef_.shape == (4, 30)
    np.testing.assert_allclose(ours2.coef_, ref2.coef_, rtol=1e-8, atol=1e-10)
    ours3 = G.RidgeCV(alphas=ALPHAS, alpha_per_target=True).fit(X, Y)
    ref3 = SkRidgeCV(alphas=ALPHAS, alpha_per_target=True).fit(X, Y)
    np.testing.assert_array_equal(ours3.alpha_, ref3.alpha_)
    np.testing.assert_allclose(ours3.coef_, ref3.coef_, rtol=1e-8, atol=1e-10)


def test_general_real_features_and_errors():
    import gadm_b200 as G

    rng = np.random.RandomState(3)
    X = rng.normal(size=(120, 17))
    y = X @ rng.normal(size=17) + 0.1 * rng.normal(size=120)
    coef, alpha, ic = oagg.datamodel_ridge(X, y[:, None], alphas=[0.1, 1.0, 10.0])
    res = G.ridge_cv_batched(X, y, [0.1, 1.0, 10.0])
    np.testing.assert_array_equal(res["alpha"], alpha)
    np.testing.assert_allclose(res["coef"], coef, rtol=1e-8, atol=1e-10)
    with pytest.raises(ValueError):
        G.ridge_cv_batched(X, y[:-1])
    with pytest.raises(ValueError):
        G.ridge_cv_batched(X, y, [0.0, 1.0])
    with pytest.raises(ValueError):
        G.ridge_cv_batched(X, y, device="cpu")
    with pytest.raises(NotImplementedError):
        G.RidgeCV(cv=5)


@pytest.mark.parametrize("tag", ["wide", "tall", "odd"])
def test_bootstrapped_datamodel_golden(tag):
    """datamodel.py:8-37 through csrc/datamodel.cuh against the coefficients the reference's function produced
    (same global numpy seed => same resamples; same alpha per resample; coefficients to 1e-7)."""
    import gadm_b200 as G

    g = np.load(os.path.join(os.path.dirname(GOLDEN), "datamodel_golden.npz"))
    X, Y = g[f"{tag}_X"], g[f"{tag}_Y"]
    want = g[f"{tag}_coeff"]
    np.random.seed(int(g[f"{tag}_seed"]))
    got, det = G.datamodel(X, Y, want.shape[0], return_details=True)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=1e-7, atol=1e-9)
    # the grid-search scores themselves, against sklearn on the same resamples
    from sklearn.linear_model import Ridge
    from sklearn.model_selection import GridSearchCV

    for b in range(want.shape[0]):
        idx = det["bootstrap_indices"][b]
        gs = GridSearchCV(Ridge(), {"alpha": [0.1, 1.0, 10.0]}, cv=5).fit(X[idx].astype(np.float64), Y[idx])
        np.testing.assert_allclose(det["mean_test_score"][b], gs.cv_results_["mean_test_score"], rtol=1e-8, atol=1e-10)
        assert det["alpha"][b] == gs.best_params_["alpha"]


def test_compute_datamodel_scores_shape_and_values():
    import argparse

    import gadm_b200 as G

    d, n = 50, 40
    rng = np.random.RandomState(2)
    X = oagg.datamodel_masks(d, list(range(n)), alpha=0.5)
    w = rng.normal(size=d)
    recs = [{"remaining_idx": np.where(X[i] == 1)[0].tolist(), "removed_idx": np.where(X[i] == 0)[0].tolist(),
             "fid": float(X[i] @ w + 0.1 * rng.normal())} for i in range(n)]
    args = argparse.Namespace(dataset="cifar", model_behavior="fid", num_runs=3)
    train_idx, val_idx = list(range(30)), list(range(30, 40))
    np.random.seed(5)
    got = G.compute_datamodel_scores(args, recs, train_idx, val_idx, total_data_num=d)
    np.random.seed(5)
    Y = np.array([r["fid"] for r in recs])
    want = X[val_idx] @ oagg.datamodel(X[train_idx], Y[train_idx], 3).T
    assert got.shape == (10, 3)
    np.testing.assert_allclose(got, want, rtol=1e-7, atol=1e-9)
