"""Datamodel (RidgeCV) oracle against vectors produced by the reference's call (lds.py:411-421 -> sklearn)."""
import os

import numpy as np
import pytest

from oracle import aggregation as oagg

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ridge_golden.npz")
ALPHAS = np.linspace(0.01, 10, 100)


@pytest.mark.parametrize("tag", ["long", "wide", "square"])
def test_oracle_matches_golden(tag):
    g = np.load(GOLDEN)
    X, Y = g[f"{tag}_X"].astype(np.float64), g[f"{tag}_Y"]
    coef, alpha, ic = oagg.datamodel_ridge(X, Y)
    np.testing.assert_array_equal(alpha, g[f"{tag}_alpha"])
    np.testing.assert_allclose(coef, g[f"{tag}_coef"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(ic, g[f"{tag}_intercept"], rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("tag", ["long", "wide", "square"])
def test_single_decomposition_form_matches_sklearn(tag):
    """The algebra the CUDA kernels implement (one eigendecomposition for all alphas and behaviours) reproduces
    sklearn's per-behaviour fits: same alpha on the grid, coefficients to 1e-9."""
    g = np.load(GOLDEN)
    X, Y = g[f"{tag}_X"].astype(np.float64), g[f"{tag}_Y"]
    coef, alpha, ic, _ = oagg.ridge_gcv_closed_form(X, Y, ALPHAS)
    np.testing.assert_array_equal(alpha, g[f"{tag}_alpha"])
    np.testing.assert_allclose(coef, g[f"{tag}_coef"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(ic, g[f"{tag}_intercept"], rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("tag", ["wide", "tall", "odd"])
def test_bootstrapped_datamodel_oracle_matches_reference_function(tag):
    """oracle.datamodel == the reference's datamodel() (datamodel.py:8-37) run on the same global numpy seed."""
    g = np.load(os.path.join(os.path.dirname(GOLDEN), "datamodel_golden.npz"))
    X, Y = g[f"{tag}_X"].astype(np.float64), g[f"{tag}_Y"]
    np.random.seed(int(g[f"{tag}_seed"]))
    coeff = oagg.datamodel(X, Y, g[f"{tag}_coeff"].shape[0])
    np.testing.assert_allclose(coeff, g[f"{tag}_coeff"], rtol=1e-9, atol=1e-12)
