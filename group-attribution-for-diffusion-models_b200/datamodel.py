"""Datamodel (ridge) attribution estimator of the LDS fit sweep, on sm_100a kernels.

The reference fits, for every model behaviour ``i`` (lds.py:411-421)::

    datamodel = RidgeCV(alphas=np.linspace(0.01, 10, 100)).fit(train_masks_fold, train_targets_fold[:, i])
    coeff = datamodel.coef_          # datamodel.alpha_ is printed

i.e. sklearn's efficient leave-one-out (GCV) ridge with an intercept, one fit -- and one decomposition of the same
mask matrix -- per behaviour.  ``ridge_cv_batched`` does the decomposition once and evaluates every (alpha,
behaviour) pair in one kernel (csrc/ridge.cuh); ``RidgeCV`` keeps sklearn's constructor / ``fit`` / attribute names
for the subset the reference uses so ``lds.py`` can swap the import.  fp64 throughout; no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .aggregation import _dev_f64, _device

_f64 = torch.float64
DEFAULT_ALPHAS = np.linspace(0.01, 10, 100)  # lds.py:413


def _gemm(h, dev, trans_a: bool, a: torch.Tensor, b: torch.Tensor, m: int, j: int, n: int) -> torch.Tensor:
    c = torch.empty(m, n, dtype=_f64, device=dev)
    _lib.check(h.lib.gadm_dgemm(h.ptr, 1 if trans_a else 0, a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), m, j, n,
                                c.data_ptr(), c.stride(0), _lib.stream_ptr(dev)))
    return c


def ridge_cv_batched(x_train, y_train, alphas=DEFAULT_ALPHAS, alpha_per_target: bool = True, device=None,
                     as_numpy: bool = True):
    """Leave-one-out ridge with intercept for all K targets at once.

    x_train [n, d], y_train [n] or [n, K], alphas [A]  ->  dict(coef [d, K], intercept [K], alpha [K],
    alpha_index [K], best_score [K], cv_scores [A, K]).  ``alpha_per_target=True`` reproduces the reference's
    per-behaviour fits; ``False`` is sklearn's multi-target default (one alpha for all targets).
    """
    dev = _device(device)
    h = _lib.get_handle(dev)
    x = _dev_f64(x_train, dev)
    y = _dev_f64(y_train, dev)
    if x.dim() != 2:
        raise ValueError(f"x_train must be [n, d], got {tuple(x.shape)}")
    if y.dim() == 1:
        y = y[:, None].contiguous()
    n, d = int(x.shape[0]), int(x.shape[1])
    if y.shape[0] != n:
        raise ValueError(f"Found input variables with inconsistent numbers of samples: [{n}, {y.shape[0]}]")
    K = int(y.shape[1])
    al = np.asarray(alphas, dtype=np.float64).reshape(-1)
    if al.size == 0 or np.any(al <= 0):
        raise ValueError("alphas must be a non-empty array of positive floats")  # sklearn: Interval(0, None, 'neither')
    al_t = _dev_f64(al, dev)
    A = int(al.size)
    st = _lib.stream_ptr(dev)
    with torch.cuda.device(dev):
        xc = torch.empty_like(x); xmean = torch.empty(d, dtype=_f64, device=dev)
        yc = torch.empty_like(y); ymean = torch.empty(K, dtype=_f64, device=dev)
        _lib.check(h.lib.gadm_center_columns(h.ptr, x.data_ptr(), n, d, xc.data_ptr(), xmean.data_ptr(), st))
        _lib.check(h.lib.gadm_center_columns(h.ptr, y.data_ptr(), n, K, yc.data_ptr(), ymean.data_ptr(), st))
        cov = _gemm(h, dev, True, xc, xc, d, n, d)                       # Xc^T Xc
        evals = torch.empty(d, dtype=_f64, device=dev)
        v = torch.empty(d, d, dtype=_f64, device=dev)
        ws = torch.empty(int(h.lib.gadm_sym_eig_workspace_bytes(d)), dtype=torch.uint8, device=dev)
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.check(h.lib.gadm_sym_eig(h.ptr, cov.data_ptr(), d, evals.data_ptr(), v.data_ptr(), ws.data_ptr(), ws.numel(),
                                      C.cast(info.data_ptr(), C.POINTER(C.c_int)), st))
        z = _gemm(h, dev, False, xc, v, n, d, d)                         # Z = Xc V
        t = _gemm(h, dev, True, z, yc, d, n, K)                          # T = Z^T Yc
        q = torch.empty(d, dtype=_f64, device=dev)
        den = torch.empty(A, n, dtype=_f64, device=dev)
        score = torch.empty(A, K, dtype=_f64, device=dev)
        _lib.check(h.lib.gadm_ridge_gcv(h.ptr, z.data_ptr(), t.data_ptr(), yc.data_ptr(), evals.data_ptr(), al_t.data_ptr(),
                                        n, d, K, A, q.data_ptr(), den.data_ptr(), score.data_ptr(), st))
        best = torch.empty(K, dtype=torch.int32, device=dev)
        best_score = torch.empty(K, dtype=_f64, device=dev)
        ts = torch.empty_like(t)
        _lib.check(h.lib.gadm_ridge_select(h.ptr, score.data_ptr(), A, K, 1 if alpha_per_target else 0, al_t.data_ptr(),
                                           evals.data_ptr(), t.data_ptr(), d, best.data_ptr(), best_score.data_ptr(),
                                           ts.data_ptr(), st))
        coef = _gemm(h, dev, False, v, ts, d, d, K)                      # coef = V diag(1 / (L + alpha)) T
        intercept = torch.empty(K, dtype=_f64, device=dev)
        _lib.check(h.lib.gadm_ridge_intercept(h.ptr, coef.data_ptr(), xmean.data_ptr(), ymean.data_ptr(), d, K,
                                              intercept.data_ptr(), st))
    out = {"coef": coef, "intercept": intercept, "alpha_index": best, "best_score": best_score, "cv_scores": score}
    if as_numpy:
        out = {k_: v_.cpu().numpy() for k_, v_ in out.items()}
        out["alpha"] = al[out["alpha_index"]]
    else:
        out["alpha"] = al_t[best.long()]
    return out


def datamodel_ridge_batched(train_masks, train_targets, alphas=DEFAULT_ALPHAS, device=None, as_numpy: bool = True):
    """coeff [d, K] of lds.py:411-421 for all behaviours (column i == ``RidgeCV(alphas).fit(masks, targets[:, i]).coef_``)."""
    return ridge_cv_batched(train_masks, train_targets, alphas, True, device, as_numpy)["coef"]


class RidgeCV:
    """``sklearn.linear_model.RidgeCV`` for the configuration the reference uses (lds.py:413): default leave-one-out
    CV, ``fit_intercept=True``, no sample weights, default scoring.  Attributes after ``fit``: ``coef_``
    ([d] for 1-D y, else [K, d]), ``intercept_``, ``alpha_``, ``best_score_``."""

    def __init__(self, alphas=(0.1, 1.0, 10.0), *, fit_intercept=True, scoring=None, cv=None, gcv_mode=None,
                 store_cv_results=False, alpha_per_target=False, device=None):
        if not fit_intercept or scoring is not None or cv is not None:
            raise NotImplementedError("only the reference's configuration is built: fit_intercept=True, scoring=None, "
                                      "cv=None (efficient leave-one-out)")
        self.alphas = alphas
        self.fit_intercept = fit_intercept
        self.gcv_mode = gcv_mode
        self.store_cv_results = store_cv_results
        self.alpha_per_target = alpha_per_target
        self.device = device

    def fit(self, X, y, sample_weight=None):
        if sample_weight is not None:
            raise NotImplementedError("sample_weight is not used by the reference (lds.py:413)")
        y = np.asarray(y, dtype=np.float64)
        one_d = y.ndim == 1
        res = ridge_cv_batched(X, y, self.alphas, alpha_per_target=bool(self.alpha_per_target) or one_d, device=self.device)
        if one_d:
            self.coef_ = res["coef"][:, 0]
            self.intercept_ = float(res["intercept"][0])
            self.alpha_ = float(res["alpha"][0])
            self.best_score_ = float(res["best_score"][0])
        else:
            self.coef_ = res["coef"].T.copy()
            self.intercept_ = res["intercept"]
            if self.alpha_per_target:
                self.alpha_, self.best_score_ = res["alpha"], res["best_score"]
            else:
                self.alpha_, self.best_score_ = float(res["alpha"][0]), float(res["best_score"][0])
        if self.store_cv_results:
            self.cv_scores_ = res["cv_scores"]
        return self

    def predict(self, X):
        X = np.asarray(X, dtype=np.float64)
        return X @ (self.coef_ if self.coef_.ndim == 1 else self.coef_.T) + self.intercept_
