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
        gws = torch.empty(int(h.lib.gadm_ridge_gcv_workspace_bytes(n, d, K, A)), dtype=torch.uint8, device=dev)
        score = torch.empty(A, K, dtype=_f64, device=dev)
        _lib.check(h.lib.gadm_ridge_gcv(h.ptr, z.data_ptr(), t.data_ptr(), yc.data_ptr(), evals.data_ptr(), al_t.data_ptr(),
                                        n, d, K, A, gws.data_ptr(), gws.numel(), score.data_ptr(), st))
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


# ---------------------------------------------------------------- bootstrapped datamodel (datamodel.py:8-37)

_SYSTEM_DTYPE = np.dtype([("run", np.int32), ("f0", np.int32), ("f1", np.int32), ("out", np.int32), ("alpha", np.float64)])


def _kfold_bounds(n: int, n_splits: int = 5):
    """sklearn KFold(n_splits, shuffle=False): the first n % n_splits folds have one extra sample."""
    sizes = np.full(n_splits, n // n_splits, dtype=np.int64)
    sizes[: n % n_splits] += 1
    ends = np.cumsum(sizes)
    return list(zip((ends - sizes).tolist(), ends.tolist()))


def datamodel(x_train, y_train, num_runs: int, alphas=(0.1, 1.0, 1e1), cv: int = 5, bootstrap_indices=None, device=None,
              return_details: bool = False):
    """Reference signature (src/attributions/methods/datamodel.py:8-37): ``num_runs`` bootstrap resamples, each fitted
    with ``RidgeCV(cv=5, alphas=[0.1, 1.0, 10.0])``; returns the stacked coefficients ``[num_runs, d]``.

    The resamples are drawn exactly like the reference, ``np.random.choice(n, n, replace=True)`` from numpy's *global*
    state (seed it the same way to reproduce a run), unless ``bootstrap_indices`` [num_runs, n] is given.  x_train
    must be 0/1 masks (they are ``remaining_idx`` indicators, datamodel.py:71); all ridge systems of all resamples are
    solved in one kernel launch per stage (csrc/datamodel.cuh)."""
    from .aggregation import PackedMasks

    dev = _device(device)
    h = _lib.get_handle(dev)
    masks = x_train if isinstance(x_train, PackedMasks) else PackedMasks(x_train, dev)
    n, d = masks.n, masks.d
    y = _dev_f64(np.asarray(y_train, dtype=np.float64).reshape(-1), dev)
    if y.shape[0] != n:
        raise ValueError(f"Found input variables with inconsistent numbers of samples: [{n}, {y.shape[0]}]")
    if n < cv:
        raise ValueError(f"Cannot have number of splits n_splits={cv} greater than the number of samples: n_samples={n}.")
    if bootstrap_indices is None:
        idx = np.stack([np.random.choice(n, n, replace=True) for _ in range(num_runs)])  # datamodel.py:27
    else:
        idx = np.asarray(bootstrap_indices).reshape(num_runs, n)
    idx_t = torch.as_tensor(np.ascontiguousarray(idx, dtype=np.int32)).to(dev)
    al = [float(a) for a in alphas]
    folds = _kfold_bounds(n, cv)
    st = _lib.stream_ptr(dev)
    with torch.cuda.device(dev):
        g0 = torch.empty(n, n, dtype=_f64, device=dev)  # X X^T: the row bit planes play the role of column planes
        _lib.check(h.lib.gadm_mask_gram(h.ptr, masks.rowbits.data_ptr(), d, n, 2, g0.data_ptr(), st))
        slot = int(h.lib.gadm_datamodel_slot_bytes(n))
        free, _ = torch.cuda.mem_get_info(dev)
        n_fold_sys = num_runs * len(al) * cv
        slots = max(1, min(n_fold_sys, 2 * torch.cuda.get_device_properties(dev).multi_processor_count,
                           int(0.5 * free) // slot))
        ws = torch.empty(slots * slot, dtype=torch.uint8, device=dev)
        # stage 1: every (resample, alpha, fold) system -> held-out R^2
        sysv = np.zeros(n_fold_sys, dtype=_SYSTEM_DTYPE)
        e = 0
        for b in range(num_runs):
            for ai, a in enumerate(al):
                for fi, (f0, f1) in enumerate(folds):
                    sysv[e] = (b, f0, f1, (b * len(al) + ai) * cv + fi, a)
                    e += 1
        scores = torch.empty(n_fold_sys, dtype=_f64, device=dev)
        wdual = torch.zeros(num_runs, n, dtype=_f64, device=dev)
        sys_t = torch.as_tensor(sysv.view(np.uint8)).to(dev)
        _lib.check(h.lib.gadm_datamodel_ridge_systems(h.ptr, g0.data_ptr(), y.data_ptr(), idx_t.data_ptr(), n, sys_t.data_ptr(),
                                                      n_fold_sys, ws.data_ptr(), ws.numel(), scores.data_ptr(),
                                                      wdual.data_ptr(), st))
        # GridSearchCV: mean test score over the folds, first best alpha (rank 'min' + argmin)
        mean_scores = scores.cpu().numpy().reshape(num_runs, len(al), cv).mean(axis=2)
        best = np.argmax(mean_scores, axis=1)
        # stage 2: refit every resample on all of its rows with its alpha
        sys2 = np.zeros(num_runs, dtype=_SYSTEM_DTYPE)
        for b in range(num_runs):
            sys2[b] = (b, 0, 0, b, al[best[b]])
        sys2_t = torch.as_tensor(sys2.view(np.uint8)).to(dev)
        _lib.check(h.lib.gadm_datamodel_ridge_systems(h.ptr, g0.data_ptr(), y.data_ptr(), idx_t.data_ptr(), n, sys2_t.data_ptr(),
                                                      num_runs, ws.data_ptr(), ws.numel(), scores.data_ptr(),
                                                      wdual.data_ptr(), st))
        coef = masks.xty(wdual.T.contiguous(), None, 0.0, 1.0)  # X^T w  -> [d, num_runs]
    coeff = coef.T.contiguous().cpu().numpy()
    if return_details:
        return coeff, {"alpha": np.asarray(al)[best], "mean_test_score": mean_scores, "bootstrap_indices": idx}
    return coeff


def compute_datamodel_scores(args, model_behavior_all, train_idx, val_idx, total_data_num: int | None = None, device=None):
    """Reference signature (datamodel.py:40-80): masks / behaviours from the jsonl records, bootstrapped datamodel on
    ``train_idx``, predictions ``X[val_idx] @ coeff.T`` -> [len(val_idx), num_runs].  ``total_data_num`` replaces
    ``len(create_dataset(args.dataset, train=True))`` (dataset I/O is out of scope)."""
    from .aggregation import PackedMasks

    if total_data_num is None:
        from src.datasets import create_dataset  # the reference's module

        total_data_num = len(create_dataset(dataset_name=args.dataset, train=True))
    train_val_index = list(train_idx) + list(val_idx)
    X = np.zeros((len(train_val_index), total_data_num), dtype=np.uint8)
    Y = np.zeros(len(train_val_index))
    for i in train_val_index:
        remaining_idx = model_behavior_all[i].get("remaining_idx", [])
        removed_idx = model_behavior_all[i].get("removed_idx", [])
        if total_data_num != len(remaining_idx) + len(removed_idx):
            print(f"AssertionError for index {i}: Total data number mismatch.")  # datamodel.py:62-76 (row stays zero)
            continue
        X[i, remaining_idx] = 1
        Y[i] = model_behavior_all[i].get(args.model_behavior)
    coeff = datamodel(X[list(train_idx)], Y[list(train_idx)], args.num_runs, device=device)  # [runs, d]
    val = PackedMasks(X[list(val_idx)], device)
    return val.times(_dev_f64(coeff.T, val.device)).cpu().numpy()
