"""Shapley / Banzhaf estimators and LDS evaluation with the reference's signatures, on sm_100a kernels.

Drop-in for
  ``data_shapley(dataset_size, x_train, y_train, v1, v0)``  src/attributions/methods/datashapley.py:8-48
  ``data_banzhaf(x_train, y_train)``                         src/attributions/methods/databanzhaf.py:5-26
  ``evaluate_lds(attrs_all, test_data_list, num_model_behaviors)``
        text_to_image/shapley_lds.py:138-150 (``attrs_all[:, k]``), lds.py:158-170 (``attrs_all[k]``)
as called from text_to_image/shapley_lds.py:246-283, text_to_image/banzhaf_lds.py:162-178 and
lds.py:403-456.  The reference calls the estimators once per model behaviour and recomputes X^T X and
an SVD-based pseudo-inverse every time; here the batched forms do that work once per mask matrix and
solve all K behaviours together -- the per-behaviour wrappers just call them with K = 1.

Masks must be 0/1 (they are ``remaining_idx`` indicator rows, shapley_lds.py:114-119); they are
bit-packed on upload.  All arithmetic is fp64.  No CPU path: inputs are copied to the CUDA device,
results come back as numpy arrays (the reference scripts are numpy programs).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

_f64 = torch.float64


def _device(device=None) -> torch.device:
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("gadm_b200.aggregation needs a CUDA device (sm_100a); there is no CPU fallback")
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise ValueError(f"gadm kernels run on CUDA devices only; got '{device}'")
    return device if device.index is not None else torch.device("cuda", torch.cuda.current_device())


class PackedMasks:
    """Bit-packed subset masks resident on the device (row-major and column-major bit planes)."""

    def __init__(self, x, device=None):
        self.device = _device(device)
        self.h = _lib.get_handle(self.device)
        x = torch.as_tensor(np.asarray(x) if not isinstance(x, torch.Tensor) else x)
        if x.dim() != 2:
            raise ValueError(f"masks must be [n, d], got {tuple(x.shape)}")
        if not bool(((x == 0) | (x == 1)).all()):
            raise ValueError("subset masks must contain only 0 and 1")
        self.n, self.d = int(x.shape[0]), int(x.shape[1])
        x8 = x.to(torch.uint8).contiguous().to(self.device, non_blocking=True)
        self.wd, self.wn = (self.d + 31) // 32, (self.n + 31) // 32
        self.rowbits = torch.empty(self.n, self.wd, dtype=torch.int32, device=self.device)
        self.colbits = torch.empty(self.d, self.wn, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.h.lib.gadm_pack_masks(self.h.ptr, x8.data_ptr(), self.n, self.d, self.rowbits.data_ptr(),
                                                 self.colbits.data_ptr(), _lib.stream_ptr(self.device)))

    # A = X^T X / n  or (X - 1/2)^T (X - 1/2)
    def gram(self, mode: int) -> torch.Tensor:
        a = torch.empty(self.d, self.d, dtype=_f64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.h.lib.gadm_mask_gram(self.h.ptr, self.colbits.data_ptr(), self.n, self.d, mode, a.data_ptr(),
                                                _lib.stream_ptr(self.device)))
        return a

    def xty(self, y: torch.Tensor, shift, half: float, scale: float) -> torch.Tensor:
        out = torch.empty(self.d, y.shape[1], dtype=_f64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.h.lib.gadm_mask_xty(self.h.ptr, self.rowbits.data_ptr(), y.data_ptr(), self.n, self.d,
                                               y.shape[1], shift.data_ptr() if shift is not None else None,
                                               float(half), float(scale), out.data_ptr(), _lib.stream_ptr(self.device)))
        return out

    def times(self, mat: torch.Tensor) -> torch.Tensor:
        out = torch.empty(self.n, mat.shape[1], dtype=_f64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.h.lib.gadm_mask_times_matrix(self.h.ptr, self.colbits.data_ptr(), mat.data_ptr(), self.n,
                                                        self.d, mat.shape[1], out.data_ptr(),
                                                        _lib.stream_ptr(self.device)))
        return out


def masks_from_remaining_idx(remaining_idx_list, num_groups: int, device=None) -> PackedMasks:
    """Bit-packed masks from the jsonl records' ``remaining_idx`` lists (shapley_lds.py:114-119, lds.py:232-233),
    built as uint8 on the host (the jsonl / pandas parsing stays host Python) and packed on the device."""
    x = np.zeros((len(remaining_idx_list), num_groups), dtype=np.uint8)
    for r, idx in enumerate(remaining_idx_list):
        x[r, np.asarray(idx, dtype=np.int64)] = 1
    return PackedMasks(x, device)


def _dev_f64(a, device) -> torch.Tensor:
    if not isinstance(a, torch.Tensor):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
        if not a.flags.writeable:
            a = a.copy()
    t = torch.as_tensor(a)
    return t.to(device=device, dtype=_f64, non_blocking=True).contiguous()


def sym_pinv(a: torch.Tensor, rcond: float, return_info: bool = False):
    """numpy.linalg.pinv semantics for a symmetric fp64 matrix on the device (one-sided Jacobi SVD)."""
    device = a.device
    h = _lib.get_handle(device)
    d = a.shape[0]
    out = torch.empty_like(a)
    ws = torch.empty(int(h.lib.gadm_sym_pinv_workspace_bytes(d)), dtype=torch.uint8, device=device)
    info = torch.zeros(2, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        _lib.check(h.lib.gadm_sym_pinv(h.ptr, a.data_ptr(), d, float(rcond), out.data_ptr(), ws.data_ptr(), ws.numel(),
                                      C.cast(info.data_ptr(), C.POINTER(C.c_int)), _lib.stream_ptr(device)))
    return (out, info) if return_info else out


def _dgemm(a: torch.Tensor, b: torch.Tensor, zero_below: float = 0.0) -> torch.Tensor:
    h = _lib.get_handle(a.device)
    c = torch.empty_like(b)
    with torch.cuda.device(a.device):
        _lib.check(h.lib.gadm_dgemm_dk(h.ptr, a.data_ptr(), b.data_ptr(), a.shape[0], b.shape[1], float(zero_below),
                                      c.data_ptr(), _lib.stream_ptr(a.device)))
    return c


def data_shapley_batched(x_train, y_train, v1, v0, device=None, as_numpy: bool = True):
    """Closed-form KernelSHAP (Covert & Lee eq. 7) for all K behaviours at once.

    x_train [n, d] 0/1 masks (or PackedMasks), y_train [n, K], v1 [K], v0 [K] -> phi [d, K]."""
    masks = x_train if isinstance(x_train, PackedMasks) else PackedMasks(x_train, device)
    dev = masks.device
    y = _dev_f64(y_train, dev)
    if y.dim() == 1:
        y = y[:, None]
    K = y.shape[1]
    v1 = _dev_f64(np.broadcast_to(np.asarray(v1, dtype=np.float64), (K,)) if not isinstance(v1, torch.Tensor) else v1, dev)
    v0 = _dev_f64(np.broadcast_to(np.asarray(v0, dtype=np.float64), (K,)) if not isinstance(v0, torch.Tensor) else v0, dev)
    if y.shape[0] != masks.n:
        raise ValueError(f"y_train has {y.shape[0]} rows, masks have {masks.n}")
    a_hat = masks.gram(0)  # datashapley.py:29
    b_hat = masks.xty(y, v0, 0.0, 1.0 / masks.n)  # datashapley.py:30
    a_inv = sym_pinv(a_hat, 1e-15)  # datashapley.py:37 (np.linalg.pinv default cut-off)
    h = masks.h
    colsum = torch.empty(masks.d + 1, dtype=_f64, device=dev)
    rhs = torch.empty_like(b_hat)
    with torch.cuda.device(dev):
        _lib.check(h.lib.gadm_shapley_rhs(h.ptr, a_inv.data_ptr(), b_hat.data_ptr(), masks.d, K, v1.data_ptr(),
                                         v0.data_ptr(), colsum.data_ptr(), rhs.data_ptr(), _lib.stream_ptr(dev)))
    coef = _dgemm(a_inv, rhs, zero_below=1e-10)  # datashapley.py:43-45
    return coef.cpu().numpy() if as_numpy else coef


def data_banzhaf_batched(x_train, y_train, device=None, as_numpy: bool = True):
    """KernelBanzhaf: min-norm least squares on the +-1/2 masks for all K behaviours -> [d, K]."""
    masks = x_train if isinstance(x_train, PackedMasks) else PackedMasks(x_train, device)
    dev = masks.device
    y = _dev_f64(y_train, dev)
    if y.dim() == 1:
        y = y[:, None]
    if y.shape[0] != masks.n:
        raise ValueError(f"y_train has {y.shape[0]} rows, masks have {masks.n}")
    a = masks.gram(1)  # databanzhaf.py:21
    rhs = masks.xty(y, None, 0.5, 1.0)  # databanzhaf.py:22
    a_inv = sym_pinv(a, np.finfo(np.float64).eps * masks.d)  # lstsq(rcond=None) cut-off (databanzhaf.py:20-25)
    coef = _dgemm(a_inv, rhs)
    return coef.cpu().numpy() if as_numpy else coef


def data_shapley(dataset_size, x_train, y_train, v1, v0):
    """Reference signature (note: v1 before v0).  Returns ndarray [d, 1] like the reference."""
    x_train = np.asarray(x_train)
    if dataset_size != x_train.shape[-1]:
        raise ValueError(f"dataset_size {dataset_size} != mask width {x_train.shape[-1]}")
    return data_shapley_batched(x_train, np.asarray(y_train, dtype=np.float64).reshape(-1, 1), [v1], [v0])


def data_banzhaf(x_train, y_train):
    """Reference signature.  Returns ndarray [d]."""
    return data_banzhaf_batched(x_train, np.asarray(y_train, dtype=np.float64).reshape(-1, 1))[:, 0]


def loo_attr_batched(train_masks, train_targets, full_targets, device=None, as_numpy: bool = True):
    """Leave-one-out attributions for all behaviours (lds.py:436-440):
    coeff[i, k] = sum_r (1 - X[r, i]) * (full[k] - y[r, k])  -> [d, K]."""
    masks = train_masks if isinstance(train_masks, PackedMasks) else PackedMasks(train_masks, device)
    y = _dev_f64(train_targets, masks.device)
    if y.dim() == 1:
        y = y[:, None]
    full = _dev_f64(np.broadcast_to(np.asarray(full_targets, dtype=np.float64).reshape(-1), (y.shape[1],)), masks.device)
    # X^T (y - full) = -sum_r X (full - y); sum_r (1 - X)(full - y) = sum_r (full - y) - sum_r X (full - y)
    xt = masks.xty(y, full, 0.0, 1.0)                       # sum_r X[r,i] (y - full)
    tot = (y - full[None, :]).sum(dim=0)                    # sum_r (y - full)   [K]
    coef = xt - tot[None, :]                                # = sum_r (1 - X)(full - y)
    return coef.cpu().numpy() if as_numpy else coef


def aoi_attr_batched(train_masks, train_targets, null_targets, device=None, as_numpy: bool = True):
    """Add-one-in attributions (lds.py:442-445): coeff[i, k] = sum_r X[r, i] * (y[r, k] - null[k]) -> [d, K]."""
    masks = train_masks if isinstance(train_masks, PackedMasks) else PackedMasks(train_masks, device)
    y = _dev_f64(train_targets, masks.device)
    if y.dim() == 1:
        y = y[:, None]
    null = _dev_f64(np.broadcast_to(np.asarray(null_targets, dtype=np.float64).reshape(-1), (y.shape[1],)), masks.device)
    coef = masks.xty(y, null, 0.0, 1.0)
    return coef.cpu().numpy() if as_numpy else coef


def spearman_matrix(x_test, y_test, attrs, idx=None, device=None, as_numpy: bool = True):
    """rho[e, k] = spearmanr(x_test[idx[e]] @ attrs[:, k], y_test[idx[e], k]); idx None -> one identity row."""
    masks = x_test if isinstance(x_test, PackedMasks) else PackedMasks(x_test, device)
    dev = masks.device
    attrs = _dev_f64(attrs, dev)
    y = _dev_f64(y_test, dev)
    pred = masks.times(attrs)
    K = attrs.shape[1]
    h = masks.h
    if idx is None:
        n_eval, rows = 1, masks.n
        idx_ptr = None
    else:
        idx_np = np.ascontiguousarray(np.asarray(idx, dtype=np.int64))
        if idx_np.size and (idx_np.min() < 0 or idx_np.max() >= masks.n):  # the kernel gathers rows without a bounds check
            raise IndexError(f"resampling indices must lie in [0, {masks.n}), got [{idx_np.min()}, {idx_np.max()}]")
        idx_t = torch.as_tensor(idx_np.astype(np.int32)).to(dev)
        if idx_t.dim() == 1:
            idx_t = idx_t[None, :]
        n_eval, rows = int(idx_t.shape[0]), int(idx_t.shape[1])
        idx_ptr = idx_t.data_ptr()
    rho = torch.empty(n_eval, K, dtype=_f64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(h.lib.gadm_lds_spearman(h.ptr, pred.data_ptr(), y.data_ptr(), masks.n, K, idx_ptr, n_eval, rows,
                                          rho.data_ptr(), _lib.stream_ptr(dev)))
    return rho.cpu().numpy() if as_numpy else rho


def lds_per_test_set(x_test, y_test, attrs, idx=None, device=None) -> np.ndarray:
    """mean_k rho * 100 for each evaluation row (the inner np.mean of evaluate_lds / lds.py:my_lds)."""
    rho = spearman_matrix(x_test, y_test, attrs, idx, device, as_numpy=False)
    h = _lib.get_handle(rho.device)
    out = torch.empty(rho.shape[0], dtype=_f64, device=rho.device)
    with torch.cuda.device(rho.device):
        _lib.check(h.lib.gadm_lds_mean(h.ptr, rho.data_ptr(), rho.shape[0], rho.shape[1], out.data_ptr(),
                                      _lib.stream_ptr(rho.device)))
    return out.cpu().numpy()


def evaluate_lds(attrs_all, test_data_list, num_model_behaviors, index_first: bool | None = None, device=None):
    """Reference signature.  ``attrs_all`` is [d, K] (shapley_lds.py) or a list / [K, d] array indexed by
    behaviour first (lds.py); ``index_first=None`` infers the convention like the callers use it."""
    if isinstance(attrs_all, (list, tuple)):
        attrs = np.stack([np.asarray(a, dtype=np.float64).reshape(-1) for a in attrs_all], axis=1)  # [d, K]
    else:
        attrs = np.asarray(attrs_all, dtype=np.float64)
        if index_first is True:
            attrs = attrs.reshape(attrs.shape[0], -1).T
    attrs = attrs[:, :num_model_behaviors]
    lds_list = []
    for (x_test, y_test) in test_data_list:
        y = np.asarray(y_test, dtype=np.float64)[:, :num_model_behaviors]
        lds_list.append(lds_per_test_set(x_test, y, attrs, None, device)[0])
    lds_mean = np.mean(lds_list)
    lds_ci = np.std(lds_list) / np.sqrt(len(lds_list)) * 1.96
    return lds_mean, lds_ci


def bootstrap_statistic(test_masks, test_targets, data_attr_list, device=None):
    """Vectorised replacement of ``my_lds`` (lds.py:460-471) for ``scipy.stats.bootstrap(..., vectorized=True)``:
    returns f(idx, axis=-1) evaluating every resampled index row in one kernel launch."""
    masks = PackedMasks(test_masks, device)
    attrs = np.stack([np.asarray(a, dtype=np.float64).reshape(-1) for a in data_attr_list], axis=1)
    y = _dev_f64(test_targets, masks.device)
    attrs_t = _dev_f64(attrs, masks.device)

    def statistic(idx, axis=-1):
        idx = np.asarray(idx)
        lead = idx.shape[:-1]
        flat = idx.reshape(-1, idx.shape[-1]).astype(np.int32)
        vals = lds_per_test_set(masks, y, attrs_t, flat)
        return vals.reshape(lead)

    return statistic


def group_reduce(values, group_ids, num_groups: int, mode: str = "sum", device=None, as_numpy: bool = True):
    """Per-group sum / mean / max in fp64 (traks.py:188-204)."""
    dev = values.device if isinstance(values, torch.Tensor) and values.is_cuda else _device(device)
    v = torch.as_tensor(values).to(dev).contiguous()
    if v.dtype not in (torch.float32, torch.float64):
        v = v.double()
    g = torch.as_tensor(np.asarray(group_ids) if not isinstance(group_ids, torch.Tensor) else group_ids)
    g = g.to(device=dev, dtype=torch.int32).contiguous()
    h = _lib.get_handle(dev)
    out = torch.empty(num_groups, dtype=_f64, device=dev)
    code = {"sum": 0, "mean": 1, "max": 2}[mode]
    with torch.cuda.device(dev):
        _lib.check(h.lib.gadm_group_reduce(h.ptr, v.data_ptr(), 0 if v.dtype == torch.float32 else 3, g.data_ptr(),
                                          v.numel(), num_groups, code, out.data_ptr(), _lib.stream_ptr(dev)))
    return out.cpu().numpy() if as_numpy else out


def stable_rank(output, device=None) -> np.ndarray:
    """np.argsort(-output.mean(axis=-1), kind="stable") (traks.py:218, shapley_lds.py:294), int64."""
    dev = _device(device)
    x = _dev_f64(output, dev)
    h = _lib.get_handle(dev)
    with torch.cuda.device(dev):
        if x.dim() == 2:
            m = torch.empty(x.shape[0], dtype=_f64, device=dev)
            _lib.check(h.lib.gadm_row_mean(h.ptr, x.data_ptr(), x.shape[0], x.shape[1], m.data_ptr(), _lib.stream_ptr(dev)))
            x = m
        rank = torch.empty(x.shape[0], dtype=torch.int64, device=dev)
        _lib.check(h.lib.gadm_stable_rank_desc(h.ptr, x.data_ptr(), x.shape[0], rank.data_ptr(), _lib.stream_ptr(dev)))
    return rank.cpu().numpy()


def lds_fit_sweep(train_masks, train_targets, test_data_list, subset_sizes, removal_dist: str, full_targets=None,
                  null_targets=None, train_indices=None, compat_loo_full_masks: bool = True, device=None):
    """The fit-size sweep of lds.py:397-456: for every ``n`` in ``subset_sizes`` fit all behaviours on the first ``n``
    rows of ``train_indices`` and evaluate the LDS on the three test sets.

    ``removal_dist`` in {"datamodel", "shapley", "uniform", "loo", "add_one_in"} selects the estimator exactly like
    ``args.removal_dist`` (RidgeCV / data_shapley / data_banzhaf / LOO / add-one-in).  The reference's LOO and
    add-one-in branches use the *full* ``train_masks`` instead of the fold (lds.py:438,444); that is kept by default
    (``compat_loo_full_masks``).  Returns a list of dicts {n, lds_mean, lds_ci, coef [d, K]}."""
    from .datamodel import datamodel_ridge_batched

    x = np.asarray(train_masks)
    y = np.asarray(train_targets, dtype=np.float64)
    if y.ndim == 1:
        y = y[:, None]
    K = y.shape[1]
    idx = np.arange(x.shape[0]) if train_indices is None else np.asarray(train_indices)
    out = []
    for n in subset_sizes:
        fold = idx[:n]
        xf, yf = x[fold], y[fold]
        if removal_dist == "datamodel":
            coef = datamodel_ridge_batched(xf, yf, device=device)
        elif removal_dist == "shapley":
            coef = data_shapley_batched(xf, yf, np.asarray(full_targets, dtype=np.float64).reshape(-1)[:K],
                                        np.asarray(null_targets, dtype=np.float64).reshape(-1)[:K], device=device)
        elif removal_dist == "uniform":
            coef = data_banzhaf_batched(xf, yf, device=device)
        elif removal_dist == "loo":
            xs, ys = (x, y) if compat_loo_full_masks else (xf, yf)
            coef = loo_attr_batched(xs, ys, full_targets, device=device)
        elif removal_dist == "add_one_in":
            xs, ys = (x, y) if compat_loo_full_masks else (xf, yf)
            coef = aoi_attr_batched(xs, ys, null_targets, device=device)
        else:
            raise ValueError(f"Removal distribution: {removal_dist} does not exist.")  # lds.py:447-450
        lds_mean, lds_ci = evaluate_lds(coef, test_data_list, K, device=device)
        out.append({"n": int(len(fold)), "lds_mean": float(lds_mean), "lds_ci": float(lds_ci), "coef": coef})
    return out


def convergence_metrics(baseline_attrs, attrs, device=None):
    """MSE, Pearson and Spearman between two attribution vectors (text_to_image/shapley_convergence.py:261-268).
    The rank correlation runs on the device kernel (scipy tie semantics); MSE / Pearson are d-element host sums."""
    a = np.asarray(baseline_attrs, dtype=np.float64).reshape(-1)
    b = np.asarray(attrs, dtype=np.float64).reshape(-1)
    if a.shape != b.shape:
        raise ValueError("attribution vectors must have the same length")
    dev = _device(device)
    h = _lib.get_handle(dev)
    ta, tb = _dev_f64(a[:, None], dev), _dev_f64(b[:, None], dev)
    rho = torch.empty(1, 1, dtype=_f64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(h.lib.gadm_lds_spearman(h.ptr, ta.data_ptr(), tb.data_ptr(), a.size, 1, None, 1, a.size, rho.data_ptr(),
                                          _lib.stream_ptr(dev)))
    am, bm = a - a.mean(), b - b.mean()
    pearson = float((am @ bm) / np.sqrt((am @ am) * (bm @ bm)))
    return {"mse": float(((a - b) ** 2).mean()), "pearson": pearson, "spearman": float(rho.item())}
