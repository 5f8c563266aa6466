"""TRAK / D-TRAK / influence scoring on sm_100a kernels, behind the reference's interfaces.

Reference arithmetic (all restated in oracle/scorer.py and pinned on golden vectors):
  text_to_image/traks.py:141-186         grad_sim, TRAK, relative / renormalised influence, journey-TRAK, D-TRAK
  src/attributions/methods/compute_gradient_score.py:13-139    numpy / fp64 version with the kernel cache
  unconditional_generation/attribute.py:15,150-152             imports a missing compute_dtrak_trak_scores

What changes under the hood: K = Phi^T Phi + lam*I is factored once by a blocked Cholesky (tensor-core
3xTF32 trailing updates) and applied to right-hand-side rows by blocked triangular solves; the explicit
inverse of the reference (torch.inverse / np.linalg.inv) is only formed when a caller asks for it
(kernel cache compatibility).  With ``torch.distributed`` initialised, training examples are sharded by
row: one all-reduce of the Gram matrix, identical factorisation on every rank, local score slices
[T, N/R], optional all-gather.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Iterable, Sequence

import numpy as np
import torch

from . import _lib
from .aggregation import group_reduce, stable_rank
from .distributed import allgather_cat, allreduce_sum_

_f32 = torch.float32


def _check_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise ValueError(f"{name} must live on a CUDA device (sm_100a); there is no CPU path")
    if t.dtype != _f32:
        t = t.float()
    if t.dim() != 2:
        raise ValueError(f"{name} must be 2-D, got {tuple(t.shape)}")
    if t.stride(1) != 1 or t.stride(0) % 4 != 0 or t.data_ptr() % 16 != 0:
        cols = t.shape[1]
        ld = -(-cols // 4) * 4
        buf = torch.zeros(t.shape[0], ld, dtype=_f32, device=t.device)
        buf[:, :cols] = t
        t = buf[:, :cols]
    return t


def _clone_padded(t: torch.Tensor) -> torch.Tensor:
    """Copy of a 2-D fp32 tensor whose row pitch is a multiple of 4 floats (16 B: TMA / float4 requirement)."""
    ld = -(-t.shape[1] // 4) * 4
    buf = torch.zeros(t.shape[0], ld, dtype=_f32, device=t.device) if ld != t.shape[1] else \
        torch.empty(t.shape[0], ld, dtype=_f32, device=t.device)
    out = buf[:, :t.shape[1]]
    out.copy_(t)
    return out


def _h(t: torch.Tensor):
    return _lib.get_handle(t.device)


def gemm_tn(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor | None = None, alpha: float = 1.0, beta: float = 0.0,
            diag_add: float = 0.0, lower_only: bool = False, b_tri: str | None = None) -> torch.Tensor:
    """out = alpha * a @ b.T + beta * out  (fp32-grade accuracy, 3xTF32 tcgen05 kernel).
    ``b_tri`` = "lower" / "upper": b is square triangular, the contraction skips its zero blocks."""
    a = _check_cuda_f32(a, "a")
    b = _check_cuda_f32(b, "b")
    if a.shape[1] != b.shape[1]:
        raise ValueError(f"contraction mismatch: {tuple(a.shape)} x {tuple(b.shape)}^T")
    m, n, k = a.shape[0], b.shape[0], a.shape[1]
    if out is None:  # row pitch padded to a multiple of 4 floats so the result is TMA-loadable as an operand later
        ld = -(-n // 4) * 4
        out = (torch.zeros(m, ld, dtype=_f32, device=a.device) if lower_only or ld != n
               else torch.empty(m, ld, dtype=_f32, device=a.device))[:, :n]
    h = _h(a)
    with torch.cuda.device(a.device):
        _lib.check(h.lib.gadm_gemm_tn(h.ptr, a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), out.data_ptr(),
                                     out.stride(0), m, n, k, float(alpha), float(beta), float(diag_add),
                                     int(lower_only) | {None: 0, "lower": 2, "upper": 4}[b_tri], _lib.stream_ptr(a.device)))
    return out


def transpose(x: torch.Tensor, min_ld: int = 0) -> torch.Tensor:
    """[R, C] -> [C, R] with a pitch that is a multiple of 4 floats (TMA-loadable) and at least ``min_ld``; the pad
    columns are zero."""
    x = _check_cuda_f32(x, "x")
    r, c = x.shape
    ld = -(-max(r, min_ld) // 4) * 4
    out = torch.empty(c, ld, dtype=_f32, device=x.device)
    if ld != r:
        out[:, r:].zero_()  # only the <= 3 pad columns (they enter the contraction of the Gram GEMM); a full zero fill
                            # of the 819 MB config-2 buffer cost 0.25 ms of the 8.7 ms score
    h = _h(x)
    with torch.cuda.device(x.device):
        _lib.check(h.lib.gadm_transpose(h.ptr, x.data_ptr(), r, c, x.stride(0), out.data_ptr(), ld,
                                       _lib.stream_ptr(x.device)))
    return out[:, :r]


def row_norms(x: torch.Tensor, reciprocal: bool = False) -> torch.Tensor:
    x = _check_cuda_f32(x, "x")
    out = torch.empty(x.shape[0], dtype=_f32, device=x.device)
    h = _h(x)
    with torch.cuda.device(x.device):
        _lib.check(h.lib.gadm_row_norms(h.ptr, x.data_ptr(), x.shape[0], x.shape[1], x.stride(0), int(reciprocal),
                                       out.data_ptr(), _lib.stream_ptr(x.device)))
    return out


def col_mean_scaled(s: torch.Tensor, row_scale: torch.Tensor | None = None,
                    col_scale: torch.Tensor | None = None) -> torch.Tensor:
    out = torch.empty(s.shape[1], dtype=_f32, device=s.device)
    h = _h(s)
    with torch.cuda.device(s.device):
        _lib.check(h.lib.gadm_col_mean_scaled(h.ptr, s.data_ptr(), s.shape[0], s.shape[1], s.stride(0),
                                             row_scale.data_ptr() if row_scale is not None else None,
                                             col_scale.data_ptr() if col_scale is not None else None,
                                             out.data_ptr(), _lib.stream_ptr(s.device)))
    return out


def scale_rows_cols_(s: torch.Tensor, row_scale: torch.Tensor | None = None, col_scale: torch.Tensor | None = None):
    h = _h(s)
    with torch.cuda.device(s.device):
        _lib.check(h.lib.gadm_scale_rows_cols(h.ptr, s.data_ptr(), s.shape[0], s.shape[1], s.stride(0),
                                             row_scale.data_ptr() if row_scale is not None else None,
                                             col_scale.data_ptr() if col_scale is not None else None,
                                             _lib.stream_ptr(s.device)))
    return s


def _gram_plan(n: int, k: int, device) -> tuple[int, int, int] | None:
    """Wave balance of the lower-triangle Gram GEMM.  With the single-CTA kernel (one 128 x 128 tile per SM) k = 4096
    gives 528 output tiles = 3.57 waves on 148 SMs and the fourth wave runs 57 % empty; with the CTA-pair kernel
    (256 x 256 tiles, 74 pairs) it is 136 tiles = 1.84 waves, where slicing does not pay (k = 8192: 7.13 waves, it does).
    Returns (r0, parts, kc): row tiles [0, r0) fill whole waves with full-contraction CTAs, the remaining row tiles are
    computed as ``parts`` contraction slices of kc columns each (short CTAs that pack into the last wave) and summed;
    None when that does not pay."""
    # long contractions run on the CTA-pair kernel: 256 x 256 tiles, one per pair of SMs (gadm.cu, GADM_GEMM_2CTA)
    pair = os.environ.get("GADM_GEMM_2CTA", "1") != "0" and n >= 2048 and k > 128
    tile = 256 if pair else 128
    nt = -(-k // tile)
    sms = torch.cuda.get_device_properties(device).multi_processor_count // (2 if pair else 1)
    tiles = nt * (nt + 1) // 2
    full = tiles // sms
    if k % tile or full < 1 or n < 8192:
        return None
    r0 = int((math.isqrt(8 * full * sms + 1) - 1) // 2)
    if r0 >= nt or r0 < 1:
        return None
    balanced = (r0 * (r0 + 1) // 2 + (nt - r0) * nt) / sms
    if balanced > 0.95 * -(-tiles // sms):
        return None
    parts = min(8, n // 2048)
    kc = -(-n // parts // 32) * 32
    return r0 * (tile // 128), parts, kc  # r0 in 128-row tiles


_SIDE_STREAMS: dict = {}


def gram_lower(phi: torch.Tensor, diag_add: float = 0.0) -> torch.Tensor:
    """phi [N, k] -> Phi^T Phi + diag_add * I  [k, k], lower block triangle valid (what gadm_cholesky reads);
    traks.py:149-151.  Transposes (contraction over examples becomes K-major) and runs the 3xTF32 GEMM, wave-balanced
    per ``_gram_plan``."""
    phi = _check_cuda_f32(phi, "phi")
    n, k = phi.shape
    plan = _gram_plan(n, k, phi.device)
    if plan is None:
        phi_t = transpose(phi)
        return gemm_tn(phi_t, phi_t, lower_only=True, diag_add=diag_add)
    r0, parts, kc = plan
    m0 = r0 * 128
    phi_t = transpose(phi, min_ld=parts * kc)
    ld = phi_t.stride(0)
    gram = torch.zeros(k, k, dtype=_f32, device=phi.device)
    partial = torch.empty(parts, k - m0, k, dtype=_f32, device=phi.device)
    h = _h(phi)
    main = torch.cuda.current_stream(phi.device)
    side = _SIDE_STREAMS.get(phi.device)
    if side is None:
        side = _SIDE_STREAMS[phi.device] = torch.cuda.Stream(device=phi.device, priority=-1)
    side.wait_stream(main)
    with torch.cuda.device(phi.device):
        with torch.cuda.stream(side):  # the short CTAs first and with priority: they fill the gaps of the main grid's waves
            _lib.check(h.lib.gadm_gemm_tn_batched(h.ptr, phi_t[m0:].data_ptr(), ld, kc, phi_t.data_ptr(), ld, kc,
                                                 partial.data_ptr(), k, (k - m0) * k, k - m0, k, kc, parts, 1.0, 0.0, 0.0, 0,
                                                 _lib.stream_ptr(phi.device)))
            torch.sum(partial, dim=0, out=gram[m0:])
            if diag_add:
                gram.diagonal()[m0:].add_(diag_add)
        gemm_tn(phi_t[:m0], phi_t[:m0], out=gram[:m0, :m0], lower_only=True, diag_add=diag_add)
    main.wait_stream(side)
    for t in (phi_t, partial, gram):
        t.record_stream(side)
    return gram


LOCAL = "local"  # pass as ``group`` to score this process's rows alone even when torch.distributed is initialised


def _dist(group=None):
    import torch.distributed as dist

    if group is LOCAL:
        return None
    return dist if (dist.is_available() and dist.is_initialized()) else None


def _grp(group):
    return None if group is LOCAL else group


# (max pivot / min pivot)^2 of the Cholesky factor is a lower bound of cond(K).  fp32 resolves 2^-24: beyond ~1e7 the
# factor of an fp32 matrix carries no correct digits in the small-eigenvalue directions, so results are refused
# instead of returned (the reference inverts in fp64 on its numpy path, compute_gradient_score.py:108-110).
MAX_FP32_PIVOT_COND = 1.0e7


def matvec_rows(x: torch.Tensor, v: torch.Tensor, col_scale: torch.Tensor | None = None) -> torch.Tensor:
    """out[n] = col_scale[n] * <x[n, :], v>  (HBM-bound GEMV, fp64 accumulation)."""
    x = _check_cuda_f32(x, "x")
    v = v.reshape(-1).contiguous().float()
    if v.numel() != x.shape[1]:
        raise ValueError(f"vector has {v.numel()} entries, rows have {x.shape[1]}")
    out = torch.empty(x.shape[0], dtype=_f32, device=x.device)
    h = _h(x)
    with torch.cuda.device(x.device):
        _lib.check(h.lib.gadm_matvec_rows(h.ptr, x.data_ptr(), x.shape[0], x.shape[1], x.stride(0), v.data_ptr(),
                                         col_scale.data_ptr() if col_scale is not None else None, out.data_ptr(),
                                         _lib.stream_ptr(x.device)))
    return out


class TrakScorer:
    """K = Phi^T Phi + lam*I factored once; rows are then multiplied by K^-1 on demand.

    When there are fewer training examples than projection dimensions (BASELINE configs 3 and 4: N = 5000 with
    k = 8192 / 32768) the k x k system is replaced by the N x N one through the push-through identity
    ``K^-1 Phi^T = Phi^T (Phi Phi^T + lam*I)^-1`` (exact algebra, no cancellation), i.e.
    ``S = (Phi_gen Phi^T) A^-1`` with ``A = Phi Phi^T + lam*I``: 2 N^2 k + N^3/3 flops instead of 2 N k^2 + k^3/3
    (15x less at config 4).  ``dual`` is chosen automatically; the primal path is what the reference computes
    literally (traks.py:149-156)."""

    def __init__(self, lam: float = 5e-1, group=None):
        self.lam = float(lam)
        self.group = group
        self.k = None          # size of the factored system (k primal, N_total dual)
        self.L = None
        self.U = None
        self._X = None         # L^-1 (lower) and
        self._Xt = None        # L^-T (upper), explicit -- built on first use (``X`` / ``Xt``)
        self.blocks = None
        self.info = None
        self.dual = False
        self.phi_all = None    # dual: every rank's training features [N_total, k]
        self.local = None      # dual: [lo, hi) of this rank's examples inside phi_all

    def fit(self, train_phi: torch.Tensor, dual: bool | None = None) -> "TrakScorer":
        """train_phi: this rank's [N_local, k] features.  traks.py:149-151 / compute_gradient_score.py:108-110."""
        phi = _check_cuda_f32(train_phi, "train_phi")
        dist = _dist(self.group)
        world = dist.get_world_size(self.group) if dist else 1
        n_local = phi.shape[0]
        if world > 1:
            sizes = [torch.zeros(1, dtype=torch.int64, device=phi.device) for _ in range(world)]
            dist.all_gather(sizes, torch.tensor([n_local], dtype=torch.int64, device=phi.device), group=self.group)
            lens = [int(x.item()) for x in sizes]
            rank = dist.get_rank(self.group)
        else:
            lens, rank = [n_local], 0
        n_total = sum(lens)
        self.dual = (n_total < phi.shape[1]) if dual is None else bool(dual)
        if self.dual:
            self.phi_all = _check_cuda_f32(phi if dist is None else allgather_cat(phi, dim=0, group=self.group), "train_phi")
            lo = sum(lens[:rank])
            self.local = (lo, lo + n_local)
            if dist is not None and world > 1 and min(lens) > 0:
                # every rank builds its own block of rows A[lo:hi, :] = Phi_r Phi_all^T and the blocks are all-gathered
                # (N x N floats) instead of every rank building all of A: 1/R of the 2 N^2 k flops per GPU.  The blocks
                # are bitwise what the unsharded GEMM computes (same kernel, same contraction order per element).
                rows = gemm_tn(phi, self.phi_all)
                gram = _clone_padded(allgather_cat(rows, dim=0, group=self.group))
                gram.diagonal().add_(self.lam)
            else:
                gram = gemm_tn(self.phi_all, self.phi_all, lower_only=True, diag_add=self.lam)  # A = Phi Phi^T + lam I
            return self.factor_(gram)
        gram = gram_lower(phi, self.lam / world)  # transposes: the contraction over examples becomes K-major
        if dist is not None:
            allreduce_sum_(gram, self.group)  # one NCCL all-reduce over NVLink (sum of per-rank Grams)
        return self.factor_(gram)

    def partial_fit(self, phi_rows: torch.Tensor) -> "TrakScorer":
        """Streaming form of ``fit`` for the featurisation loop: accumulate Phi^T Phi batch by batch (e.g. every 1024
        rows that ``CudaProjector.deferred()`` produces) so that the Gram GEMM hides behind the gradient computation.
        Call ``finalize()`` once all of this rank's rows have been seen (primal k x k form only)."""
        phi = _check_cuda_f32(phi_rows, "phi_rows")
        phi_t = transpose(phi)
        if getattr(self, "_gram_acc", None) is None:
            self._gram_acc = gemm_tn(phi_t, phi_t, lower_only=True)
        else:
            gemm_tn(phi_t, phi_t, out=self._gram_acc, beta=1.0, lower_only=True)
        return self

    def finalize(self) -> "TrakScorer":
        """All-reduce the accumulated Gram, add lam*I and factor (the tail of ``fit``)."""
        if getattr(self, "_gram_acc", None) is None:
            raise RuntimeError("finalize() before any partial_fit()")
        gram, self._gram_acc = self._gram_acc, None
        if _dist(self.group) is not None:
            allreduce_sum_(gram, self.group)
        gram.diagonal().add_(self.lam)
        self.dual = False
        return self.factor_(gram)

    def factor_(self, gram: torch.Tensor) -> "TrakScorer":
        """In-place Cholesky of an already regularised symmetric matrix (lower triangle is read).  The explicit
        triangular inverse that the many-row solves use is built on first use: the mean-first score paths solve a
        single row and go through ``gadm_cholesky_solve_vec`` instead."""
        self._X = self._Xt = None
        return self._cholesky_(gram)

    @property
    def X(self) -> torch.Tensor:
        if self._X is None:
            self._tri_inverse_()
        return self._X

    @property
    def Xt(self) -> torch.Tensor:
        if self._Xt is None:
            self._tri_inverse_()
        return self._Xt

    def _solve_vec(self, row: torch.Tensor) -> torch.Tensor:
        """One row times (factored matrix)^-1 by forward / backward substitution in one cooperative launch."""
        gram = self.L
        h = _h(gram)
        b = row.reshape(-1).contiguous().float()
        x = torch.empty(self.k, dtype=_f32, device=gram.device)
        ws = torch.empty(int(h.lib.gadm_cholesky_solve_vec_workspace_bytes(self.k)), dtype=torch.uint8, device=gram.device)
        with torch.cuda.device(gram.device):
            _lib.check(h.lib.gadm_cholesky_solve_vec(h.ptr, gram.data_ptr(), gram.stride(0), self.blocks.data_ptr(), self.k,
                                                    b.data_ptr(), x.data_ptr(), ws.data_ptr(), ws.numel(),
                                                    _lib.stream_ptr(gram.device)))
        return x

    def _cholesky_(self, gram: torch.Tensor) -> "TrakScorer":
        self.k = gram.shape[0]
        h = _h(gram)
        nbytes = int(h.lib.gadm_cholesky_workspace_bytes(self.k))
        self.blocks = torch.empty(nbytes, dtype=torch.uint8, device=gram.device)
        self.info = torch.zeros(1, dtype=torch.int32, device=gram.device)
        with torch.cuda.device(gram.device):
            _lib.check(h.lib.gadm_cholesky(h.ptr, gram.data_ptr(), gram.stride(0), self.k, self.blocks.data_ptr(), nbytes,
                                          C.cast(self.info.data_ptr(), C.POINTER(C.c_int)), _lib.stream_ptr(gram.device)))
            self.pivots = torch.empty(2, dtype=_f32, device=gram.device)
            _lib.check(h.lib.gadm_diag_minmax(h.ptr, gram.data_ptr(), gram.stride(0), self.k, self.pivots.data_ptr(),
                                             _lib.stream_ptr(gram.device)))
        self.L = gram
        self.U = None
        self._checked = False
        return self

    def _tri_inverse_(self) -> "TrakScorer":
        """Explicit L^-1 / L^-T by recursive doubling: K^-1 is then applied by two full-size GEMMs."""
        gram = self.L
        h = _h(gram)
        ld = -(-self.k // 4) * 4
        self._X = torch.empty(self.k, ld, dtype=_f32, device=gram.device)[:, :self.k]
        self._Xt = torch.empty(self.k, ld, dtype=_f32, device=gram.device)[:, :self.k]
        ws = torch.empty(int(h.lib.gadm_tri_inverse_workspace_bytes(self.k)), dtype=torch.uint8, device=gram.device)
        with torch.cuda.device(gram.device):
            _lib.check(h.lib.gadm_tri_inverse(h.ptr, gram.data_ptr(), gram.stride(0), self.blocks.data_ptr(), self.k,
                                              self._X.data_ptr(), self._X.stride(0), self._Xt.data_ptr(), self._Xt.stride(0),
                                              ws.data_ptr(), ws.numel(), _lib.stream_ptr(gram.device)))
        return self

    def check(self, max_cond: float = MAX_FP32_PIVOT_COND) -> None:
        """Host sync (one small D2H, cached): raise if the factorisation met a non-positive / NaN pivot (the kernel
        substitutes 1 and carries on, so everything downstream would be finite garbage) or if the pivot range shows
        that fp32 cannot resolve the matrix."""
        if getattr(self, "_checked", False):
            return
        bad = int(self.info.item())
        if bad:
            raise _lib.GadmError(f"Gram matrix + lam*I is not positive definite in fp32 (pivot {bad - 1}): features "
                                 "contain NaN / Inf or the system is too ill-conditioned for the fp32 factorisation")
        lo, hi = (float(x) for x in self.pivots.tolist())
        if not (lo > 0.0) or not (hi < float("inf")):
            raise _lib.GadmError(f"Cholesky factor has a non-finite or non-positive pivot range [{lo}, {hi}]")
        if (hi / lo) ** 2 > max_cond:
            raise _lib.GadmError(f"system is too ill-conditioned for fp32 (pivot-ratio bound cond >= {(hi / lo) ** 2:.3g} > "
                                 f"{max_cond:.3g}); rescale the features or increase lam")
        self._checked = True

    def _solve(self, rows: torch.Tensor, inplace: bool = False) -> torch.Tensor:
        """rows [m, self.k] -> rows @ (factored matrix)^-1 = (rows L^-T) L^-1: two GEMMs against the explicit
        triangular inverse (``inplace`` is accepted for compatibility; a new tensor is returned)."""
        y = _check_cuda_f32(rows, "rows")
        if y.shape[1] != self.k:
            raise ValueError(f"rows have {y.shape[1]} columns, the factored system has {self.k}")
        if y.shape[0] <= 2:
            # the mean-first score paths solve ONE row.  Without the explicit inverse (not built yet): substitution
            # through the factor in one cooperative launch; with it: two matrix-vector products (HBM-bound, fp64
            # accumulation, ~20 us each at k = 4096) instead of two GEMMs that run a single 128-row tile through the
            # whole contraction (84 us each)
            if self._X is None and self.k <= 8192 and self.L.stride(0) % 4 == 0 and self.L.data_ptr() % 16 == 0:
                return torch.stack([self._solve_vec(y[i]) for i in range(y.shape[0])])
            return torch.stack([matvec_rows(self.Xt, matvec_rows(self.X, y[i])) for i in range(y.shape[0])])
        return gemm_tn(gemm_tn(y, self.X, b_tri="lower"), self.Xt, b_tri="upper")

    def solve_rows_blocked(self, rows: torch.Tensor) -> torch.Tensor:
        """The same through blocked forward / backward substitution (gadm_solve_rows), kept as a cross-check."""
        y = _clone_padded(_check_cuda_f32(rows, "rows"))
        if self.U is None:
            self.U = transpose(self.L)
        h = _h(y)
        with torch.cuda.device(y.device):
            _lib.check(h.lib.gadm_solve_rows(h.ptr, self.L.data_ptr(), self.L.stride(0), self.U.data_ptr(), self.U.stride(0),
                                            self.blocks.data_ptr(), self.k, y.data_ptr(), y.stride(0), y.shape[0],
                                            _lib.stream_ptr(y.device)))
        return y

    def solve_rows(self, rows: torch.Tensor, inplace: bool = False) -> torch.Tensor:
        """rows [m, k] -> rows @ K^-1."""
        if self.dual:
            # Woodbury: K^-1 = (I - Phi^T A^-1 Phi) / lam  (normwise accurate; the score paths below never need it)
            r = _check_cuda_f32(rows, "rows")
            z = self._solve(gemm_tn(r, self.phi_all), inplace=True)          # (rows Phi^T) A^-1   [m, N]
            out = _clone_padded(r)
            return gemm_tn(z, transpose(self.phi_all), out=out, alpha=-1.0 / self.lam, beta=1.0 / self.lam)
        return self._solve(rows, inplace)

    def kernel_inverse(self) -> torch.Tensor:
        """Explicit K^-1 (what the reference caches as kernel_*.npy, compute_gradient_score.py:104-111)."""
        if not self.dual:
            return gemm_tn(self.Xt, self.Xt, b_tri="upper")  # K^-1 = L^-T L^-1
        eye = torch.eye(self.phi_all.shape[1], dtype=_f32, device=self.L.device)
        return self.solve_rows(eye, inplace=True)

    def score_matrix(self, gen_phi: torch.Tensor, train_phi: torch.Tensor) -> torch.Tensor:
        """S = gen_phi K^-1 train_phi^T  [T, N_local]  (traks.py:152-156; compute_gradient_score.py:126)."""
        if self.dual:
            z = self._solve(gemm_tn(_check_cuda_f32(gen_phi, "gen_phi"), self.phi_all), inplace=True)  # [T, N_total]
            return z[:, self.local[0]:self.local[1]]
        z = self.solve_rows(gen_phi)
        return gemm_tn(z, _check_cuda_f32(train_phi, "train_phi"))

    def train_weight_norms(self, train_phi: torch.Tensor, reciprocal: bool = True) -> torch.Tensor:
        """(1 /) ||K^-1 phi_n|| for this rank's training examples (traks.py:162, compute_gradient_score.py:120)."""
        if self.dual:
            wt = self._solve(transpose(self.phi_all), inplace=True)          # Phi^T A^-1 = K^-1 Phi^T   [k, N_total]
            return row_norms(transpose(wt[:, self.local[0]:self.local[1]]), reciprocal)
        return row_norms(self.solve_rows(train_phi), reciprocal)


TRAK_VARIANTS = ("grad_sim", "trak", "relative_influence", "renorm_influence")


def trak_scores(train_phi: torch.Tensor, gen_phi: torch.Tensor, lam: float = 5e-1,
                variants: Sequence[str] = TRAK_VARIANTS, journey_phi: torch.Tensor | None = None,
                group=None, gather: bool = True, return_scorer: bool = False, dual: bool | None = None,
                check: bool = True):
    """Per-training-example attribution vectors of text_to_image/traks.py:139-173 (mean over generated images).

    Returns dict name -> fp32 tensor [N] (all ranks' examples when ``gather`` and torch.distributed is up).
    ``journey_trak`` is added when ``journey_phi`` is given.  D-TRAK (traks.py:176-186) is the same call on the
    mean-squared-l2-norm features.

    Every variant is a mean over generated images of something linear in the generated features, so the mean is
    taken FIRST: ``mean_t(gen_t) K^-1 Phi^T`` is one solved row and a matrix-vector product over the training
    features instead of the reference's [T, N] GEMM followed by ``.mean(dim=0)`` (traks.py:156-157); the [T, N]
    matrix itself is still available from ``gradient_scores`` / ``TrakScorer.score_matrix``.  ``check`` raises (one
    small host sync) when the factorisation failed or is beyond fp32 instead of returning finite garbage."""
    train = _check_cuda_f32(train_phi, "train_phi")
    gen = _check_cuda_f32(gen_phi, "gen_phi")
    out = {}
    inv_train_norm = None
    if "grad_sim" in variants or "renorm_influence" in variants:
        inv_train_norm = row_norms(train, reciprocal=True)
    if "grad_sim" in variants:  # traks.py:141-146: mean_t cos(gen_t, phi_n) = <mean_t gen_t / |gen_t|, phi_n> / |phi_n|
        out["grad_sim"] = matvec_rows(train, col_mean_scaled(gen, row_norms(gen, reciprocal=True)), inv_train_norm)
    scorer = None
    if any(v in variants for v in ("trak", "relative_influence", "renorm_influence")) or journey_phi is not None:
        scorer = TrakScorer(lam, group).fit(train, dual=dual)  # dual=None: N x N system when N_total < k

        def mean_scores(rows_phi):
            """mean_t (rows_t K^-1 phi_n) for this rank's examples, unscaled."""
            bar = col_mean_scaled(rows_phi)  # [k]
            if scorer.dual:
                y = matvec_rows(scorer.phi_all, bar)  # (mean gen) Phi^T   [N_total]
                z = scorer._solve(y[None, :])[0]      # ... A^-1
                return z[scorer.local[0]:scorer.local[1]].contiguous()
            return matvec_rows(train, scorer._solve(bar[None, :])[0])

        need = [v for v in ("trak", "relative_influence", "renorm_influence") if v in variants]
        if need:
            m = mean_scores(gen)
            if "trak" in variants:
                out["trak"] = m  # traks.py:156-157
            if "relative_influence" in variants:  # traks.py:161-164: / ||K^-1 phi_n||
                out["relative_influence"] = _scaled(m, scorer.train_weight_norms(train))
            if "renorm_influence" in variants:  # traks.py:166-168: / ||phi_n||
                out["renorm_influence"] = _scaled(m, inv_train_norm)
        if journey_phi is not None:  # traks.py:171-173
            out["journey_trak"] = mean_scores(_check_cuda_f32(journey_phi, "journey_phi"))
        if check:
            scorer.check()
    if gather and _dist(group) is not None:
        for name, v in list(out.items()):
            out[name] = allgather_cat(v, dim=0, group=group)  # one all-gather of the per-example score slices
    return (out, scorer) if return_scorer else out


def _scaled(v: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """v[n] * scale[n] through the row/column scaling kernel (keeps elementwise torch math off the product path)."""
    out = v.clone().reshape(1, -1)
    return scale_rows_cols_(out, None, scale)[0]


def group_and_rank(sample_output_dict: dict, group_ids, num_groups: int):
    """traks.py:188-225: per-group aggregation (sum; grad_sim -> avg / max) and stable descending ranks.

    Returns (output_dict name -> float64 [G, 1], rank_dict name -> int64 [G])."""
    output_dict = {}
    for method, attrs in sample_output_dict.items():
        if method in ["grad_sim"]:
            output_dict[f"avg_{method}"] = group_reduce(attrs, group_ids, num_groups, "mean").reshape(num_groups, 1)
            output_dict[f"max_{method}"] = group_reduce(attrs, group_ids, num_groups, "max").reshape(num_groups, 1)
        else:
            output_dict[method] = group_reduce(attrs, group_ids, num_groups, "sum").reshape(num_groups, 1)
    rank_dict = {name: stable_rank(output) for name, output in output_dict.items()}
    return output_dict, rank_dict


def aggregate_by_class(scores, labels, by: str = "mean", compat_max_over_all_rows: bool = True):
    """src/attributions/methods/attribution_utils.py:15-48 with the dataset replaced by its label vector.

    ``by="max"`` in the reference takes the max over *all* rows (attribution_utils.py:46); that behaviour is
    kept by default (``compat_max_over_all_rows``)."""
    scores = scores if isinstance(scores, torch.Tensor) else torch.as_tensor(np.asarray(scores))
    if scores.dim() == 1:
        scores = scores[None, :]
    labels = np.asarray(labels)
    uniq = sorted(set(labels.tolist()))
    lut = {v: i for i, v in enumerate(uniq)}
    gid = np.array([lut[v] for v in labels.tolist()], dtype=np.int32)
    rows = []
    counts = np.bincount(gid, minlength=len(uniq))
    for r in range(scores.shape[0]):
        if by == "mean":  # np.divide(scores[:, mask].sum(axis=1), np.sum(mask)): sum in the scores' dtype, division in fp64
            rows.append(group_reduce(scores[r].contiguous(), gid, len(uniq), "sum") / counts)
        else:
            rows.append(group_reduce(scores[r].contiguous(), gid, len(uniq), "max"))
    result = np.stack(rows)
    if by == "max" and compat_max_over_all_rows:
        result[:] = result.max(axis=0, keepdims=True)
    return result


def gradient_scores(train_phi: torch.Tensor, val_phi: torch.Tensor, gradient_type: str = "trak", lam: float = 5e-1,
                    kernel_inverse: torch.Tensor | None = None, check: bool = True):
    """Score matrix [T, N] of compute_gradient_score.py:108-126 for one gradient_type; returns (scores, scorer).
    ``check``: raise instead of returning finite garbage when the factorisation failed (TrakScorer.check)."""
    train = _check_cuda_f32(train_phi, "train_phi")
    val = _check_cuda_f32(val_phi, "val_phi")
    if gradient_type == "vanilla_gradient":  # :114-117
        s = gemm_tn(val, train)
        return scale_rows_cols_(s, row_norms(val, True), row_norms(train, True)), None
    scorer = None
    if kernel_inverse is not None:  # cached kernel (:102-105): scores = val @ (train @ kernel).T
        kinv = _check_cuda_f32(kernel_inverse, "kernel_inverse")
        w = gemm_tn(train, kinv)  # kernel is symmetric
    else:
        scorer = TrakScorer(lam).fit(train)
        w = None
    if scorer is not None and scorer.dual:  # N < k: every quantity through the N x N system
        if gradient_type == "relative_if":
            col = scorer.train_weight_norms(train)
        elif gradient_type == "renormalized_if":
            col = row_norms(train, True)
        else:
            col = None
        s = scorer.score_matrix(val, train).contiguous()
    else:
        if gradient_type == "relative_if":  # :119-120
            if w is None:
                w = scorer.solve_rows(train)
            col = row_norms(w, True)
        elif gradient_type == "renormalized_if":  # :121-122
            col = row_norms(train, True)
        else:
            col = None
        s = gemm_tn(val, w) if w is not None else scorer.score_matrix(val, train)
    if col is not None:
        scale_rows_cols_(s, None, col)
    if check and scorer is not None:
        scorer.check()
    return s, scorer


def _constants_outdir(outdir):
    if outdir is not None:
        return outdir
    if os.environ.get("GADM_OUTDIR"):
        return os.environ["GADM_OUTDIR"]
    try:
        import src.constants as constants  # the reference's user-created module (README.md:19-28)

        return constants.OUTDIR
    except Exception as e:  # pragma: no cover
        raise RuntimeError("pass outdir=..., set GADM_OUTDIR or provide src/constants.py (OUTDIR)") from e


def compute_gradient_scores(args, retraining: bool = False, training_seeds: Iterable[int] | None = None, *,
                            outdir: str | None = None, n_train: int | None = None, n_val: int | None = None,
                            labels=None, device=None, compat: bool = True):
    """Reference signature: src/attributions/methods/compute_gradient_score.py:13 (called at baseline_lds.py:380-383).

    Reads the reference's raw fp32 memmaps (``train_f=..._t=..._k=..._d=...`` / ``reference_f=...``), reuses or writes
    the fp64 ``kernel_train_*.npy`` cache, and returns what the reference returns: with ``compat=True`` the
    [T, N] score matrix unless ``args.by_class`` (the ``else: coeff = scores`` of :134-137), with ``compat=False``
    the mean over T when the behaviour is global.  ``n_train`` / ``n_val`` replace ``len(dataset)`` /
    ``len(ImageDataset(sample_dir))`` (dataset I/O is out of scope); ``labels`` feeds aggregate_by_class."""
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    outdir = _constants_outdir(outdir)
    if args.gradient_type == "d_trak":
        model_behavior, t_strategy = "mean-squared-l2-norm", "uniform"
    else:
        model_behavior, t_strategy = "loss", "uniform"
    if args.gradient_type == "journey_trak":
        val_grad_path = os.path.join(outdir, args.dataset, "d_trak", "full",
                                     f"gen_f=loss_t=uniform_k={args.k_partition}_d={args.projector_dim}")
    else:
        val_grad_path = os.path.join(args.sample_dir, "d_trak",
                                     f"reference_f={model_behavior}_t={t_strategy}_k={args.k_partition}_d={args.projector_dim}")
    kdim = args.projector_dim
    if n_val is None:
        n_val = args.sample_size if args.gradient_type == "journey_trak" else os.path.getsize(val_grad_path) // (4 * kdim)
    val_phi = np.memmap(val_grad_path, dtype=np.float32, mode="r", shape=(n_val, kdim))[: args.sample_size]
    val = torch.from_numpy(np.array(val_phi, dtype=np.float32, copy=True)).to(device)

    def _load_train(path):
        n = n_train if n_train is not None else os.path.getsize(path) // (4 * kdim)
        return torch.from_numpy(np.array(np.memmap(path, dtype=np.float32, mode="r", shape=(n, kdim)), copy=True)).to(device)

    if retraining:  # :54-79
        scores = None
        seeds = list(training_seeds)
        for seed in seeds:
            removal_dir = f"{args.removal_dist}/{args.removal_dist}_seed={seed}"
            # the reference spells this directory "d_track" (:63); accept both
            for sub in ("d_track", "d_trak"):
                path = os.path.join(outdir, args.dataset, sub, removal_dir,
                                    f"train_f={args.trak_behavior}_t={args.t_strategy}_k={args.k_partition}_d={kdim}")
                if os.path.exists(path):
                    break
            train = _load_train(path)
            s, _ = gradient_scores(train, val, "trak")
            scores = s / len(seeds) if scores is None else scores + s / len(seeds)
    else:
        train_grad_dir = os.path.join(outdir, args.dataset, "d_trak", "full")
        train_grad_path = os.path.join(train_grad_dir, f"train_f={model_behavior}_t={t_strategy}_k={args.k_partition}_d={kdim}")
        kernel_path = os.path.join(train_grad_dir,
                                   f"kernel_train_f={model_behavior}_t={t_strategy}_k={args.k_partition}_d={kdim}.npy")
        train = _load_train(train_grad_path)
        kinv = None
        if os.path.isfile(kernel_path):
            kinv = torch.from_numpy(np.load(kernel_path)).to(device=device, dtype=_f32)
        scores, scorer = gradient_scores(train, val, args.gradient_type, kernel_inverse=kinv)
        if kinv is None:
            # the reference builds and caches the kernel before it looks at gradient_type (:102-111)
            if scorer is None:
                scorer = TrakScorer(5e-1).fit(train)
            scorer.check()  # never cache the inverse of a factorisation that failed or is beyond fp32
            np.save(kernel_path, scorer.kernel_inverse().double().cpu().numpy())
    is_local = args.model_behavior_key in ["ssim", "nrmse", "diffusion_loss"]
    if getattr(args, "by_class", False):
        coeff = scores if is_local else col_mean_scaled(scores)
        if labels is None:
            from src.datasets import create_dataset  # reference module; only needed for the label vector

            labels = [d[1] for d in create_dataset(dataset_name=args.dataset, train=True)]
        return aggregate_by_class(coeff, labels, getattr(args, "by", "mean"))
    if compat or is_local:
        return scores.cpu().numpy()
    return col_mean_scaled(scores).cpu().numpy()


def compute_dtrak_trak_scores(args, train_idx=None, val_idx=None, **kw):
    """The callee unconditional_generation/attribute.py:15,150-152 imports but the reference never ships.

    ``args.attribution_method`` in {d-trak, trak, relative_if, randomized_if} selects the gradient type;
    ``train_idx`` restricts the returned columns (remaining training examples)."""
    method = getattr(args, "attribution_method", "trak")
    gtype = {"d-trak": "d_trak", "d_trak": "d_trak", "trak": "trak", "relative_if": "relative_if",
             "randomized_if": "renormalized_if", "renormalized_if": "renormalized_if"}[method]
    ns = type("Args", (), dict(vars(args)))()
    ns.gradient_type = gtype
    for name, default in (("by_class", False), ("k_partition", 10), ("model_behavior_key", "global")):
        if not hasattr(ns, name):
            setattr(ns, name, default)
    scores = compute_gradient_scores(ns, **kw)
    scores = np.asarray(scores)
    if scores.ndim == 2 and scores.shape[0] > 1 and not ns.by_class:
        scores = scores.mean(axis=0)
    if train_idx is not None and not ns.by_class:
        scores = scores.reshape(-1)[np.asarray(train_idx)]
    return scores
