"""Drop-in projectors with trak's call signatures, backed by the sm_100a JL projection kernel.

Mirrors ``trak.projectors`` (traker==0.1.3, reference ``requirements.txt:10``) as used at
``src/attributions/methods/d_trak_grad.py:14,504-511,776`` and
``text_to_image/grad_text_to_image_lora.py:68,561-568,765,813``::

    projector = CudaProjector(grad_dim=..., proj_dim=..., seed=..., proj_type=ProjectionType.normal,
                              device=device, max_batch_size=8)
    emb = projector.project(emb, model_id=0)          # [B, D] (or dict of per-parameter grads) -> [B, k]

Same semantics as upstream: P has i.i.d. N(0,1) (or +-1) entries, is never materialised, is a pure
function of ``seed + 10**4 * model_id`` and is not scaled by 1/sqrt(k).  Differences, all additive:

* ``project`` also accepts the *dict / list of per-parameter gradient tensors* that
  ``torch.func.vmap(grad(f))`` returns, so ``vectorize_and_ignore_buffers``
  (``d_trak_grad.py:188-226``) and its extra B*D*4-byte copy can be dropped;
* ``deferred()`` stages up to 512 (Rademacher) / 1024 (normal) examples (16-bit rows in HBM) and projects them
  in one pass, which is what makes the tensor cores, not the random-number generation, the bound;
  ``DeferredProjection.accumulate`` sums the K timestep gradients of a batch into an fp32 slab with the 1/K mean
  folded in (``d_trak_grad.py:764-770``) and stages the batch on the last timestep, and with ``overlap`` the
  projection pass of one staged batch runs on a side stream while the next batch is being staged.

Numerics that differ from an fp32-input projection and are stated here on purpose: the gradients are rounded to 16
bits before the tensor-core GEMM -- fp16 with a power-of-two scale per (example, 32768-column group) by default
(``stage_dtype="f16"``: 11 significant bits, relative feature error ~2e-4), or plain bf16 (``"bf16"``: 8 bits,
~1.6e-3) -- and the normal-type P is a bf16-rounded Box-Muller matrix.  Accumulation is fp32.

There is no CPU path: a non-CUDA device raises ``ValueError`` exactly like upstream's CudaProjector.
"""
from __future__ import annotations

import ctypes as C
import os
from enum import Enum
from typing import Mapping, Sequence, Union

import torch

from . import _lib

SEED_MODEL_ID_STRIDE = int(1e4)  # trak.projectors.CudaProjector.project: seed + int(1e4) * model_id
MAX_STAGE_ROWS = 512
TILE_K = 64

_DTYPES = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
_STAGE_CODES = {"bf16": 0, "f16": 1}          # GADM_STAGE_BF16 / GADM_STAGE_F16G (include/gadm.h)
_STAGE_TORCH = {"bf16": torch.bfloat16, "f16": torch.float16}

GradsLike = Union[torch.Tensor, Mapping[str, torch.Tensor], Sequence[torch.Tensor]]


class ProjectionType(str, Enum):
    """trak.projectors.ProjectionType."""

    normal = "normal"
    rademacher = "rademacher"


_PROJ_CODE = {ProjectionType.normal: 0, ProjectionType.rademacher: 1}


_STAGE_NARROW = os.environ.get("GADM_STAGE_NARROW", "0") == "1"


def _as_blocks(grads: GradsLike):
    """Normalise the accepted inputs to a list of [B, numel] views (one per parameter block)."""
    if isinstance(grads, torch.Tensor):
        if grads.dim() != 2:
            raise ValueError(f"grads must be [batch, grad_dim], got shape {tuple(grads.shape)}")
        blocks = [grads]
    else:
        vals = list(grads.values()) if isinstance(grads, Mapping) else list(grads)
        if not vals:
            raise ValueError("empty gradient collection")
        blocks = [v.reshape(v.shape[0], -1) for v in vals]
    out = []
    for b in blocks:
        if b.dtype not in _DTYPES:
            b = b.float()
        if b.shape[1] > 0 and b.stride(1) != 1:
            b = b.contiguous()
        out.append(b)
    bsz = out[0].shape[0]
    if any(b.shape[0] != bsz for b in out):
        raise ValueError("all gradient blocks must share the batch dimension")
    if len({b.dtype for b in out}) > 1:  # one launch converts one source dtype
        out = [b.float() for b in out]
    return out


class _Stage:
    """One staging buffer: 16-bit tile-major gradients [d_pad / 64][m_cap][64] (+ the F16G inverse scales)."""

    def __init__(self, projector: "CudaProjector", rows: int):
        p = projector
        self.rows = rows
        self.data = torch.zeros(p.d_pad // TILE_K, rows, TILE_K, dtype=_STAGE_TORCH[p.stage_dtype], device=p.device)
        self.inv_scale = (torch.ones(rows, p.scale_groups, dtype=torch.float32, device=p.device)
                          if p.stage_dtype == "f16" else None)
        self.done = None  # event of the last projection pass that read this buffer (overlap mode)


class CudaProjector:
    """``trak.projectors.CudaProjector`` signature over ``gadm_stage_rows`` + ``gadm_project_staged``."""

    def __init__(self, grad_dim: int, proj_dim: int, seed: int, proj_type: ProjectionType,
                 device, max_batch_size: int = 32, *args, stage_rows: int | None = None, cta_group: int | None = None,
                 stage_dtype: str | None = None, **kwargs) -> None:
        self.grad_dim = int(grad_dim)
        self.proj_dim = int(proj_dim)
        self.seed = int(seed)
        if isinstance(proj_type, str) and not isinstance(proj_type, ProjectionType):
            try:
                proj_type = ProjectionType(proj_type)
            except ValueError as e:
                raise KeyError(proj_type) from e
        if proj_type not in _PROJ_CODE:
            raise KeyError(proj_type)
        self.proj_type = proj_type
        self.max_batch_size = int(max_batch_size)
        if isinstance(device, str):
            device = torch.device(device)
        if device.type != "cuda":
            raise ValueError("CudaProjector only works on cuda device! Consider using BasicProjector instead"
                             " (this package has no CPU path).")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device
        if self.proj_dim <= 0 or self.proj_dim % 512 != 0:
            raise ValueError(f"proj_dim must be a positive multiple of 512 (got {self.proj_dim})")
        if self.grad_dim <= 0:
            raise ValueError("grad_dim must be positive")
        if stage_dtype is None:
            stage_dtype = os.environ.get("GADM_STAGE_DTYPE", "f16")
        if stage_dtype not in _STAGE_CODES:
            raise ValueError(f"stage_dtype must be 'f16' or 'bf16', got {stage_dtype!r}")
        self.stage_dtype = stage_dtype
        # kernel variant: 2 = one tcgen05 CTA pair per unit (512 rows per pass); 4 = two pairs per cluster sharing the
        # generated P tiles (1024 rows per pass) -- halves the Box-Muller work that bounds the normal type.
        if cta_group is None:
            cta_group = 4 if proj_type == ProjectionType.normal else 2
        self.cta_group = int(cta_group)
        self.d_pad = -(-self.grad_dim // TILE_K) * TILE_K
        self._handle = _lib.get_handle(device)
        self.scale_groups = int(self._handle.lib.gadm_stage_scale_count(self.d_pad))
        self.num_sms = torch.cuda.get_device_properties(device).multi_processor_count
        max_rows = 256 * self.cta_group
        if stage_rows is None:
            # largest pass that leaves a fifth of the free HBM to the caller (C3: D = 274 M -> 256 rows = 140 GB)
            free, _ = torch.cuda.mem_get_info(device)
            stage_rows = max_rows
            while stage_rows > 128 and stage_rows * self.d_pad * 2 > 0.8 * free:
                stage_rows //= 2
        self.stage_rows = int(min(max(int(stage_rows), 1), max_rows))
        self._stages: list[_Stage] = []
        self._ws = None
        self._slab = None
        self._side = None
        self._owner = None  # the DeferredProjection that has rows pending in the staging buffers

    # ------------------------------------------------------------------ buffers
    def _stage(self, rows: int, index: int = 0) -> _Stage:
        while len(self._stages) <= index:
            self._stages.append(None)
        st = self._stages[index]
        if st is None or st.rows < rows:
            self._stages[index] = None
            st = self._stages[index] = _Stage(self, rows)
        return st

    def _stage_buffer(self, rows: int) -> torch.Tensor:
        """The first staging buffer's data tensor [d_pad / 64][m_cap][64] (tests and bench fill it directly)."""
        return self._stage(rows).data

    def _group_for(self, rows: int) -> int:
        """cta_group 4 (two CTA pairs sharing the generated P tiles) only pays when both pairs have rows."""
        return 2 if (self.cta_group == 4 and rows <= 512) else self.cta_group

    def _workspace(self, rows: int) -> torch.Tensor:
        need = _lib.check(self._handle.lib.gadm_project_workspace_bytes(
            self._handle.ptr, rows, self.d_pad, self.proj_dim, self._group_for(rows)))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def _side_stream(self) -> torch.cuda.Stream:
        if self._side is None:
            # high priority: when a pass and the staging launches of the next one become runnable together, the
            # pass's clusters are placed first as SMs drain (staging CTAs are short) instead of queueing behind them
            self._side = torch.cuda.Stream(device=self.device, priority=-1)
        return self._side

    def _check_free(self, who) -> None:
        if self._owner is not None and self._owner is not who:
            raise RuntimeError("a DeferredProjection of this projector still has staged rows that were not projected; "
                               "call its flush() / result() first (the staging buffers are shared)")

    def free_memory(self) -> None:
        """trak AbstractProjector.free_memory."""
        self._check_free(None)
        self._stages = []
        self._ws = None
        self._slab = None

    # ------------------------------------------------------------------ kernels
    def _block_table(self, blocks):
        total = sum(b.shape[1] for b in blocks)
        if total != self.grad_dim:
            raise ValueError(f"gradient has {total} entries per example, projector was built for {self.grad_dim}")
        bsz = blocks[0].shape[0]
        live = [b for b in blocks if b.shape[1] > 0]
        arr = (_lib.Block * len(live))()
        col, i = 0, 0
        for b in blocks:
            if b.device != self.device:
                raise ValueError(f"gradients live on {b.device}, projector on {self.device}")
            numel = b.shape[1]
            if numel == 0:
                continue
            arr[i] = _lib.Block(b.data_ptr(), numel, b.stride(0) if bsz > 1 else numel, col)
            col += numel
            i += 1
        return arr, len(live), _DTYPES[blocks[0].dtype], bsz

    def _pack(self, blocks, stage: _Stage, row0: int, scale: float = 1.0, coresident: bool = False) -> int:
        """Stage a batch (all parameter blocks in one launch) into rows row0.. of `stage`.  ``coresident``: narrow CTAs
        that run beside a CTA-pair projection pass on the same SMs (gadm_stage_rows)."""
        arr, n, dtype, bsz = self._block_table(blocks)
        lib, h = self._handle.lib, self._handle.ptr
        _lib.check(lib.gadm_stage_rows(h, arr, n, dtype, bsz, float(scale), stage.data.data_ptr(),
                                       _STAGE_CODES[self.stage_dtype], self.d_pad, stage.rows, row0,
                                       stage.inv_scale.data_ptr() if stage.inv_scale is not None else None,
                                       int(coresident), _lib.stream_ptr(self.device)))
        return bsz

    def _accumulate(self, blocks, slab: torch.Tensor, scale: float, accumulate: bool) -> None:
        arr, n, dtype, bsz = self._block_table(blocks)
        _lib.check(self._handle.lib.gadm_accumulate_rows(self._handle.ptr, arr, n, dtype, bsz, float(scale),
                                                        slab.data_ptr(), self.d_pad, slab.shape[0], 0, int(accumulate),
                                                        _lib.stream_ptr(self.device)))

    def _project_rows(self, stage, rows: int, model_id: int, out: torch.Tensor) -> None:
        """One projection pass over the first `rows` staged rows, on the current stream."""
        if isinstance(stage, torch.Tensor):  # the data tensor of buffer 0 (tests / bench fill it directly)
            stage = next(s for s in self._stages if s is not None and s.data.data_ptr() == stage.data_ptr())
        ws = self._workspace(rows)
        seed64 = (self.seed + SEED_MODEL_ID_STRIDE * int(model_id)) & 0xFFFFFFFFFFFFFFFF
        _lib.check(self._handle.lib.gadm_project_staged(
            self._handle.ptr, stage.data.data_ptr(), _STAGE_CODES[self.stage_dtype],
            stage.inv_scale.data_ptr() if stage.inv_scale is not None else None, rows, self.d_pad, stage.rows, 0,
            self.proj_dim, seed64, _PROJ_CODE[self.proj_type], out.data_ptr(), out.stride(0), 0, ws.data_ptr(), ws.numel(),
            self._group_for(rows), _lib.stream_ptr(self.device)))

    # ------------------------------------------------------------------ reference API
    def project(self, grads: GradsLike, model_id: int) -> torch.Tensor:
        """[B, grad_dim] (or per-parameter blocks) -> [B, proj_dim]; returns immediately-usable values."""
        self._check_free(None)
        blocks = _as_blocks(grads)
        bsz = blocks[0].shape[0]
        in_dtype = blocks[0].dtype
        out = torch.empty(bsz, self.proj_dim, dtype=torch.float32, device=self.device)
        cap = self.stage_rows
        with torch.cuda.device(self.device):
            for r0 in range(0, bsz, cap):
                r1 = min(bsz, r0 + cap)
                stage = self._stage(min(cap, max(r1 - r0, min(self.max_batch_size, cap))))
                self._pack([b[r0:r1] for b in blocks], stage, 0)
                self._project_rows(stage, r1 - r0, model_id, out[r0:r1])
        return out if in_dtype == torch.float32 else out.to(in_dtype)

    def deferred(self, model_id: int = 0, overlap: bool | None = None, record_events: bool = False) -> "DeferredProjection":
        """Stage examples in HBM and project ``stage_rows`` of them per kernel pass.  ``overlap``: run the passes on a
        side stream against a second staging buffer so that staging the next batch (and whatever produces it) is
        not serialised behind the projection; ``None`` = when a second buffer fits in the free HBM.
        ``record_events``: keep a (start, end) CUDA event pair per pass in ``.pass_events`` (bench.py's roofline)."""
        return DeferredProjection(self, model_id, overlap, record_events)

    def materialize(self, row0: int, nrows: int, model_id: int = 0) -> torch.Tensor:
        """P[row0:row0+nrows, :] as fp32 -- the kernel's own matrix (test / oracle hook)."""
        out = torch.empty(nrows, self.proj_dim, dtype=torch.float32, device=self.device)
        seed64 = (self.seed + SEED_MODEL_ID_STRIDE * int(model_id)) & 0xFFFFFFFFFFFFFFFF
        with torch.cuda.device(self.device):
            _lib.check(self._handle.lib.gadm_materialize_p(self._handle.ptr, row0, nrows, self.proj_dim, seed64,
                                                          _PROJ_CODE[self.proj_type], _STAGE_CODES[self.stage_dtype],
                                                          out.data_ptr(), _lib.stream_ptr(self.device)))
        return out


class DeferredProjection:
    """Context manager returned by ``CudaProjector.deferred``.

    ``add(grads, scale)`` appends a batch (tensor or per-parameter blocks; ``scale`` folds e.g. the 1/K timestep mean
    of ``d_trak_grad.py:770``).  ``accumulate(grads, scale, last)`` is the featurisation loop's timestep sum
    (``d_trak_grad.py:757-770``, ``grad_text_to_image_lora.py:804-812``): every call adds ``scale * grads`` of the
    *current* batch into an fp32 slab, and ``last=True`` stages the summed batch -- no ``[B, D]`` torch temporaries,
    no ``vectorize_and_ignore_buffers``.  ``result()`` returns the [N, proj_dim] fp32 features in insertion order,
    device resident (replaces the per-batch ``.cpu()`` of ``d_trak_grad.py:792``)."""

    def __init__(self, projector: CudaProjector, model_id: int, overlap: bool | None = None, record_events: bool = False):
        self.p = projector
        self.pass_events = [] if record_events else None
        self.model_id = model_id
        self.rows = 0
        self.cur = 0
        self.outputs: list[torch.Tensor] = []
        self._acc_rows = 0  # batch size of the timestep sum in progress (0: none)
        if overlap is None:
            need = projector.stage_rows * projector.d_pad * 2
            have_second = len(projector._stages) > 1 and projector._stages[1] is not None
            free, _ = torch.cuda.mem_get_info(projector.device)
            first = 0 if projector._stages and projector._stages[0] is not None else need
            overlap = have_second or free > first + need + (16 << 30)
        self.overlap = bool(overlap)

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc, tb):
        if exc_type is None:
            self.flush()
        else:  # give the buffers back without projecting half-staged work
            self.rows = self._acc_rows = 0
            if self.p._owner is self:
                self.p._owner = None
        return False

    def _claim(self):
        self.p._check_free(self)
        self.p._owner = self

    def add(self, grads: GradsLike, scale: float = 1.0) -> None:
        if self._acc_rows:
            raise RuntimeError("add() while a timestep sum is in progress; finish it with accumulate(..., last=True)")
        self._claim()
        self._stage_blocks(_as_blocks(grads), scale)

    def _stage_blocks(self, blocks, scale: float) -> None:
        bsz = blocks[0].shape[0]
        cap = self.p.stage_rows
        done = 0
        with torch.cuda.device(self.p.device):
            while done < bsz:
                take = min(bsz - done, cap - self.rows)
                stage = self.p._stage(cap, self.cur)
                # wide staging CTAs: beside a quad-kernel pass they run on the 20 SMs its 32-cluster grid leaves free;
                # a pair-kernel pass covers every SM, so they queue behind it and then run at full speed (GADM_STAGE_NARROW=1
                # selects the narrow shape that runs beside the pair kernel instead -- measured slower)
                self.p._pack([b[done:done + take] for b in blocks], stage, self.rows, scale,
                             coresident=_STAGE_NARROW and self.overlap and self.p._group_for(cap) != 4)
                self.rows += take
                done += take
                if self.rows == cap:
                    self.flush()

    def accumulate(self, grads: GradsLike, scale: float = 1.0, last: bool = False) -> None:
        blocks = _as_blocks(grads)
        bsz = blocks[0].shape[0]
        p = self.p
        self._claim()
        if self._acc_rows and self._acc_rows != bsz:
            raise ValueError(f"timestep sum in progress holds {self._acc_rows} examples, got a batch of {bsz}")
        if p._slab is None or p._slab.shape[0] < bsz:
            p._slab = None
            p._slab = torch.empty(bsz, p.d_pad, dtype=torch.float32, device=p.device)
        with torch.cuda.device(p.device):
            p._accumulate(blocks, p._slab, scale, accumulate=self._acc_rows > 0)
        self._acc_rows = bsz
        if last:
            self._acc_rows = 0
            slab = p._slab[:bsz]
            self._stage_blocks([slab[:, :p.grad_dim]], 1.0)

    def flush(self) -> None:
        if self._acc_rows:
            raise RuntimeError("flush() while a timestep sum is in progress; finish it with accumulate(..., last=True)")
        if self.rows == 0:
            if self.p._owner is self:
                self.p._owner = None
            return
        p = self.p
        stage = p._stage(p.stage_rows, self.cur)
        out = torch.empty(self.rows, p.proj_dim, dtype=torch.float32, device=p.device)
        with torch.cuda.device(p.device):
            if not self.overlap:
                self._timed_pass(stage, out)
            else:
                main = torch.cuda.current_stream(p.device)
                side = p._side_stream()
                staged = torch.cuda.Event()
                staged.record(main)
                side.wait_event(staged)
                with torch.cuda.stream(side):
                    p._workspace(self.rows)  # allocated (once) under the side stream
                    self._timed_pass(stage, out)
                    stage.done = torch.cuda.Event()
                    stage.done.record(side)
                out.record_stream(side)
                self.cur ^= 1
                nxt = p._stage(p.stage_rows, self.cur)
                if nxt.done is not None:  # the pass that last read the other buffer must be over before it is refilled
                    main.wait_event(nxt.done)
        self.outputs.append(out)
        self.rows = 0
        p._owner = None

    def _timed_pass(self, stage, out) -> None:
        if self.pass_events is None:
            self.p._project_rows(stage, self.rows, self.model_id, out)
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()  # on the stream the kernel is launched on
        self.p._project_rows(stage, self.rows, self.model_id, out)
        e1.record()
        self.pass_events.append((e0, e1, self.rows))

    def result(self) -> torch.Tensor:
        self.flush()
        p = self.p
        if self.overlap:
            main = torch.cuda.current_stream(p.device)
            for st in p._stages:
                if st is not None and st.done is not None:
                    main.wait_event(st.done)
        if not self.outputs:
            return torch.empty(0, p.proj_dim, dtype=torch.float32, device=p.device)
        return self.outputs[0] if len(self.outputs) == 1 else torch.cat(self.outputs, dim=0)


class BasicProjector(CudaProjector):
    """``trak.projectors.BasicProjector`` signature.  Upstream's version builds P block by block with
    torch ops on any device; here it is the same sm_100a kernel (CUDA only, no CPU path)."""

    def __init__(self, grad_dim: int, proj_dim: int, seed: int, proj_type: ProjectionType, device,
                 block_size: int = 100, dtype: torch.dtype = torch.float32, model_id: int = 0, *args, **kwargs) -> None:
        super().__init__(grad_dim, proj_dim, seed, proj_type, device, max_batch_size=32, **kwargs)
        self.block_size = block_size
        self.dtype = dtype
        self.model_id = model_id

    def project(self, grads: GradsLike, model_id: int | None = None) -> torch.Tensor:
        return super().project(grads, self.model_id if model_id is None else model_id)


def is_not_buffer(ind, params_dict) -> bool:
    """trak.utils.is_not_buffer (imported at d_trak_grad.py:15): False for BatchNorm buffers."""
    name = params_dict[ind]
    if ("running_mean" in name) or ("running_var" in name) or ("num_batches_tracked" in name):
        return False
    return True
