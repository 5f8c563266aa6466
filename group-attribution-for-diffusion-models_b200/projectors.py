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
* ``deferred()`` stages up to 512 (Rademacher) / 1024 (normal) examples (bf16 rows in HBM) and projects them
  in one pass, which is what makes the tensor cores, not the random-number generation, the bound.

There is no CPU path: a non-CUDA device raises ``ValueError`` exactly like upstream's CudaProjector.
"""
from __future__ import annotations

from enum import Enum
from typing import Iterable, Mapping, Sequence, Union

import torch

from . import _lib

SEED_MODEL_ID_STRIDE = int(1e4)  # trak.projectors.CudaProjector.project: seed + int(1e4) * model_id
MAX_STAGE_ROWS = 512
TILE_K = 64

_DTYPES = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}

GradsLike = Union[torch.Tensor, Mapping[str, torch.Tensor], Sequence[torch.Tensor]]


class ProjectionType(str, Enum):
    """trak.projectors.ProjectionType."""

    normal = "normal"
    rademacher = "rademacher"


_PROJ_CODE = {ProjectionType.normal: 0, ProjectionType.rademacher: 1}


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
    return out


class CudaProjector:
    """``trak.projectors.CudaProjector`` signature over ``gadm_project_staged``."""

    def __init__(self, grad_dim: int, proj_dim: int, seed: int, proj_type: ProjectionType,
                 device, max_batch_size: int = 32, *args, stage_rows: int | None = None, cta_group: int | None = None,
                 **kwargs) -> None:
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
        # kernel variant: 2 = one tcgen05 CTA pair per unit (512 rows per pass); 4 = two pairs per cluster sharing the
        # generated P tiles (1024 rows per pass) -- halves the Box-Muller work that bounds the normal type.
        if cta_group is None:
            cta_group = 4 if proj_type == ProjectionType.normal else 2
        self.cta_group = int(cta_group)
        self.d_pad = -(-self.grad_dim // TILE_K) * TILE_K
        self._handle = _lib.get_handle(device)
        self.num_sms = torch.cuda.get_device_properties(device).multi_processor_count
        max_rows = 256 * self.cta_group
        if stage_rows is None:
            total = torch.cuda.get_device_properties(device).total_memory
            stage_rows = max_rows
            while stage_rows > 128 and stage_rows * self.d_pad * 2 > 0.45 * total:
                stage_rows //= 2
        self.stage_rows = int(min(max(int(stage_rows), 1), max_rows))
        self._stage = None
        self._ws = None

    # ------------------------------------------------------------------ buffers
    def _stage_buffer(self, rows: int) -> torch.Tensor:
        """bf16 staging buffer in the kernel's tile-major layout [d_pad / 64][m_cap][64] (include/gadm.h)."""
        if self._stage is None or self._stage.shape[1] < rows:
            self._stage = None
            self._stage = torch.zeros(self.d_pad // TILE_K, rows, TILE_K, dtype=torch.bfloat16, device=self.device)
        return self._stage

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

    def free_memory(self) -> None:
        """trak AbstractProjector.free_memory."""
        self._stage = None
        self._ws = None

    # ------------------------------------------------------------------ kernels
    def _pack(self, blocks, stage: torch.Tensor, row0: int, scale: float = 1.0) -> int:
        bsz = blocks[0].shape[0]
        total = sum(b.shape[1] for b in blocks)
        if total != self.grad_dim:
            raise ValueError(f"gradient has {total} entries per example, projector was built for {self.grad_dim}")
        lib, h, st = self._handle.lib, self._handle.ptr, _lib.stream_ptr(self.device)
        col = 0
        for b in blocks:
            if b.device != self.device:
                raise ValueError(f"gradients live on {b.device}, projector on {self.device}")
            numel = b.shape[1]
            if numel == 0:
                continue
            _lib.check(lib.gadm_pack_block(h, b.data_ptr(), _DTYPES[b.dtype], bsz, numel, b.stride(0) if bsz > 1 else numel,
                                          stage.data_ptr(), self.d_pad, stage.shape[1], row0, col, float(scale), st))
            col += numel
        return bsz

    def _project_rows(self, stage: torch.Tensor, rows: int, model_id: int, out: torch.Tensor) -> None:
        ws = self._workspace(rows)
        seed64 = (self.seed + SEED_MODEL_ID_STRIDE * int(model_id)) & 0xFFFFFFFFFFFFFFFF
        _lib.check(self._handle.lib.gadm_project_staged(
            self._handle.ptr, stage.data_ptr(), rows, self.d_pad, stage.shape[1], 0, self.proj_dim, seed64,
            _PROJ_CODE[self.proj_type], out.data_ptr(), out.stride(0), 0, ws.data_ptr(), ws.numel(),
            self._group_for(rows), _lib.stream_ptr(self.device)))

    # ------------------------------------------------------------------ reference API
    def project(self, grads: GradsLike, model_id: int) -> torch.Tensor:
        """[B, grad_dim] (or per-parameter blocks) -> [B, proj_dim]; returns immediately-usable values."""
        blocks = _as_blocks(grads)
        bsz = blocks[0].shape[0]
        in_dtype = blocks[0].dtype
        out = torch.empty(bsz, self.proj_dim, dtype=torch.float32, device=self.device)
        cap = self.stage_rows
        with torch.cuda.device(self.device):
            for r0 in range(0, bsz, cap):
                r1 = min(bsz, r0 + cap)
                stage = self._stage_buffer(min(cap, max(r1 - r0, min(self.max_batch_size, cap))))
                self._pack([b[r0:r1] for b in blocks], stage, 0)
                self._project_rows(stage, r1 - r0, model_id, out[r0:r1])
        return out if in_dtype == torch.float32 else out.to(in_dtype)

    def deferred(self, model_id: int = 0) -> "DeferredProjection":
        """Stage examples in HBM and project ``stage_rows`` of them per kernel pass."""
        return DeferredProjection(self, model_id)

    def materialize(self, row0: int, nrows: int, model_id: int = 0) -> torch.Tensor:
        """P[row0:row0+nrows, :] as fp32 -- the kernel's own matrix (test / oracle hook)."""
        out = torch.empty(nrows, self.proj_dim, dtype=torch.float32, device=self.device)
        seed64 = (self.seed + SEED_MODEL_ID_STRIDE * int(model_id)) & 0xFFFFFFFFFFFFFFFF
        with torch.cuda.device(self.device):
            _lib.check(self._handle.lib.gadm_materialize_p(self._handle.ptr, row0, nrows, self.proj_dim, seed64,
                                                          _PROJ_CODE[self.proj_type], out.data_ptr(),
                                                          _lib.stream_ptr(self.device)))
        return out


class DeferredProjection:
    """Context manager returned by ``CudaProjector.deferred``.

    ``add(grads, scale)`` appends a batch (tensor or per-parameter blocks; ``scale`` folds e.g. the
    1/K timestep mean of ``d_trak_grad.py:770``); ``accumulate(grads, scale)`` adds into the rows of
    the *last added* batch is not supported in bf16 staging -- sum timesteps in fp32 first.
    ``result()`` returns the [N, proj_dim] fp32 features in insertion order, device resident.
    """

    def __init__(self, projector: CudaProjector, model_id: int):
        self.p = projector
        self.model_id = model_id
        self.rows = 0
        self.outputs: list[torch.Tensor] = []

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc, tb):
        if exc_type is None:
            self.flush()
        return False

    def add(self, grads: GradsLike, scale: float = 1.0) -> None:
        blocks = _as_blocks(grads)
        bsz = blocks[0].shape[0]
        cap = self.p.stage_rows
        done = 0
        with torch.cuda.device(self.p.device):
            while done < bsz:
                take = min(bsz - done, cap - self.rows)
                stage = self.p._stage_buffer(cap)
                self.p._pack([b[done:done + take] for b in blocks], stage, self.rows, scale)
                self.rows += take
                done += take
                if self.rows == cap:
                    self.flush()

    def flush(self) -> None:
        if self.rows == 0:
            return
        out = torch.empty(self.rows, self.p.proj_dim, dtype=torch.float32, device=self.p.device)
        with torch.cuda.device(self.p.device):
            self.p._project_rows(self.p._stage_buffer(self.p.stage_rows), self.rows, self.model_id, out)
        self.outputs.append(out)
        self.rows = 0

    def result(self) -> torch.Tensor:
        self.flush()
        if not self.outputs:
            return torch.empty(0, self.p.proj_dim, dtype=torch.float32, device=self.p.device)
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
