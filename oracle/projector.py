"""Projector oracles.  TEST ORACLE (see oracle/__init__.py).

Two restatements:

* ``BasicProjectorOracle`` -- trak ``BasicProjector`` (traker==0.1.3,
  ``trak/projectors.py``; requirements.txt:10).  The package is not under
  /root/reference and cannot be installed here, so this follows its published
  algorithm: ``block_size = min(proj_dim, 100)``; for block i a
  ``torch.Generator`` seeded with ``seed + 1000*i + 100000*model_id`` fills a
  ``[grad_dim, block]`` matrix with ``normal_`` (or ``bernoulli_(0.5)*2-1``) and
  ``sketch[:, st:ed] = grads.float() @ proj_matrix``; no normalisation.  This is the
  reference's CPU projection path and is what ``bench.py`` times as
  ``cpu_baseline``.  It shares no random stream with the CUDA kernel (neither does
  trak's own CudaProjector), so it is used for throughput and JL statistics only.

* ``project_explicit`` -- the same-matrix parity check: Phi = staged(G) @ P(seed) in
  float64, staged() being the kernel's 16-bit input format (``stage="f16"``: fp16 with a
  power-of-two scale per 32768-column group, the default; ``"bf16"``), with P from
  ``oracle.philox`` (bit-exact for Rademacher) or handed in by
  the caller (the kernel's own ``gadm_materialize_p`` output for the normal type,
  whose Box-Muller uses MUFU approximations and matches ``oracle.philox`` only to
  one bf16 ulp).  Call-site semantics: ``d_trak_grad.py:776``,
  ``grad_text_to_image_lora.py:765,813`` -- ``[B, D] -> [B, k]`` float32.
"""
from __future__ import annotations

import numpy as np

from . import philox


def project_explicit(grads: np.ndarray, P: np.ndarray | None = None, *, seed: int = 0, model_id: int = 0,
                     proj_type: str = "rademacher", proj_dim: int | None = None, row_offset: int = 0,
                     stage: str | None = "f16", chunk: int = 1 << 15) -> np.ndarray:
    """fp64 reference of the kernel's contraction: [B, D] x P[row_offset:row_offset+D, :k]."""
    g = philox.round_staged(np.asarray(grads, dtype=np.float32), stage)
    B, D = g.shape
    if P is not None:
        return g.astype(np.float64) @ np.asarray(P, dtype=np.float64)
    out = np.zeros((B, proj_dim), dtype=np.float64)
    for s in range(0, D, chunk):
        e = min(D, s + chunk)
        Pc = philox.projection_matrix(seed, model_id, proj_type, row_offset + s, e - s, proj_dim, stage or "f16")
        out += g[:, s:e].astype(np.float64) @ Pc.astype(np.float64)
    return out


class BasicProjectorOracle:
    """trak BasicProjector semantics on CPU torch (see module docstring)."""

    def __init__(self, grad_dim: int, proj_dim: int, seed: int, proj_type: str = "normal",
                 block_size: int = 100, model_id: int = 0):
        import torch

        self.torch = torch
        self.grad_dim = grad_dim
        self.proj_dim = proj_dim
        self.seed = seed
        self.proj_type = proj_type
        self.model_id = model_id
        self.block_size = min(proj_dim, block_size)
        self.num_blocks = -(-proj_dim // self.block_size)
        self.generator = torch.Generator(device="cpu")
        self.proj_matrix = torch.empty(grad_dim, self.block_size, dtype=torch.float32)

    def _fill(self, block: int, model_id: int):
        self.generator.manual_seed(self.seed + int(1e3) * block + int(1e5) * model_id)
        if self.proj_type == "normal":
            self.proj_matrix.normal_(generator=self.generator)
        elif self.proj_type == "rademacher":
            self.proj_matrix.bernoulli_(p=0.5, generator=self.generator)
            self.proj_matrix *= 2.0
            self.proj_matrix -= 1.0
        else:
            raise KeyError(self.proj_type)

    def project(self, grads, model_id: int | None = None, blocks: int | None = None):
        torch = self.torch
        model_id = self.model_id if model_id is None else model_id
        grads = torch.as_tensor(grads)
        sketch = torch.zeros(grads.shape[0], self.proj_dim, dtype=torch.float32)
        nb = self.num_blocks if blocks is None else min(blocks, self.num_blocks)
        for i in range(nb):
            self._fill(i, model_id)
            st = i * self.block_size
            ed = min((i + 1) * self.block_size, self.proj_dim)
            sketch[:, st:ed] = (grads.float() @ self.proj_matrix)[:, : ed - st]
        return sketch.to(grads.dtype)
