"""Philox4x32-10 and the counter -> matrix-element map of the JL projection.  TEST ORACLE.

The reference projects gradients with ``trak.projectors.CudaProjector`` (call sites
``src/attributions/methods/d_trak_grad.py:504-511,776`` and
``text_to_image/grad_text_to_image_lora.py:561-568,765,813``), which never
materialises the projection matrix P[D, k] but generates it on the fly from
``seed + 10**4 * model_id`` inside the third-party ``fast_jl`` CUDA kernel
(``requirements.txt:10-11``; not present under /root/reference).  No two trak
projectors agree on P for a given seed, so the matrix *values* are defined by this
framework; what is kept from the reference is the semantics: i.i.d. N(0,1) or
+-1 entries, no 1/sqrt(k) scaling, a pure function of (seed, model_id).

Definition used by the CUDA kernel and restated here (Salmon et al., SC'11,
Philox4x32 with 10 rounds; multipliers 0xD2511F53 / 0xCD9E8D57, Weyl key
increments 0x9E3779B9 / 0xBB67AE85):

  key            = (seed64 & 0xffffffff, seed64 >> 32),  seed64 = seed + 10**4 * model_id
  Rademacher     : ctr = (p >> 5, j >> 2, 0, TAG_RADEMACHER); word = j & 3; bit = p & 31
                   P[p, j] = -1 if bit set else +1
  Normal         : ctr = (p >> 3, j, 0, TAG_NORMAL); word = (p >> 1) & 3
                   lo = word & 0xffff, hi = word >> 16
                   u1 = (lo + 0.5) / 2**16, theta = 2*pi*(hi + 0.5) / 2**16
                   r = sqrt(-2 ln u1);  P[p, j] = r*cos(theta) if p even else r*sin(theta)
                   then rounded (nearest-even) to the 16-bit format of the staged gradients: fp16 for the
                   default "f16" staging, bfloat16 for "bf16" staging (tcgen05 kind::f16 multiplies
                   f16 x f16 or bf16 x bf16, not a mixed pair).

p is the canonical index of a parameter in the flattened gradient (row of P), j the
output feature (column of P).  The map does not depend on tiling, SM count, split-K
or the number of GPUs.
"""
from __future__ import annotations

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

TAG_RADEMACHER = 0x52414445  # "RADE"
TAG_NORMAL = 0x4E4F524D  # "NORM"

SEED_MODEL_ID_STRIDE = 10**4  # trak CudaProjector: seed + int(1e4) * model_id


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Vectorised Philox4x32-10.  c0..c3: broadcastable uint32/uint64 arrays.

    Returns four uint32 arrays (the 128-bit output block)."""
    c0 = np.asarray(c0, dtype=np.uint64) & MASK32
    c1 = np.asarray(c1, dtype=np.uint64) & MASK32
    c2 = np.asarray(c2, dtype=np.uint64) & MASK32
    c3 = np.asarray(c3, dtype=np.uint64) & MASK32
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0)
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def seed64_of(seed: int, model_id: int = 0) -> int:
    return (int(seed) + SEED_MODEL_ID_STRIDE * int(model_id)) & 0xFFFFFFFFFFFFFFFF


def _key(seed64: int):
    return seed64 & 0xFFFFFFFF, (seed64 >> 32) & 0xFFFFFFFF


def round_to_bf16(x: np.ndarray) -> np.ndarray:
    """float32 -> nearest-even bfloat16, returned as float32."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    rounding = np.uint64(0x7FFF) + ((u >> np.uint64(16)) & np.uint64(1))
    u = (u + rounding) & np.uint64(0xFFFF0000)
    return u.astype(np.uint32).view(np.float32).reshape(x.shape)


STAGE_GROUP_COLS = 32768  # GADM_STAGE_GROUP_COLS (include/gadm.h)


def _floor_log2(a: np.ndarray) -> np.ndarray:
    return ((a.astype(np.float32).view(np.uint32) >> np.uint32(23)) & np.uint32(0xFF)).astype(np.int64) - 127


def f16_group_scale_exponents(blk: np.ndarray, scale: float = 1.0) -> np.ndarray:
    """Scale exponent s per row of one 32768-column group (csrc/stage.cuh ``stage_groups_kernel``): guessed from the
    sample columns [1024 j, 1024 j + 1024), j % 8 == 0 (sampled max -> [2^11, 2^12)); kept if the true maximum then
    lands in [2^8, 65504), otherwise the exact scale (true max -> [2^13, 2^14)); 0 for an all-zero / non-finite group."""
    sc = np.float32(abs(scale))
    cols = np.arange(blk.shape[1])
    sample = blk[:, (cols // 1024) % 8 == 0]

    def amax_of(x):
        with np.errstate(invalid="ignore"):
            m = np.fmax.reduce(np.abs(x), axis=1, initial=0.0).astype(np.float32)  # fmax skips NaNs like the kernel
        return (m * sc).astype(np.float32)

    def expo(a, top):
        ok = (a > 0) & np.isfinite(a)
        e = np.clip(np.where(ok, _floor_log2(np.where(ok, a, np.float32(1))), 0), -100, 100)
        return np.where(ok, top - e, 0)

    s_guess = expo(amax_of(sample), 11)
    true_max = amax_of(blk)
    with np.errstate(over="ignore", invalid="ignore"):
        landed = (true_max * np.ldexp(np.float32(1), s_guess).astype(np.float32)).astype(np.float32)
    keep = ((landed >= 256) & (landed < 65504)) | (true_max == 0)
    return np.where(keep, s_guess, expo(true_max, 13))


def round_to_f16_groups(x: np.ndarray, scale: float = 1.0) -> np.ndarray:
    """The GADM_STAGE_F16G staging format (csrc/stage.cuh), returned as the float32 values the kernel multiplies by P:
    per (row, group of 32768 columns) a power-of-two scale 2^s (``f16_group_scale_exponents``), staged value
    float16(float32(x * (scale * 2^s))) (round-to-nearest-even twice), standing for staged * 2^-s."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    sc = np.float32(scale)
    for lo in range(0, x.shape[1], STAGE_GROUP_COLS):
        blk = x[:, lo:lo + STAGE_GROUP_COLS]
        s = f16_group_scale_exponents(blk, scale)
        mul = (sc * np.ldexp(np.float32(1.0), s).astype(np.float32)).astype(np.float32)
        with np.errstate(over="ignore", invalid="ignore"):
            staged = (blk * mul[:, None]).astype(np.float32).astype(np.float16)
        out[:, lo:lo + STAGE_GROUP_COLS] = np.ldexp(staged.astype(np.float32), -s[:, None]).astype(np.float32)
    return out


def round_staged(x: np.ndarray, stage: str | None = "f16", scale: float = 1.0) -> np.ndarray:
    """What the projection kernel multiplies by P for gradients x under a staging format ('f16', 'bf16', None)."""
    if stage == "f16":
        return round_to_f16_groups(x, scale)
    xs = (np.asarray(x, dtype=np.float32) * np.float32(scale)).astype(np.float32)
    return round_to_bf16(xs) if stage == "bf16" else xs


def rademacher_matrix(seed64: int, row0: int, nrows: int, k: int) -> np.ndarray:
    """P[row0:row0+nrows, 0:k] as int8 (+1 / -1)."""
    k0, k1 = _key(seed64)
    p = np.arange(row0, row0 + nrows, dtype=np.uint64)
    j = np.arange(k, dtype=np.uint64)
    pg = np.unique(p >> np.uint64(5))
    jg = np.unique(j >> np.uint64(2))
    words = philox4x32_10(pg[:, None], jg[None, :], 0, TAG_RADEMACHER, k0, k1)
    words = np.stack(words, axis=-1)  # [npg, njg, 4]
    pi = np.searchsorted(pg, p >> np.uint64(5))
    ji = np.searchsorted(jg, j >> np.uint64(2))
    w = words[pi[:, None], ji[None, :], (j & np.uint64(3)).astype(np.int64)[None, :]]
    bit = (w >> (p & np.uint64(31)).astype(np.uint32)[:, None]) & np.uint32(1)
    return (1 - 2 * bit.astype(np.int8)).astype(np.int8)


def normal_matrix(seed64: int, row0: int, nrows: int, k: int, fmt: str | None = "f16") -> np.ndarray:
    """P[row0:row0+nrows, 0:k] as float32, Box-Muller in float64 rounded to the kernel's operand format: "f16"
    (the default staging format: fp16 numbers that keep 8 significant bits), "bf16", or None for the unrounded
    float32 values."""
    k0, k1 = _key(seed64)
    p = np.arange(row0, row0 + nrows, dtype=np.uint64)
    j = np.arange(k, dtype=np.uint64)
    pg = np.unique(p >> np.uint64(3))
    words = philox4x32_10(pg[:, None], j[None, :], 0, TAG_NORMAL, k0, k1)
    words = np.stack(words, axis=-1)  # [npg, k, 4]
    pi = np.searchsorted(pg, p >> np.uint64(3))
    wsel = ((p >> np.uint64(1)) & np.uint64(3)).astype(np.int64)
    w = words[pi[:, None], np.arange(k)[None, :], wsel[:, None]].astype(np.uint64)
    lo = (w & np.uint64(0xFFFF)).astype(np.float64)
    hi = (w >> np.uint64(16)).astype(np.float64)
    u1 = (lo + 0.5) / 65536.0
    theta = 2.0 * np.pi * (hi + 0.5) / 65536.0
    r = np.sqrt(-2.0 * np.log(u1))
    odd = (p & np.uint64(1)).astype(bool)[:, None]
    z = np.where(odd, r * np.sin(theta), r * np.cos(theta)).astype(np.float32)
    if fmt == "f16":  # fp16 with an 8-bit significand (philox.cuh::box_muller_pair): + half ulp of the dropped bits, mask
        bits = z.astype(np.float16).view(np.uint16).astype(np.uint32)
        return ((bits + 4) & 0xFFF8).astype(np.uint16).view(np.float16).astype(np.float32)
    return round_to_bf16(z) if fmt == "bf16" else z


def projection_matrix(seed: int, model_id: int, proj_type: str, row0: int, nrows: int, k: int,
                      fmt: str | None = "f16") -> np.ndarray:
    s = seed64_of(seed, model_id)
    if proj_type == "rademacher":
        return rademacher_matrix(s, row0, nrows, k).astype(np.float32)
    if proj_type == "normal":
        return normal_matrix(s, row0, nrows, k, fmt)
    raise KeyError(proj_type)
