// Philox4x32-10 and the (p, j) -> P[p, j] element map of the JL projection, device side.
// Mirrors oracle/philox.py bit for bit (Rademacher) / to MUFU accuracy (normal).
// Replaces the on-the-fly matrix generation inside fast_jl's project_{normal,rademacher}_* kernels,
// reached from trak.projectors.CudaProjector.project (reference call sites
// src/attributions/methods/d_trak_grad.py:776, text_to_image/grad_text_to_image_lora.py:765,813).
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace gadm {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;
constexpr uint32_t kTagRademacher = 0x52414445u;  // "RADE"
constexpr uint32_t kTagNormal = 0x4E4F524Du;      // "NORM"

enum ProjType : int { kProjNormal = 0, kProjRademacher = 1 };

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = static_cast<uint64_t>(kPhiloxM0) * c0;
    const uint64_t p1 = static_cast<uint64_t>(kPhiloxM1) * c2;
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
    c1 = static_cast<uint32_t>(p1);
    c3 = static_cast<uint32_t>(p0);
    c0 = n0;
    c2 = n2;
    k0 += kPhiloxW0;
    k1 += kPhiloxW1;
  }
  return make_uint4(c0, c1, c2, c3);
}

// ---- Rademacher: one call covers 32 consecutive p (bits) x 4 consecutive j (words).
__device__ __forceinline__ uint4 rademacher_call(uint32_t p_div32, uint32_t j_div4, uint32_t k0, uint32_t k1) {
  return philox4x32_10(p_div32, j_div4, 0u, kTagRademacher, k0, k1);
}
// P is written in the 16-bit format of the staged gradients (tcgen05 kind::f16 takes f16 x f16 or bf16 x bf16; a
// mixed pair faults on sm_100a): +-1 is exact in both, the normal type is Box-Muller rounded to that format.
constexpr uint32_t kOnesBf16 = 0x3F803F80u;  // two bf16 1.0
constexpr uint32_t kOnesF16 = 0x3C003C00u;   // two fp16 1.0

// 8 sign bits -> 8 16-bit (+1 / -1) values packed in a 16-byte chunk; bit e -> element e (bit set = -1).
// `ones` = kOnesBf16 or kOnesF16.
__device__ __forceinline__ uint4 rademacher_expand8(uint32_t byte, uint32_t ones) {
  // (byte >> 2q) * (2^15 + 2^30): bit 2q -> bit 15, bit 2q+1 -> bit 31 (byte < 2^8 so the copies never overlap)
  uint4 o;
  o.x = ((byte * 0x40008000u) & 0x80008000u) | ones;
  o.y = ((byte * 0x10002000u) & 0x80008000u) | ones;
  o.z = ((byte * 0x04000800u) & 0x80008000u) | ones;
  o.w = ((byte * 0x01000200u) & 0x80008000u) | ones;
  return o;
}

// ---- Normal: one call covers 8 consecutive p for one j -> one 16-byte chunk of bf16 (kF16 = false) or fp16.
template <bool kF16>
__device__ __forceinline__ uint32_t box_muller_pair(uint32_t x) {
  // lo16 -> u1 = (lo + 0.5) / 2^16 (exact), hi16 -> turn = (hi + 0.5) / 2^16 (exact)
  const float flo = __uint_as_float(0x4B000000u | (x & 0xFFFFu));         // 2^23 + lo
  const float fhi = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7632));  // 2^23 + hi
  const float kBias = -128.0f + 0.0000076293945312f;                       // -2^23 * 2^-16 + 2^-17
  const float u1 = __fmaf_rn(flo, 0.0000152587890625f, kBias);
  const float turn = __fmaf_rn(fhi, 0.0000152587890625f, kBias);
  const float theta = __fmul_rn(turn, 6.283185307179586f);
  // .ftz forms: one MUFU each (the non-ftz lg2/sin/cos add FSETP/FMUL/FADD subnormal fix-ups; u1 >= 2^-17)
  float l2, r, s, c;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u1));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fmul_rn(-1.3862943611198906f, l2)));  // sqrt(-2 ln u1)
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(theta));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(theta));
  if constexpr (kF16) {
    const __half2 v = __floats2half2_rn(__fmul_rn(r, c), __fmul_rn(r, s));  // .x (low) = even p
    // P keeps 8 significant bits (like the bf16 path), stored as fp16: half-ulp bias on both halves, then the three
    // low mantissa bits are cleared (a mantissa carry moves into the exponent, as rounding up to the next binade
    // must; |z| < 5, so no carry leaves a half).  Measured on B200 under the 1 kW cap: with 11-bit P the fp16 x fp16
    // pass is 3.6 % slower than the bf16 x bf16 one (266 vs 257 ms, C2), with 8-bit P it is not (257 ms) -- the
    // toggling low mantissa bits of the B operand are what the tensor pipe pays for; the staged gradients keep all
    // 11 bits, which is what the accuracy comes from.
    return (*reinterpret_cast<const uint32_t*>(&v) + 0x00040004u) & 0xFFF8FFF8u;
  } else {
    const __nv_bfloat162 v = __floats2bfloat162_rn(__fmul_rn(r, c), __fmul_rn(r, s));
    return *reinterpret_cast<const uint32_t*>(&v);
  }
}
template <bool kF16>
__device__ __forceinline__ uint4 normal_chunk(uint32_t p_div8, uint32_t j, uint32_t k0, uint32_t k1) {
  const uint4 w = philox4x32_10(p_div8, j, 0u, kTagNormal, k0, k1);
  return make_uint4(box_muller_pair<kF16>(w.x), box_muller_pair<kF16>(w.y), box_muller_pair<kF16>(w.z),
                    box_muller_pair<kF16>(w.w));
}

}  // namespace gadm
