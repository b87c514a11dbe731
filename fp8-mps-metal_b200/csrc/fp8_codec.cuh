// Device-side e4m3fn codec with the REFERENCE's semantics (fp8_matmul.metal:19-92), built on the
// sm_100a conversion instructions instead of the shader's exp2/log2 arithmetic.
//
// decode (metal:19-40):  cvt.rn.f16x2.e4m3x2 is exact for every non-NaN byte (incl. subnormals and
//   0x80 -> -0.0); the two NaN bytes 0x7F/0xFF come out as fp16 NaN and are forced to +0.0 (:21).
//
// encode (metal:44-92) differs from the hardware cvt.rn.satfinite.e4m3x2 in exactly three ways,
//   each repaired on the INPUT so the hardware conversion then produces the shader's byte:
//   (1) -0.0 -> 0x00 (sign = x < 0, :46): add +0.0 first;
//   (2) |v| < 2^-9 -> signed zero (:58-60; the hardware rounds (2^-10,2^-9) up to 0x01): replace
//       such inputs by +-0;
//   (3) the mantissa is rounded to nearest-even but clamped to 7 instead of carrying into the
//       exponent (:68,:81): clamp the fraction field to just below the carry point
//       (frac <= 0.875-ulp; every fraction in (0.8125, 0.9375) encodes m = 7, and in the
//       subnormal-target binade [2^-7,2^-6) the clamp lands below the 7.5 tie), done as an
//       unsigned min on the raw bits -- sign, exponent, inf and NaN are untouched.
//   Saturation (v >= 448 -> 0x7E, :53-55, inf included) is what .satfinite does; (15,7)->(15,6)
//   (:87-89) can never arise after it.  NaN -> 0x7F (undefined in the reference).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>

namespace fp8b {

// ---------------------------------------------------------------- decode

// 2 fp8 bytes (low 16 bits of `pair`) -> f16x2, hardware conversion, NaN bytes -> fp16 NaN.
__device__ __forceinline__ uint32_t cvt_e4m3x2_f16x2_raw(uint16_t pair) {
    uint32_t r;
    asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(r) : "h"(pair));
    return r;
}

// 4 fp8 bytes -> two f16x2 registers (elements 0,1 and 2,3), raw hardware semantics.
__device__ __forceinline__ void dec4_f16x2_raw(uint32_t w, uint32_t& lo, uint32_t& hi) {
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\t"
        "cvt.rn.f16x2.e4m3x2 %0, l;\n\tcvt.rn.f16x2.e4m3x2 %1, h;\n\t}"
        : "=r"(lo), "=r"(hi) : "r"(w));
}

// f16x2 NaN lanes -> +0.0 (fp8_matmul.metal:21).
__device__ __forceinline__ uint32_t nan_to_zero_f16x2(uint32_t v) {
    __half2 h = *reinterpret_cast<__half2*>(&v);
    return v & __heq2_mask(h, h);
}

// 4 fp8 bytes -> two f16x2 registers with the reference's NaN -> 0 rule.
__device__ __forceinline__ void dec4_f16x2(uint32_t w, uint32_t& lo, uint32_t& hi) {
    dec4_f16x2_raw(w, lo, hi);
    lo = nan_to_zero_f16x2(lo);
    hi = nan_to_zero_f16x2(hi);
}

// Scalar decode to fp32 (slow paths only).
__device__ __forceinline__ float dec1_f32(uint8_t b) {
    uint32_t v = nan_to_zero_f16x2(cvt_e4m3x2_f16x2_raw((uint16_t)b));
    return __half2float(__ushort_as_half((unsigned short)(v & 0xFFFF)));
}

// ---------------------------------------------------------------- float8_e5m2 (decode only)
// e5m2 is the upper byte of an IEEE binary16: decode = byte << 8, exact, +-inf and NaN preserved (the format's own
// definition, pinned to PyTorch's CPU cast in tests/golden/e5m2_golden.npz).  The reference accepts e5m2 tensors
// but decodes them as e4m3fn (fp8_mps_patch.py:48-49,65) -- there is no reference codec to follow.
// Formats are numbered as in include/fp8_b200.h: 0 = e4m3fn, 1 = e5m2.

// 4 e5m2 bytes -> two f16x2 registers: two byte permutes, no conversion instruction.
__device__ __forceinline__ void dec4_e5m2_f16x2(uint32_t w, uint32_t& lo, uint32_t& hi) {
    lo = __byte_perm(w, 0u, 0x1404);        // bytes {0, b0, 0, b1}
    hi = __byte_perm(w, 0u, 0x3424);        // bytes {0, b2, 0, b3}
}

__device__ __forceinline__ float dec1_e5m2_f32(uint8_t b) {
    return __half2float(__ushort_as_half((unsigned short)((unsigned)b << 8)));
}

// 4 bytes of format FMT -> two f16x2 registers, reference NaN rule for e4m3fn (NaN -> 0), IEEE for e5m2.
template <int FMT>
__device__ __forceinline__ void dec4_fmt_f16x2(uint32_t w, uint32_t& lo, uint32_t& hi) {
    if (FMT == 0) dec4_f16x2(w, lo, hi); else dec4_e5m2_f16x2(w, lo, hi);
}

__device__ __forceinline__ float dec1_fmt_f32(uint8_t b, int fmt) { return fmt ? dec1_e5m2_f32(b) : dec1_f32(b); }

// ---------------------------------------------------------------- encode

// fp32 input repair, steps (1)-(3) above.
__device__ __forceinline__ float enc_prepare_f32(float x) {
    x = __fadd_rn(x, 0.0f);                                   // (1) -0.0 -> +0.0
    uint32_t u = __float_as_uint(x);
    if (fabsf(x) < 0.001953125f) u &= 0x80000000u;            // (2) flush, keeps the sign of x<0
    uint32_t cap = (u & 0xFF800000u) | 0x006FFFFFu;           // (3) no carry
    u = min(u, cap);
    return __uint_as_float(u);
}

// two prepared fp32 -> 2 fp8 bytes (element 0 in the low byte).
__device__ __forceinline__ uint16_t cvt_f32x2_e4m3x2(float e0, float e1) {
    uint16_t r;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(e1), "f"(e0));
    return r;
}

__device__ __forceinline__ uint8_t enc1_f32(float x) {
    return (uint8_t)(cvt_f32x2_e4m3x2(enc_prepare_f32(x), 0.0f) & 0xFF);
}

// 4 fp32 -> 4 fp8 bytes packed little-endian.
__device__ __forceinline__ uint32_t enc4_f32(float a, float b, float c, float d) {
    uint32_t lo = cvt_f32x2_e4m3x2(enc_prepare_f32(a), enc_prepare_f32(b));
    uint32_t hi = cvt_f32x2_e4m3x2(enc_prepare_f32(c), enc_prepare_f32(d));
    return lo | (hi << 16);
}

// bf16x2 word (2 values) -> 2 fp8 bytes.  Repairs are done on the packed pair.
__device__ __forceinline__ uint16_t enc2_bf16x2(uint32_t w) {
    __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&w);
    const __nv_bfloat162 zero = __floats2bfloat162_rn(0.0f, 0.0f);
    h = __hadd2(h, zero);                                                  // (1)
    uint32_t z = *reinterpret_cast<uint32_t*>(&h);
    uint32_t small = __hlt2_mask(__habs2(h), __floats2bfloat162_rn(0.001953125f, 0.001953125f));
    z &= (~small | 0x80008000u);                                           // (2)
    uint32_t cap = (z & 0xFF80FF80u) | 0x006F006Fu;                        // (3) 7-bit fraction
    z = __vminu2(z, cap);
    float e0 = __uint_as_float(z << 16);
    float e1 = __uint_as_float(z & 0xFFFF0000u);
    return cvt_f32x2_e4m3x2(e0, e1);
}

// f16x2 word (2 values) -> 2 fp8 bytes; the conversion is taken straight from f16x2.
__device__ __forceinline__ uint16_t enc2_f16x2(uint32_t w) {
    __half2 h = *reinterpret_cast<__half2*>(&w);
    h = __hadd2(h, __floats2half2_rn(0.0f, 0.0f));                         // (1)
    uint32_t z = *reinterpret_cast<uint32_t*>(&h);
    uint32_t small = __hlt2_mask(__habs2(h), __floats2half2_rn(0.001953125f, 0.001953125f));
    z &= (~small | 0x80008000u);                                           // (2)
    uint32_t cap = (z & 0xFC00FC00u) | 0x037F037Fu;                        // (3) 10-bit fraction
    z = __vminu2(z, cap);
    uint16_t r;
    asm("cvt.rn.satfinite.e4m3x2.f16x2 %0, %1;" : "=h"(r) : "r"(z));
    return r;
}

}  // namespace fp8b
