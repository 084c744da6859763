// rust_f32.cuh — Rust f32/usize leaf semantics as sm_100a device functions.
// Every TU that includes this is compiled with --fmad=false and without -use_fast_math: rustc never contracts
// a*b+c and euc's coverage/depth results depend on the exact rounding of every step.
#pragma once
#ifndef __CUDACC_RTC__
#include <cstdint>
#include <cuda_runtime.h>
#endif

namespace eucb {

// f32::min / f32::max (IEEE minNum / maxNum): if exactly one operand is NaN, the other one is returned.
__device__ __forceinline__ float r_min(float a, float b) { return a < b ? a : (b != b ? a : b); }
__device__ __forceinline__ float r_max(float a, float b) { return a > b ? a : (b != b ? a : b); }
// f32::fract = x - trunc(x) (keeps the sign of x)
__device__ __forceinline__ float r_fract(float x) { return x - truncf(x); }
// f32::rem_euclid
__device__ __forceinline__ float r_rem_euclid(float x, float rhs) {
    float r = fmodf(x, rhs);
    return r < 0.0f ? r + fabsf(rhs) : r;
}
// `f as usize` (64-bit): truncate, saturate, NaN -> 0.  cvt.rzi.u64.f32 has exactly these semantics.
__device__ __forceinline__ unsigned long long r_as_usize(float f) { return __float2ull_rz(f); }
// `f as usize` when the result is immediately clamped to a bound < 2^32 (saturation point is irrelevant then).
__device__ __forceinline__ uint32_t r_as_usize_clamped(float f, uint32_t lo, uint32_t hi) {
    uint32_t v = __float2uint_rz(f);
    return v < lo ? lo : (v > hi ? hi : v);
}
// `f as u8`
__device__ __forceinline__ uint32_t r_as_u8(float f) {
    uint32_t v;  // cvt.rzi to u8: truncate, saturate to [0, 255], NaN -> 0 (one F2IP.U8 instead of F2I.U32 + min)
    asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(v) : "f"(f));
    return v;
}
// `f.max(0.0).min(255.0) as u8`: the saturating cast alone has the same value for every input (NaN: max(NaN, 0) = 0 and
// the cast of NaN is 0; f < 0: both 0; f > 255: both 255; in between the clamp is the identity), so the two
// NaN-aware min/max are not materialised.
__device__ __forceinline__ uint32_t r_clamp255_as_u8(float f) { return r_as_u8(f); }
// vek Mat4<f32> * Vec4<f32> (column-major): cols[0]*x, then fused mul_add per remaining column (see DESIGN.md
// "Unpinned beliefs": vek 0.17 is not in the reference tree).  __fmaf_rn is never split or re-fused by --fmad.
__device__ __forceinline__ float4 mat4_mul_vec4(const float* __restrict__ m, float x, float y, float z, float w) {
    float4 o;
    float* op = &o.x;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float acc = m[r] * x;
        acc = __fmaf_rn(m[4 + r], y, acc);
        acc = __fmaf_rn(m[8 + r], z, acc);
        acc = __fmaf_rn(m[12 + r], w, acc);
        op[r] = acc;
    }
    return o;
}
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return ax * bx + ay * by + az * bz;
}
__device__ __forceinline__ uint32_t pack_le(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3) {
    return b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
}

}  // namespace eucb
