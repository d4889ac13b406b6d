// Per-symbol arithmetic of the int16 soft demapper, shared by demod_kernels.cu and pusch_kernels.cu.
// Restated from lib/src/phy/modem/demod_soft.c (see demod_kernels.cu for the line references).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

__device__ __forceinline__ int sat16i(int v)
{
  return max(-32768, min(32767, v));
}
__device__ __forceinline__ int16_t wrap16(int v)
{
  return (int16_t)(uint16_t)(unsigned)v;
}
__device__ __forceinline__ int abs16w(int v) // _mm_abs_epi16 / (int16_t)abs(): |-32768| stays -32768
{
  return (int)wrap16(v < 0 ? -v : v);
}
// float -> int16 with C truncation semantics for in-range values (the reference's scalar tails)
__device__ __forceinline__ int trunc16(float f)
{
  return (int)wrap16(__float2int_rz(f));
}

// One symbol -> its bps soft bits (returned in o[0..bps)).  body: the symbol falls into the SIMD body of its reference
// call (rounding) rather than the scalar tail (truncation); for QPSK fpos0 = float position of s.x inside the call and
// fbody = 16 * (nfloats / 16).
__device__ __forceinline__ void demod_one(int mod, float2 s, bool body, uint32_t fpos0, uint32_t fbody, float qpsk_scale, int16_t o[6])
{
  if (mod == 3) {
    const int t1 = 432, t2 = 216; // (int16)(4*700/sqrtf(42)), (int16)(2*700/sqrtf(42))
    int       yr, yi, ar, ai;
    if (body) {
      yr = sat16i(__float2int_rn(s.x * -700.0f));
      yi = sat16i(__float2int_rn(s.y * -700.0f));
      ar = (int)wrap16(abs16w(yr) - t1);
      ai = (int)wrap16(abs16w(yi) - t1);
    } else {
      const int pr = trunc16(700.0f * s.x), pi = trunc16(700.0f * s.y);
      yr = (int)wrap16(-pr);
      yi = (int)wrap16(-pi);
      ar = (int)wrap16(abs16w(pr) - t1);
      ai = (int)wrap16(abs16w(pi) - t1);
    }
    o[0] = (int16_t)yr;
    o[1] = (int16_t)yi;
    o[2] = (int16_t)ar;
    o[3] = (int16_t)ai;
    o[4] = wrap16(abs16w(ar) - t2);
    o[5] = wrap16(abs16w(ai) - t2);
  } else if (mod == 2) {
    if (body) {
      const int yr = sat16i(__float2int_rn(s.x * -400.0f)), yi = sat16i(__float2int_rn(s.y * -400.0f));
      o[0]         = (int16_t)yr;
      o[1]         = (int16_t)yi;
      o[2]         = wrap16(abs16w(yr) - 252); // (int16)(2*400/sqrtf(10))
      o[3]         = wrap16(abs16w(yi) - 252);
    } else {
      const int   pr = trunc16(400.0f * s.x), pi = trunc16(400.0f * s.y);
      const float th = 2 * 400 / sqrtf(10.0f);
      o[0]           = wrap16(-pr);
      o[1]           = wrap16(-pi);
      o[2]           = (int16_t)trunc16((float)abs(pr) - th); // demod_soft.c:295: int - float, then truncated
      o[3]           = (int16_t)trunc16((float)abs(pi) - th);
    }
  } else { // QPSK
    const float v[2] = {s.x * qpsk_scale, s.y * qpsk_scale};
#pragma unroll
    for (int c = 0; c < 2; c++) {
      o[c] = fpos0 + c < fbody ? (int16_t)sat16i(__float2int_rz(v[c])) : (int16_t)trunc16(v[c]);
    }
  }
}

} // namespace b200
