// Packed int16x2 arithmetic: two code blocks ride in the two halves of one 32-bit register.
//
// On sm_100a these map one-to-one onto the packed-halfword integer pipe:
//   add2    -> VIADD.16x2        (PTX add.s16x2, wraps modulo 2^16 per half, no carry across halves)
//   max2    -> VIMNMX.S16x2      (PTX max.s16x2, signed)
//   addmax2 -> VIADDMNMX.S16x2   (ptxas fuses add.s16x2 + max.s16x2)
//   pos2    -> VIMNMX.S16x2.RELU (min(x,1) clamped at 0: 1 where the half is > 0, else 0)
// Wrapping (not saturating) arithmetic is deliberate: the reference's generic decoder
// (lib/src/phy/fec/turbo/turbodecoder_gen.c:58-198) computes in plain int16_t, so parity on arbitrary inputs
// needs modulo-2^16 adds and signed compares.
//
// The same header compiles for the host (g++) with bit-identical emulation so the per-lane decoder logic can be
// unit-tested on a machine without a GPU (tests/test_tdec_emulation.py).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#else
#define B200_HD inline
#endif

namespace b200 {

B200_HD uint32_t pack2(int16_t lo, int16_t hi)
{
  return (uint32_t)(uint16_t)lo | ((uint32_t)(uint16_t)hi << 16);
}
B200_HD int16_t lo16(uint32_t v)
{
  return (int16_t)(uint16_t)(v & 0xFFFFu);
}
B200_HD int16_t hi16(uint32_t v)
{
  return (int16_t)(uint16_t)(v >> 16);
}

B200_HD uint32_t add2(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
  uint32_t r;
  asm("add.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
#else
  return ((a + b) & 0xFFFFu) | ((((a >> 16) + (b >> 16)) & 0xFFFFu) << 16);
#endif
}

B200_HD uint32_t max2(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
  uint32_t r;
  asm("max.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
#else
  int16_t l = lo16(a) > lo16(b) ? lo16(a) : lo16(b);
  int16_t h = hi16(a) > hi16(b) ? hi16(a) : hi16(b);
  return pack2(l, h);
#endif
}

// max(a + b, c)
B200_HD uint32_t addmax2(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
  uint32_t r;
  asm("{.reg .b32 t; add.s16x2 t, %1, %2; max.s16x2 %0, t, %3;}" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
#else
  return max2(add2(a, b), c);
#endif
}

B200_HD uint32_t neg2(uint32_t a)
{
  return add2(~a, 0x00010001u);
}

B200_HD uint32_t sub2(uint32_t a, uint32_t b)
{
  return add2(a, neg2(b));
}

// per half: 1 if the signed half is > 0 else 0
B200_HD uint32_t pos2(uint32_t a)
{
#if defined(__CUDA_ARCH__)
  uint32_t r;
  asm("min.s16x2.relu %0, %1, %2;" : "=r"(r) : "r"(a), "r"(0x00010001u));
  return r;
#else
  return (uint32_t)(lo16(a) > 0) | ((uint32_t)(hi16(a) > 0) << 16);
#endif
}

} // namespace b200
