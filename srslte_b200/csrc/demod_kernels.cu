// int16 soft demapper (QPSK / 16QAM / 64QAM), elementwise, bit-exact with the reference's x86 build.
//
// Restated from lib/src/phy/modem/demod_soft.c:
//   64QAM  demod_64qam_lte_s_sse :569-642   y = sat16(rne(-700 x)); |y|-432; ||y|-432|-216   (tail :629-642 truncates 700 x)
//   16QAM  demod_16qam_lte_s_sse :250-299   y = sat16(rne(-400 x)); |y|-252                  (tail :283-298)
//   QPSK   demod_qpsk_lte_s      :115-118 -> srsran_vec_convert_fi (vector_simd.c:436-472): sat16(trunc(x * -100 sqrt2)) in
//          blocks of 16 floats (AVX2 build: cvttps + packs, simd.h:1891-1895), remaining floats plain C cast
// "rne" = round to nearest even (cvtps_epi32 under the default MXCSR), "sat16" = packs_epi32 saturation.  The SIMD
// body / scalar tail split depends on the length of each reference call, so the entry takes the per-call group length.
#include <cuda_runtime.h>
#include <math.h>

#include <map>
#include <mutex>

#include "../../include/srslte_b200.h"
#include "b200_runtime.h"
#include "tdec_engine.h"

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

__global__ void demod_s_kernel(int mod, const float2* __restrict__ sym, int16_t* __restrict__ llr, uint32_t n, uint32_t group,
                               float qpsk_scale)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t g0   = (i / group) * group;               // first symbol of this reference call
  const uint32_t glen = min(group, n - g0);                // its length
  const uint32_t pos  = i - g0;
  const float2   s    = sym[i];
  if (mod == 3) {
    const bool body = pos < 4u * (glen / 4u);
    const int  t1 = 432, t2 = 216; // (int16)(4*700/sqrtf(42)), (int16)(2*700/sqrtf(42))
    int        yr, yi, ar, ai;
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
    int16_t* o = llr + 6 * (size_t)i;
    o[0]       = (int16_t)yr;
    o[1]       = (int16_t)yi;
    o[2]       = (int16_t)ar;
    o[3]       = (int16_t)ai;
    o[4]       = wrap16(abs16w(ar) - t2);
    o[5]       = wrap16(abs16w(ai) - t2);
  } else if (mod == 2) {
    const bool body = pos < 4u * (glen / 4u);
    int16_t*   o    = llr + 4 * (size_t)i;
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
    const uint32_t nf   = 2u * glen;
    const uint32_t body = 16u * (nf / 16u);
    int16_t*       o    = llr + 2 * (size_t)i;
    const float    v[2] = {s.x * qpsk_scale, s.y * qpsk_scale};
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const uint32_t fpos = 2u * pos + c;
      o[c] = fpos < body ? (int16_t)sat16i(__float2int_rz(v[c])) : (int16_t)trunc16(v[c]);
    }
  }
}

} // namespace b200

using namespace b200;

extern "C" SRSRAN_B200_API int srsran_b200_demod_soft_demodulate_s(int         device,
                                                                  int         modulation,
                                                                  const void* symbols,
                                                                  int16_t*    llr,
                                                                  uint32_t    nsymbols,
                                                                  uint32_t    symbols_per_call,
                                                                  uint32_t    flags,
                                                                  void*       stream)
{
  static const int bps_of[5] = {1, 2, 4, 6, 8};
  if (modulation < 1 || modulation > 3) {
    B200_LOG_ERROR("Invalid modulation %d", modulation); // demod_soft.c:889 (BPSK and 256QAM are not offloaded)
    return B200_ERROR;
  }
  if (!symbols || !llr) return B200_ERROR_INVALID_INPUTS;
  if (nsymbols == 0) return B200_SUCCESS;
  DeviceContext* ctx = device_context(device);
  if (!ctx) return B200_ERROR;
  B200_CUDA_TRY(cudaSetDevice(device));
  const uint32_t group = symbols_per_call ? symbols_per_call : nsymbols;
  const float    qs    = (float)(-100.0 * M_SQRT2); // -SCALE_SHORT_CONV_QPSK * M_SQRT2 converted to the float parameter
  const int      bps   = bps_of[modulation];
  const unsigned grid  = (nsymbols + 255) / 256;
  if (flags & SRSRAN_B200_FLAG_DEVICE_PTRS) {
    demod_s_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(modulation, (const float2*)symbols, llr, nsymbols, group, qs);
    g_kernel_launches++;
    B200_CUDA_TRY(cudaGetLastError());
    return B200_SUCCESS;
  }
  float2*  d_sym = nullptr;
  int16_t* d_llr = nullptr;
  B200_CUDA_TRY(cudaMalloc(&d_sym, (size_t)nsymbols * sizeof(float2)));
  B200_CUDA_TRY(cudaMalloc(&d_llr, (size_t)nsymbols * bps * sizeof(int16_t)));
  B200_CUDA_TRY(cudaMemcpy(d_sym, symbols, (size_t)nsymbols * sizeof(float2), cudaMemcpyHostToDevice));
  demod_s_kernel<<<grid, 256>>>(modulation, d_sym, d_llr, nsymbols, group, qs);
  g_kernel_launches++;
  B200_CUDA_TRY(cudaMemcpy(llr, d_llr, (size_t)nsymbols * bps * sizeof(int16_t), cudaMemcpyDeviceToHost));
  B200_CUDA_TRY(cudaFree(d_sym));
  B200_CUDA_TRY(cudaFree(d_llr));
  B200_CUDA_TRY(cudaGetLastError());
  return B200_SUCCESS;
}
