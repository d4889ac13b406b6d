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
#include "demod_core.h"
#include "tdec_engine.h"

namespace b200 {

__global__ void demod_s_kernel(int mod, const float2* __restrict__ sym, int16_t* __restrict__ llr, uint32_t n, uint32_t group,
                               float qpsk_scale)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t g0   = (i / group) * group;               // first symbol of this reference call
  const uint32_t glen = min(group, n - g0);                // its length
  const uint32_t pos  = i - g0;
  const int      bps  = 2 * mod;
  int16_t        o[6];
  demod_one(mod, sym[i], pos < 4u * (glen / 4u), 2u * pos, 16u * (2u * glen / 16u), qpsk_scale, o);
  for (int b = 0; b < bps; b++) llr[(size_t)bps * i + b] = o[b];
}

// PUSCH glue for the identity-channel pipeline (BASELINE config 4): the data-carrying OFDM symbols of each subframe's
// resource grid are demapped as ONE reference call per subframe (pusch.c:449: srsran_demod_soft_demodulate_s over
// grant.nof_re symbols), then every soft bit is shifted right arithmetically by `shift` bits to bring the reference's
// 700x / 400x / 141x fixed-point scale into the generic int16 decoder's overflow-free envelope.
struct DataSymbols {
  uint8_t l[16]; // OFDM symbol index of the d-th data-carrying symbol
};

// One block per subframe walks its data symbols (no index divisions); the 2*mod soft bits of an element leave as
// 32-bit words.
__global__ void __launch_bounds__(256) pusch_demap_kernel(int mod, const float2* __restrict__ grid, int16_t* __restrict__ llr,
                                                          uint32_t nof_symbols, uint32_t nof_re, DataSymbols ds, uint32_t n_data_sym,
                                                          int shift, float qpsk_scale)
{
  const uint32_t sf     = blockIdx.x;
  const uint32_t per_sf = n_data_sym * nof_re;
  const uint32_t body = 4u * (per_sf / 4u), fbody = 16u * (2u * per_sf / 16u);
  for (uint32_t d = 0; d < n_data_sym; d++) {
    const float2* src = grid + ((size_t)sf * nof_symbols + ds.l[d]) * nof_re;
    uint32_t*     dst = reinterpret_cast<uint32_t*>(llr + ((size_t)sf * per_sf + (size_t)d * nof_re) * (size_t)(2 * mod));
    for (uint32_t re = threadIdx.x; re < nof_re; re += blockDim.x) {
      const uint32_t pos = d * nof_re + re; // symbol index inside the subframe's reference call
      int16_t        o[6];
      demod_one(mod, __ldcs(&src[re]), pos < body, 2u * pos, fbody, qpsk_scale, o);
#pragma unroll
      for (int w = 0; w < 3; w++) {
        if (w < mod) {
          const uint32_t lo = (uint16_t)(int16_t)(o[2 * w] >> shift), hi = (uint16_t)(int16_t)(o[2 * w + 1] >> shift);
          dst[(size_t)re * mod + w] = lo | (hi << 16);
        }
      }
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

extern "C" SRSRAN_B200_API int srsran_b200_pusch_demap_batch(int         device,
                                                            int         modulation,
                                                            const void* grid,
                                                            int16_t*    llr,
                                                            uint32_t    nsf,
                                                            uint32_t    nof_symbols,
                                                            uint32_t    nof_re,
                                                            uint32_t    sym_mask,
                                                            uint32_t    llr_shift,
                                                            uint32_t    flags,
                                                            void*       stream)
{
  if (modulation < 1 || modulation > 3) {
    B200_LOG_ERROR("Invalid modulation %d", modulation);
    return B200_ERROR;
  }
  if (!grid || !llr || nof_symbols == 0 || nof_symbols > 14 || nof_re == 0 || llr_shift > 15) return B200_ERROR_INVALID_INPUTS;
  if (!(flags & SRSRAN_B200_FLAG_DEVICE_PTRS)) {
    B200_LOG_ERROR("srsran_b200_pusch_demap_batch works on device buffers (it sits between two device-side stages)");
    return B200_ERROR_INVALID_INPUTS;
  }
  sym_mask &= (1u << nof_symbols) - 1u;
  const uint32_t nd = (uint32_t)__builtin_popcount(sym_mask);
  if (nsf == 0 || nd == 0) return B200_SUCCESS;
  DeviceContext* ctx = device_context(device);
  if (!ctx) return B200_ERROR;
  B200_CUDA_TRY(cudaSetDevice(device));
  const float qs = (float)(-100.0 * M_SQRT2);
  DataSymbols ds = {};
  for (uint32_t l = 0, d = 0; l < nof_symbols; l++) {
    if ((sym_mask >> l) & 1u) ds.l[d++] = (uint8_t)l;
  }
  pusch_demap_kernel<<<nsf, 256, 0, (cudaStream_t)stream>>>(modulation, (const float2*)grid, llr, nof_symbols, nof_re, ds, nd,
                                                          (int)llr_shift, qs);
  g_kernel_launches++;
  B200_CUDA_TRY(cudaGetLastError());
  return B200_SUCCESS;
}
