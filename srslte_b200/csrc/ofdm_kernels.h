// OFDM receive kernel interface (ofdm_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

constexpr int OFDM_THREADS   = 128;
constexpr int OFDM_MAX_PASSES = 8;

// Everything the kernel needs about one srsran_ofdm_t configuration (device pointers into the object's tables)
struct OfdmPlanDev {
  int N;       // symbol_sz
  int R;       // nof_re = 12 * nof_prb
  int nsym;    // symbols per subframe: 14 normal CP, 12 extended
  int cp1;     // CP of the first symbol of a slot
  int cp2;     // CP of the other symbols
  int noff;    // window_offset_n (ofdm.c:133)
  int dc;      // 1: skip bin 0 in the upper half (ofdm.c:397,411)
  int sf_sz;   // 15 N
  int slot_sz; // 7.5 N
  int tps;     // threads cooperating on one symbol
  int npass;
  int radix[OFDM_MAX_PASSES];
  // generic batched DFT mode (srsran_dft_* entries): transform h reads in + h*idist, writes all N bins to out + h*odist
  int generic;
  int idist;
  int odist;
  int inverse; // 1: e^{+2 pi i kn/N} (conjugate in, conjugate out)
  float norm;  // 1/sqrt(N) when normalising, else 1 (already folded into ramp[]; the specialised kernel rebuilds the ramp)
  const float2* W;     // exp(-2 pi i m / N), m < N
  const float2* shift; // N entries: half-subcarrier rotation inside the FFT window, or nullptr
  const float2* ramp;  // R entries: window-offset phase fix x normalisation per output element, or nullptr
};

int launch_ofdm_rx(const OfdmPlanDev& p, const float2* in_dev, float2* out_dev, uint32_t nsf, int sm_count, cudaStream_t stream);

} // namespace b200
