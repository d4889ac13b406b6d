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
  int   iq16;     // OFDM mode: the input samples are int16 I/Q pairs (the radio's wire format); value = int16 * iq_scale
  float iq_scale; // 1/32768: what the host-side conversion to float does (power of two, so the float result is the same)
  float gscale; // generic mode: every output bin is multiplied by this when non-zero (1/sqrt(N) of dft_fftw.c:343-350)
  // PUSCH transform de-precoding mode (generic = 2): transform h = (subframe sf = h / pusch_nd, data symbol d = h % pusch_nd) reads
  // the 12*L_prb allocated elements of OFDM symbol pusch_l[d] out of the subframe's resource grid, EQUALISED on the fly with
  // the slot's channel estimate (precoding.c:182-305: y conj(h) / (|h|^2 + noise)), and writes N bins to out + h*N.
  int           pusch_nd;     // data-carrying symbols per subframe (12 normal CP / 10 extended)
  unsigned char pusch_l[16];  // their OFDM symbol indices
  int           grid_nsym;    // symbols per subframe in the grid
  int           grid_R;       // elements per symbol in the grid (12 * cell nof_prb)
  int           grid_off;     // first allocated element (12 * n_prb)
  const float2* eq_ce;        // [nsf][2 slots][N] channel estimates
  const float*  eq_noise;     // noise estimate of subframe sf at eq_noise[sf * eq_noise_stride]
  int           eq_noise_stride;
  float norm;  // 1/sqrt(N) when normalising, else 1 (already folded into ramp[]; the specialised kernel rebuilds the ramp)
  const float2* W;     // exp(-2 pi i m / N), m < N
  const float2* shift; // N entries: half-subcarrier rotation inside the FFT window, or nullptr
  const float2* ramp;  // R entries: window-offset phase fix x normalisation per output element, or nullptr
};

// Split N into the register radices of the FFT core (16, 8, 4, 2, 3, 5), at least two passes.  Returns the number of
// passes, or 0 when N has another prime factor / needs more than OFDM_MAX_PASSES passes.
inline int fft_factorise(int N, int radix[OFDM_MAX_PASSES])
{
  static const int cand[6] = {16, 8, 4, 2, 3, 5};
  int              n = 0, rem = N;
  for (int i = 0; i < 6; i++) {
    while (rem % cand[i] == 0 && rem > 1) {
      if (n == OFDM_MAX_PASSES) return 0;
      radix[n++] = cand[i];
      rem /= cand[i];
    }
  }
  if (rem != 1 || n == 0) return 0;
  if (n == 1) { // a single pass would have to be first and last at once: split it
    const int r = radix[0];
    if (r == 16) { radix[0] = 4; radix[1] = 4; }
    else if (r == 8) { radix[0] = 4; radix[1] = 2; }
    else if (r == 4) { radix[0] = 2; radix[1] = 2; }
    else return 0; // 2, 3, 5 points: not worth a kernel
    n = 2;
  }
  for (int i = n; i < OFDM_MAX_PASSES; i++) radix[i] = 1;
  return n;
}

int launch_ofdm_rx(const OfdmPlanDev& p, const float2* in_dev, float2* out_dev, uint32_t nsf, int sm_count, cudaStream_t stream);

} // namespace b200
