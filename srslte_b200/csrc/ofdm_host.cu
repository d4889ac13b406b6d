// Host side of the OFDM receiver: plan construction (the reference's ofdm_init_mbsfn_ arithmetic), table generation and
// the batched entry.  Reference: lib/src/phy/dft/ofdm.c:38-212,334-362; lib/include/srsran/phy/common/phy_common.h:113-134;
// lib/src/phy/common/phy_common.c:322-385.
#include <math.h>

#include <complex>
#include <new>
#include <vector>

#include "../../include/srslte_b200.h"
#include "b200_runtime.h"
#include "ofdm_kernels.h"
#include "tdec_engine.h"

namespace b200 {

static bool g_standard_symbol_sz = false; // srsran_use_standard_symbol_size (phy_common.c:322)

int symbol_sz_for_prb(uint32_t nof_prb)
{
  if (nof_prb == 0) return -1;
  if (g_standard_symbol_sz) { // srsran_symbol_sz_power2, phy_common.c:342-359
    if (nof_prb <= 6) return 128;
    if (nof_prb <= 15) return 256;
    if (nof_prb <= 25) return 512;
    if (nof_prb <= 50) return 1024;
    if (nof_prb <= 75) return 1536;
    if (nof_prb <= 110) return 2048;
    return -1;
  }
  if (nof_prb <= 6) return 128; // phy_common.c:366-381
  if (nof_prb <= 15) return 256;
  if (nof_prb <= 25) return 384;
  if (nof_prb <= 50) return 768;
  if (nof_prb <= 75) return 1024;
  if (nof_prb <= 110) return 1536;
  return -1;
}

static int cp_len(int c, int N)
{
  return (int)ceilf(((float)c * (float)N) / 2048.0f); // SRSRAN_CP_LEN, phy_common.h:125
}

struct OfdmEngine {
  DeviceContext* ctx = nullptr;
  OfdmPlanDev    plan{};
  float2 *       dW = nullptr, *dShift = nullptr, *dRamp = nullptr;
  int            sm_count = 148;
  cudaStream_t   stream   = nullptr;
  DeviceArena    io;
  srsran_b200_ofdm_cfg_t cfg{};

  void free_tables()
  {
    if (dW) cudaFree(dW);
    if (dShift) cudaFree(dShift);
    if (dRamp) cudaFree(dRamp);
    dW = dShift = dRamp = nullptr;
  }
  void destroy()
  {
    if (ctx) cudaSetDevice(ctx->device);
    free_tables();
    if (stream) cudaStreamDestroy(stream);
    io.release();
  }

  int configure(const srsran_b200_ofdm_cfg_t& c)
  {
    int N = (int)c.symbol_sz;
    if (N == 0) {
      N = symbol_sz_for_prb(c.nof_prb);
      if (N <= 0) {
        B200_LOG_ERROR("Invalid number of PRB %u", c.nof_prb); // ofdm.c:43
        return B200_ERROR;
      }
    }
    // factor N into the register radices the kernel has
    int       radix[OFDM_MAX_PASSES];
    const int npass = N >= 16 ? fft_factorise(N, radix) : 0;
    if (npass == 0 || 12 * (int)c.nof_prb > N || c.nof_prb == 0) {
      B200_LOG_ERROR("unsupported OFDM size: symbol_sz=%d nof_prb=%u (sizes 2^a 3^b 5^c, nof_re <= symbol_sz)", N, c.nof_prb);
      return B200_ERROR;
    }
    cfg          = c;
    const bool ext = c.cp_ext != 0;
    plan.N       = N;
    plan.R       = 12 * (int)c.nof_prb;
    plan.nsym    = ext ? 12 : 14;
    plan.cp1     = ext ? cp_len(512, N) : cp_len(160, N);
    plan.cp2     = ext ? cp_len(512, N) : cp_len(144, N);
    plan.sf_sz   = 15 * N;
    plan.slot_sz = 15 * N / 2;
    plan.npass   = npass;
    for (int i = 0; i < OFDM_MAX_PASSES; i++) plan.radix[i] = radix[i];
    int tps = N / 16;
    if (tps < 8) tps = 8;
    if (tps > OFDM_THREADS) tps = OFDM_THREADS;
    plan.tps  = tps;
    plan.noff = 0;
    float off = c.rx_window_offset;
    if (isnormal(off)) { // ofdm.c:130-133
      if (off < 0) off = 0;
      if (off > 1) {
        B200_LOG_ERROR("rx_window_offset %f > 1 would start the DFT window before the symbol's own CP", off);
        return B200_ERROR;
      }
      plan.noff = (int)roundf((float)plan.cp2 * off);
    }
    const bool shift = isnormal(c.freq_shift_f);
    plan.dc          = (!c.keep_dc && !shift) ? 1 : 0; // ofdm.c:209

    B200_CUDA_TRY(cudaSetDevice(ctx->device));
    free_tables();
    std::vector<float2> W(N);
    for (int m = 0; m < N; m++) {
      double a = -2.0 * M_PI * (double)m / (double)N;
      W[m]     = make_float2((float)cos(a), (float)sin(a));
    }
    B200_CUDA_TRY(cudaMalloc(&dW, N * sizeof(float2)));
    B200_CUDA_TRY(cudaMemcpy(dW, W.data(), N * sizeof(float2), cudaMemcpyHostToDevice));
    plan.W     = dW;
    plan.shift = nullptr;
    plan.ramp  = nullptr;
    plan.norm  = c.normalize ? 1.0f / sqrtf((float)N) : 1.0f;
    if (shift) {
      // ofdm.c:347-355: shift[t] = cexpf(I 2 pi (t - cplen) f / N); inside the FFT window t - cplen = n - noff for every symbol
      std::vector<float2> S(N);
      for (int n = 0; n < N; n++) {
        float  rel = (float)(n - plan.noff);
        double arg = 2.0 * M_PI * (double)rel * (double)c.freq_shift_f / (double)N;
        float  af  = (float)arg; // the reference hands a float-precision angle to cexpf
        S[n]       = make_float2((float)cos((double)af), (float)sin((double)af));
      }
      B200_CUDA_TRY(cudaMalloc(&dShift, N * sizeof(float2)));
      B200_CUDA_TRY(cudaMemcpy(dShift, S.data(), N * sizeof(float2), cudaMemcpyHostToDevice));
      plan.shift = dShift;
    }
    if (plan.noff || c.normalize) {
      // ofdm.c:134-136,405-407,414-416: tmp[i] *= cexpf(I pi 2 noff i / N), then the kept bins, then 1/sqrt(N)
      std::vector<float2> Rp(plan.R);
      const float         norm = c.normalize ? 1.0f / sqrtf((float)N) : 1.0f;
      for (int r = 0; r < plan.R; r++) {
        int    bin = r < plan.R / 2 ? N - plan.R / 2 + r : plan.dc + (r - plan.R / 2);
        float2 v   = make_float2(1.f, 0.f);
        if (plan.noff) {
          double arg = M_PI * 2.0 * (double)(float)plan.noff * (double)(float)bin / (double)(float)N;
          float  af  = (float)arg;
          v          = make_float2((float)cos((double)af), (float)sin((double)af));
        }
        Rp[r] = make_float2(v.x * norm, v.y * norm);
      }
      B200_CUDA_TRY(cudaMalloc(&dRamp, plan.R * sizeof(float2)));
      B200_CUDA_TRY(cudaMemcpy(dRamp, Rp.data(), plan.R * sizeof(float2), cudaMemcpyHostToDevice));
      plan.ramp = dRamp;
    }
    return B200_SUCCESS;
  }

  int init(int device, const srsran_b200_ofdm_cfg_t& c)
  {
    ctx = device_context(device);
    if (!ctx) {
      B200_LOG_ERROR("no usable CUDA device %d (this library has no CPU fallback)", device);
      return B200_ERROR;
    }
    B200_CUDA_TRY(cudaSetDevice(device));
    B200_CUDA_TRY(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
    B200_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    return configure(c);
  }

  int run(const float2* in, float2* out, uint32_t nsf, uint32_t flags, cudaStream_t user)
  {
    if (!in || !out) return B200_ERROR_INVALID_INPUTS;
    if (nsf == 0) return B200_SUCCESS;
    B200_CUDA_TRY(cudaSetDevice(ctx->device));
    OfdmPlanDev plan = this->plan; // per call: the sample format is a property of the call, not of the object
    plan.iq16        = (flags & SRSRAN_B200_FLAG_IQ_INT16) ? 1 : 0;
    plan.iq_scale    = 1.0f / 32768.0f;
    const size_t in_b = (size_t)nsf * plan.sf_sz * (plan.iq16 ? 2 * sizeof(int16_t) : sizeof(float2)),
                 out_b = (size_t)nsf * plan.nsym * plan.R * sizeof(float2);
    if (flags & SRSRAN_B200_FLAG_DEVICE_PTRS) {
      int rc = launch_ofdm_rx(plan, in, out, nsf, sm_count, user);
      g_kernel_launches++;
      return rc;
    }
    if (io.reserve(in_b + out_b + 1024) != B200_SUCCESS) return B200_ERROR;
    io.reset();
    float2* d_in  = (float2*)io.take(in_b);
    float2* d_out = (float2*)io.take(out_b);
    B200_CUDA_TRY(cudaMemcpyAsync(d_in, in, in_b, cudaMemcpyHostToDevice, stream));
    int rc = launch_ofdm_rx(plan, d_in, d_out, nsf, sm_count, stream);
    g_kernel_launches++;
    B200_CUDA_TRY(cudaMemcpyAsync(out, d_out, out_b, cudaMemcpyDeviceToHost, stream));
    B200_CUDA_TRY(cudaStreamSynchronize(stream));
    return rc;
  }
};

} // namespace b200

using namespace b200;

struct srsran_b200_ofdm {
  OfdmEngine eng;
};

extern "C" {

void srsran_b200_use_standard_symbol_size(int enabled)
{
  g_standard_symbol_sz = enabled != 0;
}

int srsran_b200_symbol_sz(uint32_t nof_prb)
{
  return symbol_sz_for_prb(nof_prb);
}

int srsran_b200_ofdm_rx_init(srsran_b200_ofdm_t** q, int device, const srsran_b200_ofdm_cfg_t* cfg)
{
  if (!q || !cfg) return B200_ERROR_INVALID_INPUTS;
  *q                    = nullptr;
  srsran_b200_ofdm_t* h = new (std::nothrow) srsran_b200_ofdm_t();
  if (!h) return B200_ERROR;
  if (h->eng.init(device, *cfg) != B200_SUCCESS) {
    h->eng.destroy();
    delete h;
    return B200_ERROR;
  }
  *q = h;
  return B200_SUCCESS;
}

int srsran_b200_ofdm_rx_reconfigure(srsran_b200_ofdm_t* q, const srsran_b200_ofdm_cfg_t* cfg)
{
  if (!q || !cfg) return B200_ERROR_INVALID_INPUTS;
  return q->eng.configure(*cfg);
}

void srsran_b200_ofdm_rx_free(srsran_b200_ofdm_t* q)
{
  if (q) {
    q->eng.destroy();
    delete q;
  }
}

int srsran_b200_ofdm_rx_geometry(const srsran_b200_ofdm_t* q, uint32_t* symbol_sz, uint32_t* sf_sz, uint32_t* nof_symbols, uint32_t* nof_re)
{
  if (!q) return B200_ERROR_INVALID_INPUTS;
  if (symbol_sz) *symbol_sz = (uint32_t)q->eng.plan.N;
  if (sf_sz) *sf_sz = (uint32_t)q->eng.plan.sf_sz;
  if (nof_symbols) *nof_symbols = (uint32_t)q->eng.plan.nsym;
  if (nof_re) *nof_re = (uint32_t)q->eng.plan.R;
  return B200_SUCCESS;
}

int srsran_b200_ofdm_rx_sf_batch(srsran_b200_ofdm_t* q, const void* in, void* out, uint32_t nsf, uint32_t flags, void* stream)
{
  if (!q) return B200_ERROR_INVALID_INPUTS;
  return q->eng.run((const float2*)in, (float2*)out, nsf, flags, (cudaStream_t)stream);
}

} // extern "C"
