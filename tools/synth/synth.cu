// Synthetic workload generator (device side), used by bench.py and the full-size GPU tests so that a 65,536-block
// batch (2.4 GB of LLRs) does not have to be produced on the host: random payload + CRC24B, LTE turbo encoding,
// BPSK over AWGN, quantisation to int16.  This is the recipe of SURVEY.md section 8(d) config 2 and follows what
// lib/src/phy/fec/turbo/test/turbodecoder_test.c:211-255 does on the CPU (encode -> +-1 + noise -> scale to int16),
// with an explicit clip so the inputs stay inside the range where the reference's generic decoder never wraps.
// Not on the decode path; nothing here is timed.  Built as tools/synth/libsrslte_b200_synth.so by srslte_b200/build.py: a
// test/bench helper OUTSIDE the product library (libsrslte_b200.so does not contain or link it).
#include <cuda_runtime.h>
#include <math.h>

#include <stdio.h>

#include <vector>

#include "lte_tables.h" // header-only host tables of the library (QPP parameters, CRC polynomials); nothing is linked

namespace b200 {

__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// One thread per code block: payload bits, CRC24B in the last 24 positions, both constituent encoders.
// coded[cb][3K+12] one bit per byte in the order of turbocoder.c:77-185.
__global__ void synth_encode_kernel(uint8_t* __restrict__ bits,  // [ncb][K] scratch, one bit per byte
                                    uint8_t* __restrict__ coded, // [ncb][3K+12]
                                    uint8_t* __restrict__ truth, // [ncb][K/8] packed payload+CRC (may be null)
                                    const uint16_t* __restrict__ qpp_fwd,
                                    uint32_t ncb,
                                    int      K,
                                    uint64_t seed,
                                    int      attach_crc)
{
  const uint32_t cb = blockIdx.x * blockDim.x + threadIdx.x;
  if (cb >= ncb) return;
  uint8_t* b = bits + (size_t)cb * K;
  uint8_t* c = coded + (size_t)cb * (3 * (size_t)K + 12);

  uint32_t  crc     = 0;
  const int payload = attach_crc ? K - 24 : K;
  for (int i = 0; i < payload; i += 64) {
    uint64_t r = mix64(seed ^ ((uint64_t)cb << 20) ^ (uint64_t)i);
    for (int j = 0; j < 64 && i + j < payload; j++) {
      uint32_t bit = (uint32_t)(r >> j) & 1u;
      b[i + j]     = (uint8_t)bit;
      uint32_t top = ((crc >> 23) & 1u) ^ bit;
      crc          = (crc << 1) & 0xFFFFFFu;
      if (top) crc ^= (CRC24B_POLY & 0xFFFFFFu);
    }
  }
  if (attach_crc) {
    for (int j = 0; j < 24; j++) b[payload + j] = (uint8_t)((crc >> (23 - j)) & 1u);
  }
  if (truth) {
    for (int i = 0; i < K / 8; i++) {
      uint32_t v = 0;
      for (int j = 0; j < 8; j++) v = (v << 1) | b[8 * i + j];
      truth[(size_t)cb * (K / 8) + i] = (uint8_t)v;
    }
  }
  uint32_t a0 = 0, a1 = 0, a2 = 0, e0 = 0, e1 = 0, e2 = 0; // shift registers of encoder 1 and 2
  for (int i = 0; i < K; i++) {
    uint32_t x  = b[i];
    uint32_t fb = x ^ a2 ^ a1;
    c[3 * i]    = (uint8_t)x;
    c[3 * i + 1] = (uint8_t)(a2 ^ a0 ^ fb);
    a2 = a1; a1 = a0; a0 = fb;
    uint32_t xi = b[qpp_fwd[i]];
    uint32_t fi = xi ^ e2 ^ e1;
    c[3 * i + 2] = (uint8_t)(e2 ^ e0 ^ fi);
    e2 = e1; e1 = e0; e0 = fi;
  }
  int k = 3 * K;
  for (int t = 0; t < 3; t++) {
    uint32_t x = a2 ^ a1, fb = 0;
    c[k++]     = (uint8_t)x;
    c[k++]     = (uint8_t)(a2 ^ a0 ^ fb);
    a2 = a1; a1 = a0; a0 = fb;
  }
  for (int t = 0; t < 3; t++) {
    uint32_t x = e2 ^ e1, fb = 0;
    c[k++]     = (uint8_t)x;
    c[k++]     = (uint8_t)(e2 ^ e0 ^ fb);
    e2 = e1; e1 = e0; e0 = fb;
  }
}

// One thread per LLR: y = (2c-1) + sigma*n ; llr = clip(rint(scale*y), +-clip)
__global__ void synth_channel_kernel(const uint8_t* __restrict__ coded,
                                     int16_t* __restrict__ llr,
                                     size_t   n,
                                     float    sigma,
                                     float    scale,
                                     int      clip,
                                     uint64_t seed)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t r  = mix64(seed ^ (0xA5A5ull << 48) ^ (uint64_t)i);
  float    u1 = ((float)(uint32_t)(r >> 40) + 1.0f) * (1.0f / 16777217.0f); // (0,1)
  float    u2 = (float)(uint32_t)((r >> 8) & 0xFFFFFFu) * (1.0f / 16777216.0f);
  float    g  = sqrtf(-2.0f * __logf(u1)) * __cosf(6.28318530718f * u2);
  float    y  = (coded[i] ? 1.0f : -1.0f) + sigma * g;
  int      q  = __float2int_rn(scale * y);
  q           = max(-clip, min(clip, q));
  llr[i]      = (int16_t)q;
}

} // namespace b200

using namespace b200;

#define SYNTH_CUDA_TRY(expr)                                                                                           \
  do {                                                                                                                 \
    cudaError_t e__ = (expr);                                                                                          \
    if (e__ != cudaSuccess) {                                                                                          \
      fprintf(stderr, "[b200_synth] %s: %s\n", #expr, cudaGetErrorString(e__));                                        \
      return -1;                                                                                                       \
    }                                                                                                                  \
  } while (0)

extern "C" __attribute__((visibility("default"))) int b200_synth_llr(int      device,
                                                                     int16_t* llr_dev,
                                                                     uint8_t* truth_dev,
                                                                     uint32_t ncb,
                                                                     uint32_t K,
                                                                     float    sigma,
                                                                     float    scale,
                                                                     int      clip,
                                                                     uint64_t seed,
                                                                     int      attach_crc,
                                                                     void*    stream)
{
  const int cb_idx = cb_index_exact(K);
  if (cb_idx < 0 || !llr_dev) {
    return -2;
  }
  SYNTH_CUDA_TRY(cudaSetDevice(device));
  std::vector<uint16_t> fwd, rev;
  qpp_tables(cb_idx, fwd, rev);
  uint16_t* d_qpp = nullptr;
  SYNTH_CUDA_TRY(cudaMalloc(&d_qpp, fwd.size() * sizeof(uint16_t)));
  SYNTH_CUDA_TRY(cudaMemcpy(d_qpp, fwd.data(), fwd.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
  cudaStream_t st   = (cudaStream_t)stream;
  const size_t nllr = 3 * (size_t)K + 12;
  uint8_t *    bits = nullptr, *coded = nullptr;
  // generate in slabs so the byte-per-bit scratch stays small next to a multi-GB LLR batch
  const uint32_t slab = 8192;
  SYNTH_CUDA_TRY(cudaMalloc(&bits, (size_t)slab * K));
  SYNTH_CUDA_TRY(cudaMalloc(&coded, (size_t)slab * nllr));
  for (uint32_t first = 0; first < ncb; first += slab) {
    const uint32_t n = (ncb - first) < slab ? (ncb - first) : slab;
    synth_encode_kernel<<<(n + 63) / 64, 64, 0, st>>>(bits,
                                                      coded,
                                                      truth_dev ? truth_dev + (size_t)first * (K / 8) : nullptr,
                                                      d_qpp,
                                                      n,
                                                      (int)K,
                                                      seed + 0x1000003ull * first,
                                                      attach_crc);
    const size_t tot = (size_t)n * nllr;
    synth_channel_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(
        coded, llr_dev + (size_t)first * nllr, tot, sigma, scale, clip, seed + 0x9E37ull * (first + 1));
  }
  SYNTH_CUDA_TRY(cudaStreamSynchronize(st));
  SYNTH_CUDA_TRY(cudaFree(bits));
  SYNTH_CUDA_TRY(cudaFree(coded));
  SYNTH_CUDA_TRY(cudaFree(d_qpp));
  SYNTH_CUDA_TRY(cudaGetLastError());
  return 0;
}


// ---------------------------------------------------------------------------------------------------------------
// PUSCH input for the multi-cell benchmark (BASELINE config 5): every subframe of a batch is one of `nbase` noiseless
// time-domain subframes (distinct payloads / RNTI / TTI / fading, synthesised on the host) plus its OWN realisation of
// complex AWGN, drawn here from a counter-based generator, quantised to the radio's int16 I/Q wire format:
//   out[sf][i] = sat16(rint(scale * (base[sf % nbase][i] + sigma * amp[sf % nbase] * (n1 + j n2) / sqrt(2))))
namespace b200 {
__global__ void synth_pusch_iq16_kernel(const float2* __restrict__ base, const float* __restrict__ amp, uint32_t nbase, uint32_t nsf,
                                        uint32_t sf_sz, float sigma, float scale, uint64_t seed, short2* __restrict__ out)
{
  const size_t n = (size_t)nsf * sf_sz, stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t sf = (uint32_t)(i / sf_sz), k = (uint32_t)(i % sf_sz), b = sf % nbase;
    const uint64_t r  = mix64(seed ^ (0x5A5Aull << 48) ^ (uint64_t)i);
    const float    u1 = ((float)(uint32_t)(r >> 40) + 1.0f) * (1.0f / 16777217.0f); // (0,1)
    const float    u2 = (float)(uint32_t)((r >> 8) & 0xFFFFFFu) * (1.0f / 16777216.0f);
    const float    m  = sqrtf(-2.0f * __logf(u1)) * sigma * amp[b] * 0.70710678f;
    float          sn, cs;
    __sincosf(6.28318530718f * u2, &sn, &cs);
    const float2 x  = base[(size_t)b * sf_sz + k];
    const int    re = max(-32768, min(32767, __float2int_rn(scale * (x.x + m * cs))));
    const int    im = max(-32768, min(32767, __float2int_rn(scale * (x.y + m * sn))));
    out[i]          = make_short2((short)re, (short)im);
  }
}
} // namespace b200

extern "C" __attribute__((visibility("default"))) int b200_synth_pusch_iq16(int device, const void* base_dev, const float* amp_dev,
                                                                            uint32_t nbase, uint32_t nsf, uint32_t sf_sz, float sigma,
                                                                            float scale, uint64_t seed, void* out_dev, void* stream)
{
  if (!base_dev || !amp_dev || !out_dev || nbase == 0) return -2;
  SYNTH_CUDA_TRY(cudaSetDevice(device));
  synth_pusch_iq16_kernel<<<148 * 16, 256, 0, (cudaStream_t)stream>>>((const float2*)base_dev, amp_dev, nbase, nsf, sf_sz, sigma, scale, seed,
                                                                       (short2*)out_dev);
  SYNTH_CUDA_TRY(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Issue rate of the packed-int16 instructions the turbo decoder is built from, measured in this process on this GPU: the
// denominator of the decoder's integer roofline.  Dependency-free streams, `warps_per_sm` resident warps per
// SM; lanes_per_clk_sm[0..2] = VIADD.16x2 alone, VIADDMNMX.S16x2 alone, the two interleaved 1:1 (they issue on different
// pipes); sm_clock_mhz = the SM clock observed while the streams ran (clock64 ticks / event time).
namespace b200 {
constexpr int UB_ILP = 8, UB_ITER = 4096;
template <int MODE>
__global__ void ubench_kernel(uint32_t* out, uint32_t seed, long long* cycles)
{
  uint32_t r[UB_ILP], b = seed | 1u, c = seed ^ 0x12345u;
#pragma unroll
  for (int i = 0; i < UB_ILP; i++) r[i] = threadIdx.x * 7 + i + seed;
  const long long t0 = clock64();
  for (int it = 0; it < UB_ITER; it++) {
#pragma unroll
    for (int i = 0; i < UB_ILP; i++) {
      if (MODE == 0 || (MODE == 2 && !(i & 1))) asm volatile("add.s16x2 %0, %1, %2;" : "=r"(r[i]) : "r"(r[i]), "r"(b));
      if (MODE == 1 || (MODE == 2 && (i & 1)))
        asm volatile("{.reg .b32 t; add.s16x2 t, %1, %2; max.s16x2 %0, t, %3;}" : "=r"(r[i]) : "r"(r[i]), "r"(b), "r"(c));
    }
  }
  const long long t1 = clock64();
  uint32_t        s  = 0;
#pragma unroll
  for (int i = 0; i < UB_ILP; i++) s ^= r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  // the SM is busy until its LAST warp is done: first start to last end over the warps of the block
  __shared__ long long s0, s1;
  if (threadIdx.x == 0) {
    s0 = t0;
    s1 = t1;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&s0, t0);
    atomicMax(&s1, t1);
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = s1 - s0;
}

template <int MODE>
static int ubench_run(int nsm, int warps_per_sm, uint32_t* out, long long* cyc, double* lanes, double* mhz)
{
  const int threads = warps_per_sm * 32;
  ubench_kernel<MODE><<<nsm, threads>>>(out, 3, cyc);
  cudaEvent_t e0, e1;
  SYNTH_CUDA_TRY(cudaEventCreate(&e0));
  SYNTH_CUDA_TRY(cudaEventCreate(&e1));
  SYNTH_CUDA_TRY(cudaEventRecord(e0));
  ubench_kernel<MODE><<<nsm, threads>>>(out, 5, cyc);
  SYNTH_CUDA_TRY(cudaEventRecord(e1));
  SYNTH_CUDA_TRY(cudaDeviceSynchronize());
  float ms = 0;
  SYNTH_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> h(nsm);
  SYNTH_CUDA_TRY(cudaMemcpy(h.data(), cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0;
  for (int i = 0; i < nsm; i++) avg += (double)h[i];
  avg /= nsm;
  *lanes = 32.0 * (double)UB_ITER * UB_ILP * warps_per_sm / avg;
  *mhz   = avg / (ms * 1e-3) / 1e6;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return 0;
}
} // namespace b200

extern "C" __attribute__((visibility("default"))) int b200_ubench_int16_issue(int device, int warps_per_sm, double* lanes_per_clk_sm,
                                                                              double* sm_clock_mhz)
{
  if (!lanes_per_clk_sm || warps_per_sm < 1 || warps_per_sm > 32) return -2;
  SYNTH_CUDA_TRY(cudaSetDevice(device));
  int nsm = 0;
  SYNTH_CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
  uint32_t*  out = nullptr;
  long long* cyc = nullptr;
  SYNTH_CUDA_TRY(cudaMalloc(&out, (size_t)nsm * warps_per_sm * 32 * 4));
  SYNTH_CUDA_TRY(cudaMalloc(&cyc, nsm * sizeof(long long)));
  double mhz[3] = {0, 0, 0};
  int    rc     = ubench_run<0>(nsm, warps_per_sm, out, cyc, &lanes_per_clk_sm[0], &mhz[0]);
  if (!rc) rc = ubench_run<1>(nsm, warps_per_sm, out, cyc, &lanes_per_clk_sm[1], &mhz[1]);
  if (!rc) rc = ubench_run<2>(nsm, warps_per_sm, out, cyc, &lanes_per_clk_sm[2], &mhz[2]);
  if (sm_clock_mhz) *sm_clock_mhz = mhz[2];
  cudaFree(out);
  cudaFree(cyc);
  return rc;
}
