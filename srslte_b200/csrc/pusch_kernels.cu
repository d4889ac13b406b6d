// PUSCH receive chain between the OFDM demodulator and the rate de-matcher, for a batch of subframes:
//
//   DMRS known sequence        srsran_refsignal_dmrs_pusch_gen       lib/src/phy/ch_estimation/refsignal_ul.c:95-181,227-249,337-357
//                              srsran_zc_sequence_generate_lte       lib/src/phy/common/zc_sequence.c:205-235,273-300
//   channel estimation         srsran_chest_ul_estimate_pusch        lib/src/phy/ch_estimation/chest_ul.c:197-357,370-400
//                              srsran_conv_same_cf (3-tap smoothing) lib/src/phy/utils/convolution.c:181-218
//   equaliser                  srsran_predecoding_single             lib/src/phy/mimo/precoding.c:182-305,357  } one kernel: the FFT core of
//   transform de-precoding     srsran_dft_precoding                  lib/src/phy/dft/dft_precoding.c:114-126   } ofdm_kernels.cu in PUSCH mode
//   soft demapping             srsran_demod_soft_demodulate_s        lib/src/phy/modem/demod_soft.c:871          }
//   descrambling               srsran_sequence_pusch_apply_s         lib/src/phy/phch/sequences.c:120-147        } one kernel
//   UL-SCH de-interleaving     ulsch_deinterleave                    lib/src/phy/phch/sch.c:660-681,993-1020     }
//
// in the order lib/src/phy/phch/pusch.c:392-456 (srsran_pusch_decode) and sch.c:1121-1190 (srsran_ulsch_decode) call them.
//   control information        uci_decode_ri_ack, ulsch_deinterleave  lib/src/phy/phch/sch.c:993-1119 (UCI = true instantiation + uci_host.cu)
// Scope: one receive antenna (like pusch.c:413), no intra-subframe hopping (the reference's estimator refuses it too,
// chest_ul.c:330), every allocation srsran_dft_precoding_valid_prb accepts.
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "../../include/srslte_b200.h"
#include "b200_runtime.h"
#include "demod_core.h"
#include "ofdm_kernels.h"
#include "lte_tables.h"
#include "tdec_engine.h"
#include "uci_host.h"

namespace b200 {

#include "dmrs_phi_table.inc"

struct PuschSfParam {
  uint32_t c_init;   // scrambling seed of the subframe: (rnti << 14) + (sf_idx << 9) + cell_id  (sequences.c:120-123)
  uint32_t dmrs_idx; // n_dmrs * 10 + sf_idx: row of the DMRS table
};

// Small per-call parameter arrays reach the device through this kernel, which reads the page-locked host buffer directly
// (unified addressing), NOT through a host-to-device copy: a copy would queue on the one H2D engine behind the megabytes of
// samples that a pipelining caller has already submitted for its later chunks, and the chunk's kernels would wait for all of
// them (measured: the front end of a 4096-subframe batch started only after the last sample copy, enb_ul.cu).
__global__ void pusch_words_to_device_kernel(const uint32_t* __restrict__ host_mapped, uint32_t* __restrict__ dev, uint32_t nwords)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nwords) dev[i] = host_mapped[i];
}

// ---------------------------------------------------------------------------------------------------------------
// Channel estimation: one block per subframe.  LS estimates of the two DMRS symbols (received x conj(known)), 3-tap
// smoothing with the reference's extrapolated edges, noise from the difference between smoothed and raw estimates,
// received pilot power and the slot-to-slot phase (CFO).  Output: the two slots' estimates [nsf][2][M] (the reference
// copies them to every symbol of the slot, chest_ul.c:246-259: the equaliser reads them per slot instead) and
// meas[nsf][4] = {noise_estimate, snr, cfo_hz, ta_us = 0}.
__global__ void __launch_bounds__(256) pusch_chest_kernel(const float2* __restrict__ grid, const float2* __restrict__ dmrs_tab,
                                                          const PuschSfParam* __restrict__ prm, float2* __restrict__ ce,
                                                          float* __restrict__ meas, int nsym, int R, int off, int M, float w, float noise_cal)
{
  extern __shared__ __align__(16) float2 ls[]; // [2][M]
  __shared__ float red[4][8];
  const uint32_t sf   = blockIdx.x;
  const float2*  known = dmrs_tab + (size_t)prm[sf].dmrs_idx * 2 * M;
  const int      half = nsym / 2;
  float          rxpow = 0.f, dre = 0.f, dim = 0.f, npow = 0.f;
  for (int s = 0; s < 2; s++) {
    const float2* y = grid + ((size_t)sf * nsym + (size_t)((s + 1) * half - 4)) * R + off; // SRSRAN_REFSIGNAL_UL_L (refsignal_ul.h:43)
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
      const float2 v = y[i], r = known[s * M + i];
      ls[s * M + i]  = make_float2(v.x * r.x + v.y * r.y, v.y * r.x - v.x * r.y); // srsran_vec_prod_conj_ccc
      rxpow += v.x * v.x + v.y * v.y;
    }
  }
  __syncthreads();
  const float f0 = w, f1 = 1.f - 2.f * w;
  for (int s = 0; s < 2; s++) {
    const float2* in = ls + s * M;
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
      // convolution.c:187-204: samples beyond the edges are 3 in[1] - 2 in[0] and 3 in[M-1] - 2 in[M-2]
      const float2 c = in[i];
      float2       a, b;
      if (i > 0) a = in[i - 1];
      else a = make_float2(3.f * in[1].x - 2.f * in[0].x, 3.f * in[1].y - 2.f * in[0].y);
      if (i < M - 1) b = in[i + 1];
      else b = make_float2(3.f * in[M - 1].x - 2.f * in[M - 2].x, 3.f * in[M - 1].y - 2.f * in[M - 2].y);
      const float2 o = make_float2(a.x * f0 + c.x * f1 + b.x * f0, a.y * f0 + c.y * f1 + b.y * f0);
      ce[((size_t)sf * 2 + s) * M + i] = o;
      const float ex = o.x - c.x, ey = o.y - c.y;
      npow += ex * ex + ey * ey;
      if (s == 0) { // CFO: sum ls0 conj(ls1) (chest_ul.c:252-255)
        const float2 o1 = in[M + i];
        dre += c.x * o1.x + c.y * o1.y;
        dim += c.y * o1.x - c.x * o1.y;
      }
    }
  }
  float v[4] = {npow, rxpow, dre, dim};
#pragma unroll
  for (int k = 0; k < 4; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xFFFFFFFFu, v[k], o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < 4; k++)
      for (int q = 0; q < (int)(blockDim.x >> 5); q++) t[k] += red[k][q];
    // chest_ul.c:197-222: mean over the slots of the per-slot average power, calibrated for the 3-tap filter
    const float noise = (t[0] / (float)M) / 2.f / noise_cal;
    const bool  ok    = isfinite(noise) && fabsf(noise) >= 1.17549435e-38f; // isnormal()
    meas[4 * sf + 0]  = noise;
    meas[4 * sf + 1]  = ok ? (t[1] / (float)(2 * M)) / noise : nanf("");
    meas[4 * sf + 2]  = atan2f(t[3], t[2]) / (2.0f * 3.14159265358979323846f * 0.0005f);
    meas[4 * sf + 3]  = 0.f;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Scrambling sequences: GOLD_CHUNKS threads per subframe.  c(n) = x1(n+1600) ^ x2(n+1600) (TS 36.211 7.2; sequence.c:33-120).
// x1 does not depend on the seed: its words come from a table.  x2 is linear in the 31 seed bits, so every thread starts at its
// own chunk of the sequence from the XOR of per-seed-bit jump states, then steps 16 bits at a time (all taps of the
// recurrence x2(n+31) = x2(n+3)^x2(n+2)^x2(n+1)^x2(n) lie inside the 31-bit state for 16 new bits).  The chunks are short
// (a 20 MHz 64QAM subframe is 2701 words: 22 per thread) because the per-thread chain is serial.
constexpr uint32_t GOLD_CHUNKS = 128;
__global__ void __launch_bounds__(256) pusch_gold_kernel(const PuschSfParam* __restrict__ prm, const uint32_t* __restrict__ x1w,
                                                         const uint32_t* __restrict__ jump, uint32_t* __restrict__ seq, uint32_t nsf,
                                                         uint32_t nwords, uint32_t wpl)
{
  extern __shared__ uint32_t gold_stage[]; // [2][nwords]: the block's two subframes, written out coalesced at the end
  const uint32_t half = threadIdx.x / GOLD_CHUNKS, lane = threadIdx.x % GOLD_CHUNKS, sf = blockIdx.x * 2u + half;
  if (sf < nsf) {
    const uint32_t seed = prm[sf].c_init;
    uint32_t       s    = 0;
#pragma unroll
    for (int b = 0; b < 31; b++) s ^= ((seed >> b) & 1u) ? jump[lane * 31 + b] : 0u;
    const uint32_t w0 = lane * wpl, w1 = min(w0 + wpl, nwords);
    for (uint32_t w = w0; w < w1; w++) {
      uint32_t out = s & 0xFFFFu;
      uint32_t f   = (s ^ (s >> 1) ^ (s >> 2) ^ (s >> 3)) & 0xFFFFu;
      s            = (s >> 16) | (f << 15);
      out |= (s & 0xFFFFu) << 16;
      f = (s ^ (s >> 1) ^ (s >> 2) ^ (s >> 3)) & 0xFFFFu;
      s = (s >> 16) | (f << 15);
      gold_stage[half * nwords + w] = out;
    }
  }
  __syncthreads();
  const uint32_t n = min(2u, nsf - blockIdx.x * 2u) * nwords;
  uint32_t*      o = seq + (size_t)blockIdx.x * 2u * nwords;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) o[i] = gold_stage[i] ^ x1w[i < nwords ? i : i - nwords];
}

// ---------------------------------------------------------------------------------------------------------------
// Soft demapping + arithmetic shift + descrambling + UL-SCH de-interleaving.  The demapper sees the subframe's nd*M
// de-precoded symbols as ONE reference call (pusch.c:421: vector body / scalar tail split over grant.nof_re); soft bit
// n = (i*M + j)*Qm + k of data symbol i, subcarrier j changes sign when c(n) = 1 (int16 wrap, sequence.c:521-532) and lands
// at g[(j*nd + i)*Qm + k] (sch.c:660-681 with rows = M, cols = nd, no RI bits).  A block stages a tile of 64 subcarriers x nd
// symbols in shared memory (coalesced row reads), then walks it in output order so the stores are contiguous.
constexpr int DEMOD_TJ = 64;
//
// UCI = true (TS 36.212 5.2.2.7-5.2.2.8; sch.c:993-1119, uci.c:346-393,637-690): the interleaver matrix has M rows (subcarriers) and
// nd columns (data symbols).  HARQ-ACK symbol a sits in row M-1-a/4 of the column set {2,3,8,9} ({1,2,6,7} with the extended
// prefix) taken in the order (3a)%4, RI symbol r likewise in {1,4,7,10} ({0,3,5,8}).  The de-interleaved stream leaves the RI
// positions out (everything after them moves up) and carries zeros at the HARQ-ACK positions; the fields' own soft bits go, at
// the demapper's scale (no llr_shift), to the subframe's row of `uci_llr`: [Q_ack*Qm | Q_ri*Qm | Q_cqi*Qm], the last being a copy of
// the front of the stream.  In the 1-bit forms the repeated bit's scrambling is undone here (uci.c:678-682: it was scrambled
// like its predecessor).  Element 0 of the stream repeats the reference's scatter quirk: see include/srslte_b200.h.
struct PuschUciSf {
  uint32_t Q_ack, Q_ri, Q_cqi;
  uint32_t flags; // bit 0: 1-bit HARQ-ACK, bit 1: 1-bit RI
};
struct PuschUciArgs {
  const PuschUciSf* sf;      // [nsf]
  int16_t*          llr;     // [nsf][stride]
  uint32_t          stride;  // int16 per subframe
  uint32_t          off_ri, off_cqi;
};

template <int ND>
__device__ __forceinline__ int uci_col_order(uint32_t col, bool ri) // position of `col` in the field's filling order, -1 = not its column
{
  if (ND > 10) {
    if (ri) return col == 1 ? 0 : col == 10 ? 1 : col == 7 ? 2 : col == 4 ? 3 : -1;
    return col == 2 ? 0 : col == 9 ? 1 : col == 8 ? 2 : col == 3 ? 3 : -1;
  }
  if (ri) return col == 0 ? 0 : col == 8 ? 1 : col == 5 ? 2 : col == 3 ? 3 : -1;
  return col == 1 ? 0 : col == 7 ? 1 : col == 6 ? 2 : col == 2 ? 3 : -1;
}

template <int MOD, int ND, bool UCI = false> // srsran_mod_t 1..3; data symbols per subframe 12 (normal CP) or 10 (extended): constants, so no run-time divisions
__global__ void __launch_bounds__(256) pusch_demod_descramble_kernel(const float2* __restrict__ d, const uint32_t* __restrict__ seq,
                                                                     int16_t* __restrict__ g, uint32_t M, uint32_t nwords, int shift,
                                                                     float qpsk_scale, PuschUciArgs uci)
{
  constexpr int      mod = MOD;
  constexpr uint32_t nd  = ND;
  __shared__ float2 tile[12][DEMOD_TJ];
  const uint32_t sf = blockIdx.y, j0 = blockIdx.x * DEMOD_TJ;
  const uint32_t per_sf = nd * M;
  const float2*  src    = d + (size_t)sf * per_sf;
  for (uint32_t idx = threadIdx.x; idx < nd * DEMOD_TJ; idx += blockDim.x) {
    const uint32_t i = idx / DEMOD_TJ, jl = idx % DEMOD_TJ;
    if (j0 + jl < M) tile[i][jl] = __ldcs(&src[(size_t)i * M + j0 + jl]);
  }
  __syncthreads();
  const uint32_t  body = 4u * (per_sf / 4u), fbody = 16u * (2u * per_sf / 16u);
  constexpr int   Qm   = 2 * MOD;
  const uint32_t* sq   = seq + (size_t)sf * nwords;
  uint32_t*       dst  = reinterpret_cast<uint32_t*>(g + (size_t)sf * per_sf * Qm);
  for (uint32_t idx = threadIdx.x; idx < nd * DEMOD_TJ; idx += blockDim.x) {
    const uint32_t jl = idx / nd, i = idx % nd, j = j0 + jl;
    if (j >= M) break;
    const uint32_t pos = i * M + j;
    int16_t        o[6];
    demod_one(mod, tile[i][jl], pos < body, 2u * pos, fbody, qpsk_scale, o);
    const uint32_t n = pos * (uint32_t)Qm, wd = n >> 5, sh = n & 31u;
    uint32_t       bits = sq[wd] >> sh;
    if (sh + (uint32_t)Qm > 32u) bits |= sq[wd + 1] << (32u - sh);
    if (UCI) {
      const PuschUciSf u  = uci.sf[sf];
      const uint32_t   jr = M - 1u - j; // rows counted from the bottom, where both fields start
      const int        o_ri = uci_col_order<ND>(i, true), o_ack = uci_col_order<ND>(i, false);
      const bool       is_ri = o_ri >= 0 && 4u * jr + (uint32_t)o_ri < u.Q_ri, is_ack = o_ack >= 0 && 4u * jr + (uint32_t)o_ack < u.Q_ack;
      // RI symbols in the rows above this one, plus those of this row in earlier columns
      uint32_t before = u.Q_ri > 4u * (jr + 1u) ? u.Q_ri - 4u * (jr + 1u) : 0u;
#pragma unroll
      for (uint32_t c = 0; c < nd; c++) {
        const int o = uci_col_order<ND>(c, true);
        if (o >= 0 && c < i && 4u * jr + (uint32_t)o < u.Q_ri) before++;
      }
      const uint32_t sym  = j * nd + i - before; // index of this symbol in the de-interleaved stream (unless it is RI)
      int16_t*       row  = uci.llr + (size_t)sf * uci.stride;
      int16_t*       gs   = g + (size_t)sf * per_sf * Qm;
      const bool     last_ri = is_ri && jr == 0u && o_ri == (u.Q_ri >= 2u ? 1 : 0); // the largest RI position of the matrix
#pragma unroll
      for (int k = 0; k < Qm; k++) {
        int v = (int)o[k];
        if ((bits >> k) & 1u) v = -v;
        const int16_t raw = (int16_t)v;                       // reference scale, int16 wrap like srsran_sequence_pusch_apply_s
        int           vs  = (int)o[k] >> shift;
        if ((bits >> k) & 1u) vs = -vs;
        const int16_t sh = (int16_t)vs;
        if (is_ri) {
          int16_t f = raw;
          if (k == 1 && (u.flags & 2u) && (((bits >> 1) ^ bits) & 1u)) f = (int16_t)(-(int)raw);
          row[uci.off_ri + (4u * jr + (uint32_t)o_ri) * Qm + k] = f;
          if (last_ri && k == Qm - 1) {
            gs[0] = sh;
            if (u.Q_cqi) row[uci.off_cqi] = raw;
          }
        } else {
          if (is_ack) {
            int16_t f = raw;
            if (k == 1 && (u.flags & 1u) && (((bits >> 1) ^ bits) & 1u)) f = (int16_t)(-(int)raw);
            row[(4u * jr + (uint32_t)o_ack) * Qm + k] = f;
          }
          if (!(sym == 0u && k == 0 && u.Q_ri)) {
            gs[(size_t)sym * Qm + k] = is_ack ? (int16_t)0 : sh;
            if (sym < u.Q_cqi) row[uci.off_cqi + sym * Qm + k] = is_ack ? (int16_t)0 : raw;
          }
        }
      }
      continue;
    }
    const size_t ow = ((size_t)j * nd + i) * (size_t)mod;
#pragma unroll
    for (int w = 0; w < 3; w++) {
      if (w < mod) {
        int lo = (int)o[2 * w] >> shift, hi = (int)o[2 * w + 1] >> shift;
        if ((bits >> (2 * w)) & 1u) lo = -lo;
        if ((bits >> (2 * w + 1)) & 1u) hi = -hi;
        dst[ow + w] = (uint32_t)(uint16_t)(int16_t)lo | ((uint32_t)(uint16_t)(int16_t)hi << 16);
      }
    }
  }
}

// The same stage without control information, vectorised: a thread owns G consecutive data symbols of one subcarrier (4 with the
// normal prefix, 2 with the extended one), i.e. G*Qm consecutive soft bits of the de-interleaved stream, and writes them as
// 16-byte (8-byte) vectors; consecutive threads write consecutive chunks.  The tile's scrambling bits are staged in shared memory
// once per row instead of one scattered word load per symbol.  d and g must be 16-byte aligned (the host picks the kernel).
template <int MOD, int ND>
__global__ void __launch_bounds__(192) pusch_demod_descramble_vec_kernel(const float2* __restrict__ d, const uint32_t* __restrict__ seq,
                                                                         int16_t* __restrict__ g, uint32_t M, uint32_t nwords, int shift,
                                                                         float qpsk_scale)
{
  constexpr int Qm = 2 * MOD, G = (ND % 4 == 0) ? 4 : 2, NG = ND / G, TJ = DEMOD_TJ, TP = TJ + 2;
  constexpr int SW = (TJ * Qm + 31) / 32 + 2; // a row's scrambling bits: TJ*Qm of them, not word aligned, plus the funnel shift's upper word
  __shared__ __align__(16) float2 tile[ND][TP];
  __shared__ uint32_t sbits[ND][SW];
  const uint32_t  sf = blockIdx.y, j0 = blockIdx.x * TJ;
  const uint32_t  per_sf = (uint32_t)ND * M;
  const float2*   src = d + (size_t)sf * per_sf;
  const uint32_t* sq  = seq + (size_t)sf * nwords;
  for (uint32_t idx = threadIdx.x; idx < (uint32_t)(ND * (TJ / 2)); idx += blockDim.x) {
    const uint32_t i = idx / (TJ / 2), jl = 2u * (idx % (TJ / 2));
    if (j0 + jl < M) *reinterpret_cast<float4*>(&tile[i][jl]) = __ldcs(reinterpret_cast<const float4*>(&src[(size_t)i * M + j0 + jl]));
  }
  for (uint32_t idx = threadIdx.x; idx < (uint32_t)(ND * SW); idx += blockDim.x) {
    const uint32_t i = idx / SW, t = idx % SW, w = (((i * M + j0) * (uint32_t)Qm) >> 5) + t;
    sbits[i][t] = w < nwords ? sq[w] : 0u;
  }
  __syncthreads();
  const uint32_t body = 4u * (per_sf / 4u), fbody = 16u * (2u * per_sf / 16u);
  uint32_t*      dst  = reinterpret_cast<uint32_t*>(g + (size_t)sf * per_sf * Qm);
  for (uint32_t item = threadIdx.x; item < (uint32_t)(TJ * NG); item += blockDim.x) {
    const uint32_t jl = item / NG, iq = item % NG, j = j0 + jl;
    if (j >= M) break;
    uint32_t words[G * MOD];
#pragma unroll
    for (int r = 0; r < G; r++) {
      const uint32_t i = iq * G + r, pos = i * M + j;
      int16_t        o[6];
      demod_one(MOD, tile[i][jl], pos < body, 2u * pos, fbody, qpsk_scale, o);
      const uint32_t rel  = (((i * M + j0) * (uint32_t)Qm) & 31u) + jl * (uint32_t)Qm;
      const uint32_t bits = __funnelshift_r(sbits[i][rel >> 5], sbits[i][(rel >> 5) + 1], rel & 31u);
#pragma unroll
      for (int w = 0; w < MOD; w++) {
        int lo = (int)o[2 * w] >> shift, hi = (int)o[2 * w + 1] >> shift;
        if ((bits >> (2 * w)) & 1u) lo = -lo;
        if ((bits >> (2 * w + 1)) & 1u) hi = -hi;
        words[r * MOD + w] = (uint32_t)(uint16_t)(int16_t)lo | ((uint32_t)(uint16_t)(int16_t)hi << 16);
      }
    }
    if (G == 4) {
      uint4* o4 = reinterpret_cast<uint4*>(dst) + ((size_t)j * (ND / 4) + iq) * MOD;
#pragma unroll
      for (int q = 0; q < MOD; q++) o4[q] = make_uint4(words[4 * q], words[4 * q + 1], words[4 * q + 2], words[4 * q + 3]);
    } else {
      uint2* o2 = reinterpret_cast<uint2*>(dst) + ((size_t)j * (ND / 2) + iq) * MOD;
#pragma unroll
      for (int q = 0; q < MOD; q++) o2[q] = make_uint2(words[2 * q], words[2 * q + 1]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side

// TS 36.211 7.2 generator, bit-serial (init-time tables only)
struct Gold31 {
  uint32_t x1 = 1, x2 = 0;
  explicit Gold31(uint32_t c_init) : x2(c_init) {}
  static uint32_t step1(uint32_t x) { return (x >> 1) | ((((x >> 3) ^ x) & 1u) << 30); }
  static uint32_t step2(uint32_t x) { return (x >> 1) | ((((x >> 3) ^ (x >> 2) ^ (x >> 1) ^ x) & 1u) << 30); }
  void            bits(uint8_t* c, uint32_t len)
  {
    for (uint32_t n = 0; n < 1600 + len; n++) {
      if (n >= 1600) c[n - 1600] = (uint8_t)((x1 ^ x2) & 1u);
      x1 = step1(x1);
      x2 = step2(x2);
    }
  }
};

static uint32_t prime_lower_than(uint32_t n)
{
  for (uint32_t p = n - 1; p >= 2; p--) {
    bool ok = true;
    for (uint32_t d = 2; d * d <= p; d++) {
      if (p % d == 0) {
        ok = false;
        break;
      }
    }
    if (ok) return p;
  }
  return 0;
}

struct PuschRx {
  DeviceContext*          ctx = nullptr;
  srsran_b200_pusch_cfg_t cfg{};
  int                     nsym = 14, nd = 12, M = 0, R = 0, Qm = 0;
  uint32_t                nwords = 0, wpl = 0;
  OfdmPlanDev             plan{};
  int                     sm_count = 148;
  float2*                 dW      = nullptr;
  float2*                 d_dmrs  = nullptr; // [8][10][2][M]
  uint32_t*               d_x1w   = nullptr;
  uint32_t*               d_jump  = nullptr;
  std::vector<float2>     h_dmrs;
  // per-call scratch, grow-only
  PuschSfParam* d_prm = nullptr;
  uint32_t*     d_seq = nullptr;
  float2*       d_ce  = nullptr;
  float2*       d_d   = nullptr;
  float*        d_meas = nullptr;
  uint32_t      cap_sf = 0;
  // per-subframe parameters are staged in a ring of page-locked buffers (a copy from pageable memory would first
  // synchronise the stream, i.e. wait for the OFDM kernel queued just before this call).  The ring is deep because a caller
  // that pipelines a batch chunk by chunk queues all its chunks before the first one's samples have arrived: with two
  // buffers the third call waited for the first chunk's copy and the host fell into step with PCIe (enb_ul.cu).
  static constexpr int PRM_RING = 16;
  PuschSfParam* h_prm[PRM_RING]  = {};
  cudaEvent_t   prm_ev[PRM_RING] = {};
  uint32_t      h_prm_cap  = 0;
  int           prm_cur    = 0;
  // control information: one buffer set per rx_uci_batch call that has not been collected yet (srsran_b200_pusch_uci_collect)
  struct UciBatch {
    PuschUciSf* d_sf = nullptr; // device and page-locked host copies of the per-subframe field sizes
    PuschUciSf* h_sf = nullptr;
    int16_t*    d_llr = nullptr; // the fields' soft bits, [nsf][stride]
    int16_t*    h_llr = nullptr;
    size_t      cap_sf = 0, cap_llr = 0;
    cudaEvent_t done = nullptr;
    bool        pending = false;
    uint32_t    nsf = 0, stride = 0, off_ri = 0, off_cqi = 0;
    std::vector<UciGeometry>           geo;
    std::vector<srsran_b200_uci_cfg_t> cfg;
  };
  std::vector<UciBatch*> uci_pool;
  std::vector<UciBatch*> uci_pending; // in call order

  ~PuschRx()
  {
    if (ctx) cudaSetDevice(ctx->device);
    for (UciBatch* b : uci_pool) {
      if (b->done) {
        cudaEventSynchronize(b->done);
        cudaEventDestroy(b->done);
      }
      if (b->d_sf) cudaFree(b->d_sf);
      if (b->d_llr) cudaFree(b->d_llr);
      if (b->h_sf) cudaFreeHost(b->h_sf);
      if (b->h_llr) cudaFreeHost(b->h_llr);
      delete b;
    }
    for (void* p : {(void*)dW, (void*)d_dmrs, (void*)d_x1w, (void*)d_jump, (void*)d_prm, (void*)d_seq, (void*)d_ce, (void*)d_d, (void*)d_meas}) {
      if (p) cudaFree(p);
    }
    for (int i = 0; i < PRM_RING; i++) {
      if (h_prm[i]) cudaFreeHost(h_prm[i]);
      if (prm_ev[i]) cudaEventDestroy(prm_ev[i]);
    }
  }

  // refsignal_ul.c:337-357 for every (n_dmrs, sf_idx); float arithmetic of zc_sequence.c:205-300 as the x86 build of the
  // reference evaluates it: the ZC argument is a double expression rounded to float, the cyclic-shift term joins it in one
  // fused multiply-add, then a single-precision complex exponential.
  int gen_dmrs()
  {
    static const uint32_t n_dmrs_1[8] = {0, 2, 3, 4, 6, 8, 9, 10}, n_dmrs_2[8] = {0, 6, 3, 4, 2, 8, 10, 9}; // TS 36.211 5.5.2.1.1
    const uint32_t cell_id = cfg.cell_id, L = cfg.L_prb, dss = cfg.dmrs_delta_ss;
    const uint32_t nsl = (uint32_t)nsym / 2;
    uint8_t        c[8 * 7 * 20], cg[160];
    Gold31(((cell_id / 30) << 5) + (((cell_id % 30) + dss) % 30)).bits(c, 8 * nsl * 20); // refsignal_ul.c:95-118
    Gold31(cell_id / 30).bits(cg, 160);                                                  // phy_common.c:471-489
    const uint32_t Nzc = prime_lower_than((uint32_t)M);
    h_dmrs.assign((size_t)80 * 2 * M, make_float2(0.f, 0.f));
    for (uint32_t nd_ = 0; nd_ < 8; nd_++) {
      for (uint32_t sf_idx = 0; sf_idx < 10; sf_idx++) {
        float2* r = h_dmrs.data() + (size_t)(nd_ * 10 + sf_idx) * 2 * M;
        for (uint32_t ns = 2 * sf_idx; ns < 2 * sf_idx + 2; ns++) {
          uint32_t n_prs = 0, f_gh = 0;
          for (int i = 0; i < 8; i++) n_prs += (uint32_t)c[8 * nsl * ns + i] << i;
          if (cfg.group_hopping_en) {
            for (int i = 0; i < 8; i++) f_gh += (uint32_t)cg[8 * ns + i] << i;
          }
          const uint32_t n_cs  = (n_dmrs_1[cfg.dmrs_cyclic_shift] + n_dmrs_2[nd_] + n_prs) % 12;
          const float    alpha = (float)(2 * M_PI * (n_cs) / 12); // refsignal_ul.c:172-181
          const uint32_t u     = (f_gh + (cell_id % 30) + dss) % 30;
          uint32_t       v     = 0;
          if (L >= 6 && cfg.sequence_hopping_en) v = c[ns]; // refsignal_ul.c:121-132,242-245
          const float n_sz  = (float)Nzc;
          const float q_hat = n_sz * (float)(u + 1) / 31;
          float       qf;
          if ((((uint32_t)(2 * q_hat)) % 2) == 0) qf = (float)((double)q_hat + 0.5 + (double)v);
          else qf = (float)((double)q_hat + 0.5 - (double)v);
          const float q = (float)(uint32_t)qf;
          for (int i = 0; i < M; i++) {
            float arg;
            if (M <= 24) { // 1 and 2 PRB: phi(n) pi / 4 from the tables of TS 36.211 5.5.1.2 (zc_sequence.c:175-183, float product)
              const uint64_t w   = (M == 12 ? g_phi12 : g_phi24)[u];
              const float    phi = (float)(2 * (int)((w >> (2 * i)) & 3u) - 3);
              arg                = phi * (float)M_PI_4;
            } else {
              const float m = (float)((uint32_t)i % Nzc);
              arg           = (float)(-M_PI * (double)q * (double)m * (double)(m + 1) / (double)n_sz);
            }
            const float a   = fmaf(alpha, (float)i, arg);
            float       sn, cs;
            sincosf(a, &sn, &cs);
            r[(ns % 2) * M + i] = make_float2(cs, sn);
          }
        }
      }
    }
    return B200_SUCCESS;
  }

  int configure(const srsran_b200_pusch_cfg_t& c)
  {
    cfg  = c;
    nsym = c.cp_ext ? 12 : 14;
    nd   = nsym - 2 - (c.shortened ? 1 : 0); // ra_ul.c:232: the last symbol of a shortened subframe carries the SRS
    M    = 12 * (int)c.L_prb;
    R    = 12 * (int)c.cell_nof_prb;
    Qm   = 2 * c.modulation;
    int       radix[OFDM_MAX_PASSES];
    const int npass = fft_factorise(M, radix);
    if (c.modulation < 1 || c.modulation > 3 || c.L_prb < 1 || c.n_prb + c.L_prb > c.cell_nof_prb || c.cell_nof_prb > 110 || npass == 0 ||
        c.dmrs_cyclic_shift > 7 || c.dmrs_delta_ss > 29 || c.cell_id > 503 || c.llr_shift > 15) {
      // dft_precoding.c:88-104 accepts exactly the allocations whose 12 L_prb is 2^a 3^b 5^c
      B200_LOG_ERROR("unsupported PUSCH configuration (L_prb=%u n_prb=%u cell_nof_prb=%u mod=%d)", c.L_prb, c.n_prb, c.cell_nof_prb, c.modulation);
      return B200_ERROR_INVALID_INPUTS;
    }
    B200_CUDA_TRY(cudaSetDevice(ctx->device));
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, ctx->device);
    // transform de-precoding plan: backward DFT of 12 L_prb points scaled by 1/sqrt(N) (dft_precoding.c:46-50)
    plan         = OfdmPlanDev{};
    plan.N       = M;
    plan.R       = M;
    plan.nsym    = 1;
    plan.generic = 2;
    plan.inverse = 1;
    plan.gscale  = 1.0f / sqrtf((float)M);
    plan.npass   = npass;
    for (int i = 0; i < OFDM_MAX_PASSES; i++) plan.radix[i] = radix[i];
    int tps = M / 16;
    if (tps < 8) tps = 8;
    if (tps > OFDM_THREADS) tps = OFDM_THREADS;
    plan.tps       = tps;
    plan.pusch_nd  = nd;
    for (int l = 0, k = 0; l < nsym; l++) {
      if (l != nsym / 2 - 4 && l != nsym - 4 && !(c.shortened && l == nsym - 1)) plan.pusch_l[k++] = (unsigned char)l; // pusch.c:63-72
    }
    plan.grid_nsym = nsym;
    plan.grid_R    = R;
    plan.grid_off  = 12 * (int)c.n_prb;
    std::vector<float2> W(M);
    for (int m = 0; m < M; m++) {
      const double a = -2.0 * M_PI * (double)m / (double)M;
      W[m]           = make_float2((float)cos(a), (float)sin(a));
    }
    B200_CUDA_TRY(cudaMalloc(&dW, M * sizeof(float2)));
    B200_CUDA_TRY(cudaMemcpy(dW, W.data(), M * sizeof(float2), cudaMemcpyHostToDevice));
    plan.W = dW;
    // DMRS table
    if (gen_dmrs() != B200_SUCCESS) return B200_ERROR;
    B200_CUDA_TRY(cudaMalloc(&d_dmrs, h_dmrs.size() * sizeof(float2)));
    B200_CUDA_TRY(cudaMemcpy(d_dmrs, h_dmrs.data(), h_dmrs.size() * sizeof(float2), cudaMemcpyHostToDevice));
    // scrambling tables: x1 words and the jump states of the 31 seed bits at every lane's chunk start
    const uint32_t nbits = (uint32_t)(nd * M * Qm);
    nwords               = (nbits + 31) / 32 + 1; // one spare word: the kernel may look one word ahead
    wpl                  = (nwords + GOLD_CHUNKS - 1) / GOLD_CHUNKS;
    std::vector<uint32_t> x1w(nwords, 0u), jump(GOLD_CHUNKS * 31, 0u);
    {
      uint32_t x1 = 1;
      for (uint32_t n = 0; n < 1600; n++) x1 = Gold31::step1(x1);
      for (uint32_t w = 0; w < nwords; w++) {
        uint32_t v = 0;
        for (int b = 0; b < 32; b++) {
          v |= (x1 & 1u) << b;
          x1 = Gold31::step1(x1);
        }
        x1w[w] = v;
      }
      for (int b = 0; b < 31; b++) {
        uint32_t x2 = 1u << b;
        for (uint32_t n = 0; n < 1600; n++) x2 = Gold31::step2(x2);
        for (uint32_t lane = 0; lane < GOLD_CHUNKS; lane++) {
          jump[lane * 31 + b] = x2;
          for (uint32_t n = 0; n < 32 * wpl; n++) x2 = Gold31::step2(x2);
        }
      }
    }
    B200_CUDA_TRY(cudaMalloc(&d_x1w, nwords * sizeof(uint32_t)));
    B200_CUDA_TRY(cudaMemcpy(d_x1w, x1w.data(), nwords * sizeof(uint32_t), cudaMemcpyHostToDevice));
    B200_CUDA_TRY(cudaMalloc(&d_jump, jump.size() * sizeof(uint32_t)));
    B200_CUDA_TRY(cudaMemcpy(d_jump, jump.data(), jump.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    return B200_SUCCESS;
  }

  int reserve(uint32_t nsf)
  {
    if (nsf <= cap_sf) return B200_SUCCESS;
    for (void* p : {(void*)d_prm, (void*)d_seq, (void*)d_ce, (void*)d_d, (void*)d_meas}) {
      if (p) cudaFree(p);
    }
    d_prm = nullptr; d_seq = nullptr; d_ce = nullptr; d_d = nullptr; d_meas = nullptr;
    cap_sf = 0;
    B200_CUDA_TRY(cudaMalloc(&d_prm, (size_t)nsf * sizeof(PuschSfParam)));
    B200_CUDA_TRY(cudaMalloc(&d_seq, (size_t)nsf * nwords * sizeof(uint32_t)));
    B200_CUDA_TRY(cudaMalloc(&d_ce, (size_t)nsf * 2 * M * sizeof(float2)));
    B200_CUDA_TRY(cudaMalloc(&d_d, (size_t)nsf * nd * M * sizeof(float2)));
    B200_CUDA_TRY(cudaMalloc(&d_meas, (size_t)nsf * 4 * sizeof(float)));
    cap_sf = nsf;
    return B200_SUCCESS;
  }

  // per-subframe parameters -> device (the copy is stream ordered; the host vector is staged by the runtime)
  int upload_params(uint32_t nsf, const uint32_t* rnti, const uint32_t* tti, const uint32_t* n_dmrs, cudaStream_t st)
  {
    if (nsf > h_prm_cap) {
      for (int i = 0; i < PRM_RING; i++) {
        if (prm_ev[i]) B200_CUDA_TRY(cudaEventSynchronize(prm_ev[i]));
        if (h_prm[i]) cudaFreeHost(h_prm[i]);
        h_prm[i] = nullptr;
        B200_CUDA_TRY(cudaHostAlloc((void**)&h_prm[i], (size_t)nsf * sizeof(PuschSfParam), cudaHostAllocDefault));
        if (!prm_ev[i]) B200_CUDA_TRY(cudaEventCreateWithFlags(&prm_ev[i], cudaEventDisableTiming));
      }
      h_prm_cap = nsf;
    }
    prm_cur = (prm_cur + 1) % PRM_RING;
    B200_CUDA_TRY(cudaEventSynchronize(prm_ev[prm_cur])); // the copy that last read this buffer has completed
    PuschSfParam* hp = h_prm[prm_cur];
    for (uint32_t i = 0; i < nsf; i++) {
      const uint32_t sf_idx = tti ? tti[i] % 10 : 0, nd_ = n_dmrs ? n_dmrs[i] : 0, r = rnti ? rnti[i] : 0;
      if (nd_ > 7) {
        B200_LOG_ERROR("n_dmrs %u out of range (refsignal_ul.c:327)", nd_);
        return B200_ERROR_INVALID_INPUTS;
      }
      hp[i].c_init   = ((r & 0xFFFFu) << 14) + (sf_idx << 9) + cfg.cell_id;
      hp[i].dmrs_idx = nd_ * 10 + sf_idx;
    }
    const uint32_t nw = nsf * (uint32_t)(sizeof(PuschSfParam) / 4);
    pusch_words_to_device_kernel<<<(nw + 255) / 256, 256, 0, st>>>((const uint32_t*)hp, (uint32_t*)d_prm, nw);
    B200_CUDA_TRY(cudaGetLastError());
    B200_CUDA_TRY(cudaEventRecord(prm_ev[prm_cur], st));
    return B200_SUCCESS;
  }

  int chest(const float2* grid, float2* ce, float* meas, uint32_t nsf, cudaStream_t st)
  {
    const float  w   = 0.3333f;                                                         // chest_ul.c:82-83
    const float  cal = (float)((7.419 * w * w + 0.1117 * w - 0.005387) * 0.8);           // chest_ul.c:216-219
    const size_t sm  = (size_t)2 * M * sizeof(float2);
    if (sm > 48 * 1024) { // per device and cheap: set before every launch that needs the opt-in
      B200_CUDA_TRY(cudaFuncSetAttribute(pusch_chest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    }
    pusch_chest_kernel<<<nsf, 256, sm, st>>>(grid, d_dmrs, d_prm, ce, meas, nsym, R, plan.grid_off, M, w, cal);
    g_kernel_launches++;
    B200_CUDA_TRY(cudaGetLastError());
    return B200_SUCCESS;
  }

  int equalize_deprecode(const float2* grid, const float2* ce, const float* meas, float2* d, uint32_t nsf, cudaStream_t st)
  {
    OfdmPlanDev p     = plan;
    p.eq_ce           = ce;
    p.eq_noise        = meas;
    p.eq_noise_stride = 4;
    if (launch_ofdm_rx(p, grid, d, nsf * (uint32_t)nd, sm_count, st) != B200_SUCCESS) return B200_ERROR;
    g_kernel_launches++;
    return B200_SUCCESS;
  }

  // field sizes and the UL-SCH span of one grant (uci.c:172-190,395-418; sch.c:1136,1186-1190)
  int uci_span(uint32_t tbs, const srsran_b200_uci_cfg_t& c, UciGeometry* geo, uint32_t* e_offset, uint32_t* nof_e_bits) const
  {
    CbSegm seg;
    if (cb_segmentation(tbs, seg) != 0) return B200_ERROR_INVALID_INPUTS;
    const int rc = uci_geometry(c, seg.C1 * seg.K1 + seg.C2 * seg.K2, (uint32_t)M, (uint32_t)nd, geo);
    if (rc != B200_SUCCESS) return rc;
    // (Q'_cqi is capped at what the RI symbols leave, uci.c:186: the transport block may be left with nothing -- the reference then
    // de-matches zero soft bits and the block fails its CRC; the same happens here)
    if (e_offset) *e_offset = geo->Q_cqi * (uint32_t)Qm;
    if (nof_e_bits) *nof_e_bits = ((uint32_t)(nd * M) - geo->Q_ri - geo->Q_cqi) * (uint32_t)Qm;
    return B200_SUCCESS;
  }

  // buffers of one rx_uci_batch call: field sizes to the device now, soft bits back after the kernel
  int uci_begin(uint32_t nsf, const uint32_t* tbs, const srsran_b200_uci_cfg_t* uci, cudaStream_t st, UciBatch** out)
  {
    UciBatch* b = nullptr;
    for (UciBatch* c : uci_pool) {
      if (!c->pending) {
        b = c;
        break;
      }
    }
    if (!b) {
      b = new (std::nothrow) UciBatch();
      if (!b) return B200_ERROR;
      uci_pool.push_back(b);
      B200_CUDA_TRY(cudaEventCreateWithFlags(&b->done, cudaEventDisableTiming));
    }
    b->geo.assign(nsf, UciGeometry());
    b->cfg.assign(uci, uci + nsf);
    uint32_t qa = 0, qr = 0, qc = 0;
    for (uint32_t i = 0; i < nsf; i++) {
      const int rc = uci_span(tbs[i], uci[i], &b->geo[i], nullptr, nullptr);
      if (rc != B200_SUCCESS) return rc;
      qa = b->geo[i].Q_ack > qa ? b->geo[i].Q_ack : qa;
      qr = b->geo[i].Q_ri > qr ? b->geo[i].Q_ri : qr;
      qc = b->geo[i].Q_cqi > qc ? b->geo[i].Q_cqi : qc;
    }
    b->nsf     = nsf;
    b->off_ri  = qa * (uint32_t)Qm;
    b->off_cqi = (qa + qr) * (uint32_t)Qm;
    b->stride  = ((qa + qr + qc) * (uint32_t)Qm + 7u) & ~7u;
    if (b->stride == 0) b->stride = 8;
    if (nsf > b->cap_sf) {
      if (b->d_sf) cudaFree(b->d_sf);
      if (b->h_sf) cudaFreeHost(b->h_sf);
      b->d_sf = nullptr; b->h_sf = nullptr; b->cap_sf = 0;
      B200_CUDA_TRY(cudaMalloc(&b->d_sf, (size_t)nsf * sizeof(PuschUciSf)));
      B200_CUDA_TRY(cudaHostAlloc((void**)&b->h_sf, (size_t)nsf * sizeof(PuschUciSf), cudaHostAllocDefault));
      b->cap_sf = nsf;
    }
    const size_t need = (size_t)nsf * b->stride;
    if (need > b->cap_llr) {
      if (b->d_llr) cudaFree(b->d_llr);
      if (b->h_llr) cudaFreeHost(b->h_llr);
      b->d_llr = nullptr; b->h_llr = nullptr; b->cap_llr = 0;
      B200_CUDA_TRY(cudaMalloc(&b->d_llr, need * sizeof(int16_t)));
      B200_CUDA_TRY(cudaHostAlloc((void**)&b->h_llr, need * sizeof(int16_t), cudaHostAllocDefault));
      b->cap_llr = need;
    }
    for (uint32_t i = 0; i < nsf; i++) {
      b->h_sf[i] = PuschUciSf{b->geo[i].Q_ack, b->geo[i].Q_ri, b->geo[i].Q_cqi, (b->geo[i].ack_one_bit ? 1u : 0u) | (b->geo[i].ri_one_bit ? 2u : 0u)};
    }
    const uint32_t nw = nsf * (uint32_t)(sizeof(PuschUciSf) / 4);
    pusch_words_to_device_kernel<<<(nw + 255) / 256, 256, 0, st>>>((const uint32_t*)b->h_sf, (uint32_t*)b->d_sf, nw);
    B200_CUDA_TRY(cudaGetLastError());
    *out = b;
    return B200_SUCCESS;
  }

  int uci_end(UciBatch* b, cudaStream_t st)
  {
    B200_CUDA_TRY(cudaMemcpyAsync(b->h_llr, b->d_llr, (size_t)b->nsf * b->stride * sizeof(int16_t), cudaMemcpyDeviceToHost, st));
    B200_CUDA_TRY(cudaEventRecord(b->done, st));
    b->pending = true;
    uci_pending.push_back(b);
    return B200_SUCCESS;
  }

  int uci_collect(srsran_b200_uci_value_t* out, uint32_t nof_out)
  {
    uint32_t total = 0;
    for (UciBatch* b : uci_pending) total += b->nsf;
    if (!out) { // discard: the caller gave up on the batch
      for (UciBatch* b : uci_pending) {
        cudaEventSynchronize(b->done);
        b->pending = false;
      }
      uci_pending.clear();
      return B200_SUCCESS;
    }
    if (total != nof_out) {
      B200_LOG_ERROR("uci_collect: %u subframes pending, room for %u", total, nof_out);
      return B200_ERROR_INVALID_INPUTS;
    }
    uint32_t o = 0;
    int      rc = B200_SUCCESS;
    for (UciBatch* b : uci_pending) {
      if (cudaEventSynchronize(b->done) != cudaSuccess) rc = B200_ERROR;
      for (uint32_t i = 0; i < b->nsf && rc == B200_SUCCESS; i++, o++) {
        const int16_t*     row = b->h_llr + (size_t)i * b->stride;
        const UciGeometry& g   = b->geo[i];
        uci_decide(b->cfg[i], g, (uint32_t)Qm, row, row + b->off_ri, row + b->off_cqi, &out[o]);
        out[o].Q_prime_ack = g.Q_ack;
        out[o].Q_prime_ri  = g.Q_ri;
        out[o].Q_prime_cqi = g.Q_cqi;
        out[o].e_offset    = g.Q_cqi * (uint32_t)Qm;
        out[o].nof_e_bits  = ((uint32_t)(nd * M) - g.Q_ri - g.Q_cqi) * (uint32_t)Qm;
      }
      b->pending = false;
    }
    uci_pending.clear();
    return rc;
  }

  int demod(const float2* d, int16_t* g, uint32_t nsf, cudaStream_t st, const UciBatch* ub = nullptr)
  {
    pusch_gold_kernel<<<(nsf + 1) / 2, 2 * GOLD_CHUNKS, (size_t)2 * nwords * sizeof(uint32_t), st>>>(d_prm, d_x1w, d_jump, d_seq, nsf, nwords, wpl);
    g_kernel_launches++;
    B200_CUDA_TRY(cudaGetLastError());
    dim3 grid((unsigned)((M + DEMOD_TJ - 1) / DEMOD_TJ), nsf);
    const float qs = (float)(-100.0 * M_SQRT2);
    const int   sh = (int)cfg.llr_shift;
    PuschUciArgs ua{};
    if (ub) ua = PuschUciArgs{ub->d_sf, ub->d_llr, ub->stride, ub->off_ri, ub->off_cqi};
#define B200_DEMOD(MOD, ND)                                                                                                            \
  do {                                                                                                                                 \
    if (ub) pusch_demod_descramble_kernel<MOD, ND, true><<<grid, 256, 0, st>>>(d, d_seq, g, (uint32_t)M, nwords, sh, qs, ua);          \
    else pusch_demod_descramble_kernel<MOD, ND, false><<<grid, 256, 0, st>>>(d, d_seq, g, (uint32_t)M, nwords, sh, qs, ua);            \
  } while (0)
    const bool vec = !ub && ((((uintptr_t)d | (uintptr_t)g) & 15u) == 0) && !getenv("SRSLTE_B200_PUSCH_DEMOD_SCALAR");
#define B200_DEMODV(MOD, ND) pusch_demod_descramble_vec_kernel<MOD, ND><<<grid, 192, 0, st>>>(d, d_seq, g, (uint32_t)M, nwords, sh, qs)
#define B200_DEMOD_MOD(KIND, ND)                                                                                                       \
  do {                                                                                                                                 \
    if (cfg.modulation == 1) KIND(1, ND);                                                                                              \
    else if (cfg.modulation == 2) KIND(2, ND);                                                                                         \
    else KIND(3, ND);                                                                                                                  \
  } while (0)
    if (vec && nd == 12) B200_DEMOD_MOD(B200_DEMODV, 12);
    else if (vec && nd == 10) B200_DEMOD_MOD(B200_DEMODV, 10);
    else if (nd == 12) B200_DEMOD_MOD(B200_DEMOD, 12);
    else if (nd == 11) B200_DEMOD_MOD(B200_DEMOD, 11); // shortened subframes: one symbol per thread (11 and 9 do not split into vectors)
    else if (nd == 10) B200_DEMOD_MOD(B200_DEMOD, 10);
    else B200_DEMOD_MOD(B200_DEMOD, 9);
#undef B200_DEMOD_MOD
#undef B200_DEMOD
#undef B200_DEMODV
    g_kernel_launches++;
    B200_CUDA_TRY(cudaGetLastError());
    return B200_SUCCESS;
  }
};

} // namespace b200

using namespace b200;

struct srsran_b200_pusch {
  PuschRx rx;
};

extern "C" SRSRAN_B200_API int srsran_b200_pusch_init(srsran_b200_pusch_t** q, int device, const srsran_b200_pusch_cfg_t* cfg)
{
  if (!q || !cfg) return B200_ERROR_INVALID_INPUTS;
  *q                 = nullptr;
  DeviceContext* ctx = device_context(device);
  if (!ctx) return B200_ERROR;
  srsran_b200_pusch_t* h = new (std::nothrow) srsran_b200_pusch_t();
  if (!h) return B200_ERROR;
  h->rx.ctx = ctx;
  const int rc = h->rx.configure(*cfg);
  if (rc != B200_SUCCESS) {
    delete h;
    return rc;
  }
  *q = h;
  return B200_SUCCESS;
}

extern "C" SRSRAN_B200_API void srsran_b200_pusch_free(srsran_b200_pusch_t* q)
{
  delete q;
}

extern "C" SRSRAN_B200_API int srsran_b200_pusch_geometry(const srsran_b200_pusch_t* q, uint32_t* nof_re, uint32_t* nof_bits,
                                                         uint32_t* nof_data_symbols)
{
  if (!q) return B200_ERROR_INVALID_INPUTS;
  if (nof_re) *nof_re = (uint32_t)(q->rx.nd * q->rx.M);
  if (nof_bits) *nof_bits = (uint32_t)(q->rx.nd * q->rx.M * q->rx.Qm);
  if (nof_data_symbols) *nof_data_symbols = (uint32_t)q->rx.nd;
  return B200_SUCCESS;
}

extern "C" SRSRAN_B200_API int srsran_b200_refsignal_dmrs_pusch_gen(const srsran_b200_pusch_t* q, uint32_t sf_idx, uint32_t n_dmrs, void* r)
{
  if (!q || !r || sf_idx > 9 || n_dmrs > 7) return B200_ERROR_INVALID_INPUTS;
  const size_t n = (size_t)2 * q->rx.M;
  memcpy(r, q->rx.h_dmrs.data() + (size_t)(n_dmrs * 10 + sf_idx) * n, n * sizeof(float2));
  return B200_SUCCESS;
}

static int need_device(uint32_t flags, const char* who)
{
  if (!(flags & SRSRAN_B200_FLAG_DEVICE_PTRS)) {
    B200_LOG_ERROR("%s works on device buffers (it sits between two device-side stages)", who);
    return B200_ERROR_INVALID_INPUTS;
  }
  return B200_SUCCESS;
}

extern "C" SRSRAN_B200_API int srsran_b200_chest_ul_pusch_batch(srsran_b200_pusch_t* q, const void* grid, uint32_t nsf, const uint32_t* tti,
                                                               const uint32_t* n_dmrs, void* ce, float* meas, uint32_t flags, void* stream)
{
  if (!q || !grid || !ce || !meas) return B200_ERROR_INVALID_INPUTS;
  if (need_device(flags, "srsran_b200_chest_ul_pusch_batch")) return B200_ERROR_INVALID_INPUTS;
  if (nsf == 0) return B200_SUCCESS;
  PuschRx& rx = q->rx;
  B200_CUDA_TRY(cudaSetDevice(rx.ctx->device));
  if (rx.reserve(nsf) != B200_SUCCESS) return B200_ERROR;
  int rc = rx.upload_params(nsf, nullptr, tti, n_dmrs, (cudaStream_t)stream);
  if (rc != B200_SUCCESS) return rc;
  return rx.chest((const float2*)grid, (float2*)ce, meas, nsf, (cudaStream_t)stream);
}

extern "C" SRSRAN_B200_API int srsran_b200_pusch_equalize_deprecode_batch(srsran_b200_pusch_t* q, const void* grid, const void* ce,
                                                                         const float* meas, void* d, uint32_t nsf, uint32_t flags, void* stream)
{
  if (!q || !grid || !ce || !d) return B200_ERROR_INVALID_INPUTS;
  if (need_device(flags, "srsran_b200_pusch_equalize_deprecode_batch")) return B200_ERROR_INVALID_INPUTS;
  if (nsf == 0) return B200_SUCCESS;
  PuschRx& rx = q->rx;
  B200_CUDA_TRY(cudaSetDevice(rx.ctx->device));
  return rx.equalize_deprecode((const float2*)grid, (const float2*)ce, meas, (float2*)d, nsf, (cudaStream_t)stream);
}

extern "C" SRSRAN_B200_API int srsran_b200_pusch_demod_descramble_batch(srsran_b200_pusch_t* q, const void* d, int16_t* g, uint32_t nsf,
                                                                       const uint32_t* rnti, const uint32_t* tti, uint32_t flags, void* stream)
{
  if (!q || !d || !g) return B200_ERROR_INVALID_INPUTS;
  if (need_device(flags, "srsran_b200_pusch_demod_descramble_batch")) return B200_ERROR_INVALID_INPUTS;
  if (nsf == 0) return B200_SUCCESS;
  PuschRx& rx = q->rx;
  B200_CUDA_TRY(cudaSetDevice(rx.ctx->device));
  if (rx.reserve(nsf) != B200_SUCCESS) return B200_ERROR;
  int rc = rx.upload_params(nsf, rnti, tti, nullptr, (cudaStream_t)stream);
  if (rc != B200_SUCCESS) return rc;
  return rx.demod((const float2*)d, g, nsf, (cudaStream_t)stream);
}

extern "C" SRSRAN_B200_API int srsran_b200_pusch_rx_batch(srsran_b200_pusch_t* q, const void* grid, int16_t* g, float* meas, uint32_t nsf,
                                                         const uint32_t* rnti, const uint32_t* tti, const uint32_t* n_dmrs, uint32_t flags,
                                                         void* stream)
{
  if (!q || !grid || !g) return B200_ERROR_INVALID_INPUTS;
  if (need_device(flags, "srsran_b200_pusch_rx_batch")) return B200_ERROR_INVALID_INPUTS;
  if (nsf == 0) return B200_SUCCESS;
  PuschRx&     rx = q->rx;
  cudaStream_t st = (cudaStream_t)stream;
  B200_CUDA_TRY(cudaSetDevice(rx.ctx->device));
  if (rx.reserve(nsf) != B200_SUCCESS) return B200_ERROR;
  int rc = rx.upload_params(nsf, rnti, tti, n_dmrs, st);
  if (rc != B200_SUCCESS) return rc;
  float* m = meas ? meas : rx.d_meas;
  if ((rc = rx.chest((const float2*)grid, rx.d_ce, m, nsf, st)) != B200_SUCCESS) return rc;
  if ((rc = rx.equalize_deprecode((const float2*)grid, rx.d_ce, m, rx.d_d, nsf, st)) != B200_SUCCESS) return rc;
  return rx.demod(rx.d_d, g, nsf, st);
}

extern "C" SRSRAN_B200_API int srsran_b200_pusch_uci_geometry(const srsran_b200_pusch_t* q, uint32_t tbs, const srsran_b200_uci_cfg_t* uci,
                                                             srsran_b200_uci_value_t* out)
{
  if (!q || !uci || !out) return B200_ERROR_INVALID_INPUTS;
  UciGeometry g;
  const int   rc = q->rx.uci_span(tbs, *uci, &g, &out->e_offset, &out->nof_e_bits);
  if (rc != B200_SUCCESS) return rc;
  out->Q_prime_ack = g.Q_ack;
  out->Q_prime_ri  = g.Q_ri;
  out->Q_prime_cqi = g.Q_cqi;
  return B200_SUCCESS;
}

extern "C" SRSRAN_B200_API int srsran_b200_pusch_rx_uci_batch(srsran_b200_pusch_t* q, const void* grid, int16_t* g, float* meas, uint32_t nsf,
                                                             const uint32_t* rnti, const uint32_t* tti, const uint32_t* n_dmrs,
                                                             const uint32_t* tbs, const srsran_b200_uci_cfg_t* uci, uint32_t flags, void* stream)
{
  if (!q || !grid || !g || !tbs || !uci) return B200_ERROR_INVALID_INPUTS;
  if (need_device(flags, "srsran_b200_pusch_rx_uci_batch")) return B200_ERROR_INVALID_INPUTS;
  if (nsf == 0) return B200_SUCCESS;
  PuschRx&     rx = q->rx;
  cudaStream_t st = (cudaStream_t)stream;
  B200_CUDA_TRY(cudaSetDevice(rx.ctx->device));
  if (rx.reserve(nsf) != B200_SUCCESS) return B200_ERROR;
  PuschRx::UciBatch* ub = nullptr;
  int                rc = rx.uci_begin(nsf, tbs, uci, st, &ub); // validates every grant before anything is enqueued
  if (rc != B200_SUCCESS) return rc;
  if ((rc = rx.upload_params(nsf, rnti, tti, n_dmrs, st)) != B200_SUCCESS) return rc;
  float* m = meas ? meas : rx.d_meas;
  if ((rc = rx.chest((const float2*)grid, rx.d_ce, m, nsf, st)) != B200_SUCCESS) return rc;
  if ((rc = rx.equalize_deprecode((const float2*)grid, rx.d_ce, m, rx.d_d, nsf, st)) != B200_SUCCESS) return rc;
  if ((rc = rx.demod(rx.d_d, g, nsf, st, ub)) != B200_SUCCESS) return rc;
  return rx.uci_end(ub, st);
}

extern "C" SRSRAN_B200_API int srsran_b200_pusch_uci_collect(srsran_b200_pusch_t* q, srsran_b200_uci_value_t* out, uint32_t nof_out)
{
  if (!q) return B200_ERROR_INVALID_INPUTS;
  B200_CUDA_TRY(cudaSetDevice(q->rx.ctx->device));
  return q->rx.uci_collect(out, nof_out);
}
