// CUDA kernels of the batched turbo decoder (sm_100a).  The arithmetic lives in tdec_core.h; this file maps it onto
// the grid and adds the two layout-conversion kernels at the ends of a decode.
//
//   tdec_load_natural_kernel   natural [cb][3K+12] int16 (turbodecoder_gen.c:238-258 order)  ->  S/P0/P1/S2T tiles
//   tdec_siso_pass_kernel      one SISO pass (turbodecoder_iter.h:72-144) for every still-active code block
//   tdec_decide_kernel         HB (visiting order of the last pass)  ->  packed bytes, natural order
//                              (turbodecoder.c:370-378 + turbodecoder_gen.c:260-277)
#include <cuda_runtime.h>
#include <stdlib.h>

#include "tdec_core.h"
#include "tdec_kernels.h"

namespace b200 {

// ---------------------------------------------------------------------------------------------------------------
// SISO pass kernel.  One warp per tile of 64 code blocks, one block per warp: with the 65,536-block benchmark batch
// that is 1024 independent warps, ~7 per SM, each a long serial recursion.  Throughput therefore hangs on (a)
// instruction-level parallelism inside a trellis step and (b) keeping enough bytes in flight per warp.  A first
// version prefetched windows into registers with plain LDG; ncu showed 61 % long-scoreboard stalls at 2.7-3.9 TB/s
// because a warp has only six scoreboard slots, so a consumer of window w also waits for the loads of windows
// w-1..w-3 that alias its slot.  This version streams every window through a per-warp shared-memory ring filled by
// the TMA unit (cp.async.bulk -> UBLKCP, completion on an mbarrier): no scoreboard, prefetch depth = NSTAGE windows.
//
// Ring stage (4224 B):  S 1 KB | P 1 KB | E 1 KB | CK 1 KB | 8 CRC weights 64 B | 8 QPP entries 16 B
//   DEC1: S, P0, E are 1 KB contiguous runs of the tile (8 rows x 128 B or 2 uint4 rows x 512 B)
//   DEC2: P1 contiguous; E is eight 128-byte rows at PI(8w..8w+7), one bulk copy per row
namespace ring {
constexpr uint32_t OFF_S = 0, OFF_P = 1024, OFF_E = 2048, OFF_CK = 3072, OFF_CRC = 4096, OFF_QPP = 4160;
constexpr uint32_t STAGE_BYTES = 4224;
} // namespace ring

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy by the TMA unit; bytes multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src),
               "r"(bytes),
               "r"(bar)
               : "memory");
}

template <bool DEC2, bool FIRST, int NSTAGE>
struct WarpRing {
  uint32_t        base;  // shared-space address of stage 0
  uint32_t        bars;  // shared-space address of the NSTAGE mbarriers
  uint8_t*        gen;   // generic pointer to stage 0
  const TdecView& v;
  int             tile, lane;
  const uint8_t * gS, *gP, *gCK; // tile bases (bytes)
  const uint8_t*  gE;
  const uint8_t*  gCRC;
  uint32_t        issued = 0, consumed = 0;

  __device__ WarpRing(const TdecView& v_, uint8_t* smem, int tile_, int lane_) : v(v_), tile(tile_), lane(lane_)
  {
    gen  = smem;
    base = smem_u32(smem);
    bars = base + NSTAGE * ring::STAGE_BYTES;
    gS   = reinterpret_cast<const uint8_t*>(v.S + vec_row(v, tile, 0, 0));
    gP   = reinterpret_cast<const uint8_t*>((DEC2 ? v.P1 : v.P0) + vec_row(v, tile, 0, 0));
    gCK  = reinterpret_cast<const uint8_t*>(v.CK + ck_idx(v, tile, 0, 0, 0));
    gE   = reinterpret_cast<const uint8_t*>(v.E + e_idx(v, tile, 0, 0));
    gCRC = reinterpret_cast<const uint8_t*>(DEC2 ? v.crc_perm : v.crc_nat);
    if (lane == 0) {
      for (int s = 0; s < NSTAGE; s++) mbar_init(bars + 8u * s, 1);
    }
    if (FIRST) { // pass 0 has no a-priori: the E slot of every stage reads as zero and is never refilled
      for (int s = 0; s < NSTAGE; s++) {
        for (int t = 0; t < 8; t++) *reinterpret_cast<uint32_t*>(gen + s * ring::STAGE_BYTES + ring::OFF_E + t * 128 + lane * 4) = 0u;
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncwarp();
  }

  // Enqueue the copies of window w.  ALPHA adds the checkpoint, the CRC weights and (DEC2) the interleaver entries.
  // pi = PI(8w + lane - 8) for lanes 8..15 (DEC2 only), fetched by the caller one issue ahead.
  template <bool ALPHA>
  __device__ __forceinline__ void issue(uint32_t w, uint32_t pi)
  {
    const uint32_t s   = issued % NSTAGE;
    const uint32_t dst = base + s * ring::STAGE_BYTES;
    const uint32_t bar = bars + 8u * s;
    const bool     crc = ALPHA && gCRC != nullptr;
    uint32_t       tx  = 1024u;                       // P
    if (!DEC2) tx += 1024u;                           // S
    if (!FIRST) tx += 1024u;                          // E (contiguous or 8 rows)
    if (ALPHA) tx += 1024u + (DEC2 ? 16u : 0u);       // CK (+ QPP)
    if (crc) tx += 64u;
    if (lane == 0) mbar_expect_tx(bar, tx);
    __syncwarp();
    if (lane == 1) bulk_g2s(dst + ring::OFF_P, gP + (size_t)w * 1024u, 1024u, bar);
    if (!DEC2) {
      if (lane == 0) bulk_g2s(dst + ring::OFF_S, gS + (size_t)w * 1024u, 1024u, bar);
      if (!FIRST && lane == 2) bulk_g2s(dst + ring::OFF_E, gE + (size_t)w * 1024u, 1024u, bar);
    } else {
      if (lane >= 8 && lane < 16) bulk_g2s(dst + ring::OFF_E + (uint32_t)(lane - 8) * 128u, gE + (size_t)pi * 128u, 128u, bar);
    }
    if (ALPHA) {
      if (lane == 3) bulk_g2s(dst + ring::OFF_CK, gCK + (size_t)w * 1024u, 1024u, bar);
      if (crc && lane == 4) bulk_g2s(dst + ring::OFF_CRC, gCRC + (size_t)w * 64u, 64u, bar);
      if (DEC2 && lane == 5) bulk_g2s(dst + ring::OFF_QPP, v.qpp_fwd + 8u * w, 16u, bar);
    }
    issued++;
  }

  // Block until the oldest outstanding window has landed; returns its stage (generic pointer).
  __device__ __forceinline__ const uint8_t* acquire()
  {
    const uint32_t s = consumed % NSTAGE;
    mbar_wait(bars + 8u * s, (consumed / NSTAGE) & 1u);
    consumed++;
    return gen + s * ring::STAGE_BYTES;
  }

  __device__ __forceinline__ void read(WinIn<DEC2>& in, const uint8_t* st, bool alpha) const
  {
    in.p[0] = *reinterpret_cast<const u4*>(st + ring::OFF_P + lane * 16);
    in.p[1] = *reinterpret_cast<const u4*>(st + ring::OFF_P + 512 + lane * 16);
    if (!DEC2) {
      in.s[0] = *reinterpret_cast<const u4*>(st + ring::OFF_S + lane * 16);
      in.s[1] = *reinterpret_cast<const u4*>(st + ring::OFF_S + 512 + lane * 16);
    }
#pragma unroll
    for (int t = 0; t < 8; t++) in.e[t] = *reinterpret_cast<const uint32_t*>(st + ring::OFF_E + t * 128 + lane * 4);
    if (DEC2 && alpha) in.q = *reinterpret_cast<const u4*>(st + ring::OFF_QPP);
  }
};

template <bool DEC2, bool FIRST, int NSTAGE>
__global__ void __launch_bounds__(32) tdec_siso_pass_kernel(TdecView v, int pass_idx)
{
  extern __shared__ __align__(128) uint8_t smem[];
  const int tile = blockIdx.x;
  const int lane = threadIdx.x;
  if (tile >= v.ntiles) return;
  // whole tile finished (early stop): nothing to do for this warp
  CbStatus*  stp  = v.status + ((size_t)tile * TDEC_TILE_CB + 2 * (size_t)lane);
  CbStatus   s_lo = stp[0], s_hi = stp[1];
  const bool act_lo = s_lo.active != 0, act_hi = s_hi.active != 0;
  if (__ballot_sync(0xFFFFFFFFu, act_lo || act_hi) == 0u) return;

  const int      K  = v.K;
  const uint32_t nw = (uint32_t)K / 8u;
  WarpRing<DEC2, FIRST, NSTAGE> rg(v, smem, tile, lane);
  const LanePtrs p = lane_ptrs<DEC2>(v, tile, lane);
  const uint32_t qlane = (uint32_t)(lane & 7); // lanes 8..15 carry the row index of DEC2's E gather

  // ---------------- backward sweep ----------------
  uint32_t pi_next = 0;
  uint32_t next_w  = nw; // windows are issued nw-1, nw-2, ...
  auto     issue_beta = [&]() {
    next_w--;
    const uint32_t pi = pi_next;
    if (DEC2 && next_w > 0) pi_next = v.qpp_fwd[8u * (next_w - 1) + qlane];
    rg.template issue<false>(next_w, pi);
  };
  if (DEC2) pi_next = v.qpp_fwd[8u * (nw - 1) + qlane];
#pragma unroll 1
  for (int i = 0; i < NSTAGE && next_w > 0; i++) issue_beta();

  uint32_t B[8];
  B[0] = 0;
#pragma unroll
  for (int i = 1; i < 8; i++) B[i] = NEG_INF2;
  {
    u4 pt = p.P[(uint32_t)(K / 4) * 32u];
    u4 st = DEC2 ? v.S2T[(size_t)tile * 32 + lane] : p.S[(uint32_t)(K / 4) * 32u];
#pragma unroll
    for (int t = 2; t >= 0; t--) {
      uint32_t x = u4_get(st, t), y = u4_get(pt, t);
      beta_step(B, x, y, add2(x, y));
    }
  }
  store_ck(p, nw - 1, B);

#pragma unroll 1
  for (int w = (int)nw - 1; w >= 0; w--) {
    const uint8_t* st = rg.acquire();
    WinIn<DEC2>    cur;
    rg.read(cur, st, false);
#pragma unroll
    for (int t = 7; t >= 0; t--) {
      uint32_t x, y;
      win_xy<DEC2>(cur, t, x, y);
      beta_step(B, x, y, add2(x, y));
      if (t == 0 && w > 0) store_ck(p, (uint32_t)(w - 1), B);
      if ((t & 3) == 0) normalise(B);
    }
    __syncwarp(); // every lane has consumed the stage before the TMA unit overwrites it
    if (next_w > 0) issue_beta();
  }

  // ---------------- forward sweep ----------------
  // the checkpoints were written with ordinary stores and are now read back by the async proxy
  asm volatile("fence.proxy.async;" ::: "memory");
  __syncwarp();
  uint32_t up_w = 0;
  auto     issue_alpha = [&]() {
    const uint32_t pi = pi_next;
    if (DEC2 && up_w + 1 < nw) pi_next = v.qpp_fwd[8u * (up_w + 1) + qlane];
    rg.template issue<true>(up_w, pi);
    up_w++;
  };
  if (DEC2) pi_next = v.qpp_fwd[qlane];
#pragma unroll 1
  for (int i = 0; i < NSTAGE && up_w < nw; i++) issue_alpha();

  uint32_t   A[8];
  LaneResult res = {0u, 0u};
  A[0]           = 0;
#pragma unroll
  for (int i = 1; i < 8; i++) A[i] = NEG_INF2;
  const bool have_crc = v.crc_nat != nullptr;
#pragma unroll 1
  for (uint32_t w = 0; w < nw; w++) {
    const uint8_t* st = rg.acquire();
    WinIn<DEC2>    cur;
    rg.read(cur, st, true);
    const u4 c0 = *reinterpret_cast<const u4*>(st + ring::OFF_CK + lane * 16);
    const u4 c1 = *reinterpret_cast<const u4*>(st + ring::OFF_CK + 512 + lane * 16);
    alpha_window<DEC2>(v, p, w, cur, c0, c1, have_crc ? reinterpret_cast<const CrcPow*>(st + ring::OFF_CRC) : nullptr, A, res,
                       act_lo, act_hi);
    __syncwarp();
    if (up_w < nw) issue_alpha();
  }

  finish_pass(v, stp, s_lo, s_hi, act_lo, act_hi, res, pass_idx);
}

template <int NSTAGE>
static void launch_siso_pass_n(const TdecView& v, int pass_idx, cudaStream_t stream)
{
  const size_t smem = NSTAGE * ring::STAGE_BYTES + NSTAGE * 8;
  dim3         grid((unsigned)v.ntiles), block(32);
  static bool  attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(tdec_siso_pass_kernel<false, true, NSTAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(tdec_siso_pass_kernel<true, false, NSTAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(tdec_siso_pass_kernel<false, false, NSTAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_done = true;
  }
  if (pass_idx == 0) {
    tdec_siso_pass_kernel<false, true, NSTAGE><<<grid, block, smem, stream>>>(v, pass_idx);
  } else if (pass_idx & 1) {
    tdec_siso_pass_kernel<true, false, NSTAGE><<<grid, block, smem, stream>>>(v, pass_idx);
  } else {
    tdec_siso_pass_kernel<false, false, NSTAGE><<<grid, block, smem, stream>>>(v, pass_idx);
  }
}

void launch_siso_pass(const TdecView& v, int pass_idx, cudaStream_t stream)
{
  static int nstage = -1;
  if (nstage < 0) {
    const char* e = getenv("SRSLTE_B200_TDEC_STAGES"); // tuning knob, 4 / 6 / 8
    nstage        = e ? atoi(e) : 6;
  }
  switch (nstage) {
    case 4:
      launch_siso_pass_n<4>(v, pass_idx, stream);
      break;
    case 8:
      launch_siso_pass_n<8>(v, pass_idx, stream);
      break;
    default:
      launch_siso_pass_n<6>(v, pass_idx, stream);
      break;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// natural -> tiled.  Block = (tile, chunk of 32 trellis rows).  Phase 1 stages the 64 blocks' 96 contiguous int16
// each through shared memory with 8-byte loads (a block's natural vector is only 8-byte aligned: (3K+12)*2 bytes);
// phase 2 emits one uint4 (4 rows x 2 blocks) per thread and stream.
constexpr int LOAD_ROWS = 32;
constexpr int LOAD_PITCH = 3 * LOAD_ROWS + 4; // int16 per staged block, +4 keeps 8-byte alignment and skews banks

// `offsets` (optional): int16 offset of each block's vector inside llr (soft buffers scattered in a HARQ pool); without it
// the vectors are contiguous, block cb at cb*(3K+12).  ALIGNED8: every vector starts on an 8-byte boundary.
template <bool ALIGNED8>
__global__ void __launch_bounds__(256)
    tdec_load_natural_kernel(TdecView v, const int16_t* __restrict__ llr, const uint64_t* __restrict__ offsets, uint32_t ncb)
{
  __shared__ __align__(16) int16_t sm[TDEC_TILE_CB * LOAD_PITCH];
  const int    tile  = blockIdx.y;
  const int    chunk = blockIdx.x;
  const int    K     = v.K;
  const int    k0    = chunk * LOAD_ROWS;
  const size_t nllr  = 3 * (size_t)K + 12;
  const int    tid   = threadIdx.x;

  if (k0 < K) {
    const int rows  = min(LOAD_ROWS, K - k0); // multiple of 8
    const int nvec  = rows * 3 / 4;           // 8-byte vectors per block in this chunk
    for (int idx = tid; idx < TDEC_TILE_CB * nvec; idx += 256) {
      const int      c  = idx / nvec, q = idx % nvec;
      const uint32_t cb = (uint32_t)tile * TDEC_TILE_CB + c;
      uint2          val = make_uint2(0u, 0u);
      if (cb < ncb) {
        const int16_t* src = llr + (offsets ? offsets[cb] : cb * nllr) + 3 * (size_t)k0 + 4 * (size_t)q;
        if (ALIGNED8) {
          val = __ldcs(reinterpret_cast<const uint2*>(src));
        } else {
          val.x = (uint32_t)(uint16_t)src[0] | ((uint32_t)(uint16_t)src[1] << 16);
          val.y = (uint32_t)(uint16_t)src[2] | ((uint32_t)(uint16_t)src[3] << 16);
        }
      }
      *reinterpret_cast<uint2*>(&sm[c * LOAD_PITCH + 4 * q]) = val;
    }
    __syncthreads();
    const int lane = tid & 31;
    for (int r4 = tid >> 5; r4 < rows / 4; r4 += 8) {
      uint32_t w[3][4];
#pragma unroll
      for (int t = 0; t < 4; t++) {
#pragma unroll
        for (int s = 0; s < 3; s++) {
          const int o = 3 * (4 * r4 + t) + s;
          w[s][t]     = pack2(sm[(2 * lane) * LOAD_PITCH + o], sm[(2 * lane + 1) * LOAD_PITCH + o]);
        }
      }
      const size_t row = vec_row(v, tile, k0 / 4 + r4, lane);
      v.S[row]         = u4{w[0][0], w[0][1], w[0][2], w[0][3]};
      v.P0[row]        = u4{w[1][0], w[1][1], w[1][2], w[1][3]};
      v.P1[row]        = u4{w[2][0], w[2][1], w[2][2], w[2][3]};
    }
  } else {
    // the chunk past the payload carries the 12 tail values: rows K..K+3 of S/P0/P1 and S2T
    if (tid < 32) {
      const int lane = tid;
      uint32_t  w[4][4];
#pragma unroll
      for (int s = 0; s < 4; s++) {
#pragma unroll
        for (int t = 0; t < 4; t++) {
          int16_t a = 0, b = 0;
          const uint32_t cb0 = (uint32_t)tile * TDEC_TILE_CB + 2 * lane;
          if (cb0 < ncb) a = natural_pick(llr + (offsets ? offsets[cb0] : cb0 * nllr), K, s, K + t);
          if (cb0 + 1 < ncb) b = natural_pick(llr + (offsets ? offsets[cb0 + 1] : (cb0 + 1) * nllr), K, s, K + t);
          w[s][t] = pack2(a, b);
        }
      }
      const size_t row = vec_row(v, tile, K / 4, lane);
      v.S[row]         = u4{w[0][0], w[0][1], w[0][2], w[0][3]};
      v.P0[row]        = u4{w[1][0], w[1][1], w[1][2], w[1][3]};
      v.P1[row]        = u4{w[2][0], w[2][1], w[2][2], w[2][3]};
      v.S2T[(size_t)tile * 32 + lane] = u4{w[3][0], w[3][1], w[3][2], w[3][3]};
      // arm the per-block state for a fresh decode (srsran_tdec_new_cb, turbodecoder.c:510-525)
      const uint32_t cb0 = (uint32_t)tile * TDEC_TILE_CB + 2 * lane;
      v.status[cb0]      = CbStatus{(uint8_t)(cb0 < ncb), 0, 0, 0};
      v.status[cb0 + 1]  = CbStatus{(uint8_t)(cb0 + 1 < ncb), 0, 0, 0};
    }
  }
}

void launch_load_natural(const TdecView& v,
                         const int16_t*  llr_dev,
                         const uint64_t* offsets_dev,
                         bool            aligned8,
                         uint32_t        ncb,
                         cudaStream_t    stream)
{
  const int chunks = (v.K + LOAD_ROWS - 1) / LOAD_ROWS + 1; // +1: the tail chunk
  dim3      grid((unsigned)chunks, (unsigned)v.ntiles), block(256);
  if (aligned8) {
    tdec_load_natural_kernel<true><<<grid, block, 0, stream>>>(v, llr_dev, offsets_dev, ncb);
  } else {
    tdec_load_natural_kernel<false><<<grid, block, 0, stream>>>(v, llr_dev, offsets_dev, ncb);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// HB -> bytes.  Block = (tile, 32 output bytes per code block); decisions are transposed through shared memory so
// that both the HB reads (64 B rows) and the byte writes (32 B runs per block) are contiguous.
__global__ void __launch_bounds__(256) tdec_decide_kernel(TdecView v,
                                                          const uint16_t* __restrict__ qpp_rev,
                                                          uint8_t* __restrict__ out,
                                                          uint8_t* __restrict__ crc_ok,
                                                          uint8_t* __restrict__ npass,
                                                          uint8_t* __restrict__ npass_run,
                                                          uint32_t ncb)
{
  __shared__ uint8_t sm[TDEC_TILE_CB][33];
  const int tile = blockIdx.y;
  const int jb0  = blockIdx.x * 32;
  const int nb   = v.K / 8;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

  for (int j = wid; j < 32; j += 8) {
    const int jb = jb0 + j;
    if (jb < nb) {
      sm[2 * lane][j]     = decide_byte(v, qpp_rev, tile * TDEC_TILE_CB + 2 * lane, jb);
      sm[2 * lane + 1][j] = decide_byte(v, qpp_rev, tile * TDEC_TILE_CB + 2 * lane + 1, jb);
    }
  }
  __syncthreads();
  for (int c = wid; c < TDEC_TILE_CB; c += 8) {
    const uint32_t cb = (uint32_t)tile * TDEC_TILE_CB + c;
    const int      jb = jb0 + lane;
    if (cb < ncb && jb < nb) {
      out[(size_t)cb * nb + jb] = sm[c][lane];
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < TDEC_TILE_CB) {
    const uint32_t cb = (uint32_t)tile * TDEC_TILE_CB + threadIdx.x;
    if (cb < ncb) {
      const CbStatus s = v.status[cb];
      if (crc_ok) crc_ok[cb] = s.crc_ok;
      // the caller's loop counter (sch.c:431-432): pass at which the CRC matched, else the passes spent
      if (npass) npass[cb] = s.crc_ok ? s.npass_crc : s.npass_run;
      if (npass_run) npass_run[cb] = s.npass_run;
    }
  }
}

void launch_decide(const TdecView& v,
                   const uint16_t* qpp_rev_dev,
                   uint8_t*        out_dev,
                   uint8_t*        crc_ok_dev,
                   uint8_t*        npass_dev,
                   uint8_t*        npass_run_dev,
                   uint32_t        ncb,
                   cudaStream_t    stream)
{
  dim3 grid((unsigned)((v.K / 8 + 31) / 32), (unsigned)v.ntiles), block(256);
  tdec_decide_kernel<<<grid, block, 0, stream>>>(v, qpp_rev_dev, out_dev, crc_ok_dev, npass_dev, npass_run_dev, ncb);
}

} // namespace b200
