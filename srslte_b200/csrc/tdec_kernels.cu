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
// SISO pass kernel.  One CTA of two warps per tile of 64 code blocks; the warps work from the two ends of the trellis
// (tdec_core.h).  With the 65,536-block benchmark batch that is 2048 independent warps, ~3.5 per scheduler, each a long
// serial recursion, so throughput hangs on (a) instruction-level parallelism inside a trellis step and (b) keeping
// enough bytes in flight per warp.  (b): every window of 8 steps is streamed through a per-warp shared-memory ring
// filled with cp.async (LDGSTS): the copies do not occupy register scoreboards (a first version prefetched into
// registers with LDG and sat at 61 % long-scoreboard stalls), take per-lane addresses (DEC2's permuted E rows are one
// 4-byte copy per lane and row, no uniform-register serialisation as with per-lane bulk copies), and most bytes are
// lane-private so the only cross-lane hand-over is one __syncwarp per window.
//
// Ring stage:  P | S (DEC1) | E 1 KB | CK 1 KB | 8 CRC weights 64 B | 8 QPP entries 16 B      (P, S: 512 B int8 / 1 KB int16)
namespace ring {
constexpr uint32_t RING_BYTES = 12800; // per warp; 7 CTAs x 2 warps x 12.5 KB = 175 KB of the SM's shared memory
template <bool DEC2, bool IN8>
struct Lay {
  static constexpr uint32_t SP      = IN8 ? 512u : 1024u;
  static constexpr uint32_t OFF_P   = 0;
  static constexpr uint32_t OFF_S   = SP;
  static constexpr uint32_t OFF_E   = DEC2 ? SP : 2 * SP;
  static constexpr uint32_t OFF_CK  = OFF_E + 1024;
  static constexpr uint32_t OFF_CRC = OFF_CK + 1024;
  static constexpr uint32_t OFF_QPP = OFF_CRC + 64;
  static constexpr uint32_t BYTES   = (OFF_QPP + 16 + 127) / 128 * 128;
  static constexpr int      NST     = (int)(RING_BYTES / BYTES);
};
} // namespace ring

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void cp16(uint32_t dst, const void* src)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(__cvta_generic_to_global(src)) : "memory");
}
__device__ __forceinline__ void cp4(uint32_t dst, const void* src)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(__cvta_generic_to_global(src)) : "memory");
}
__device__ __forceinline__ void cp_commit()
{
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_wait()
{
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <bool DEC2, bool FIRST, bool IN8>
struct WarpRing {
  using L = ring::Lay<DEC2, IN8>;
  uint8_t*        gen;  // generic pointer to stage 0
  uint32_t        base; // shared-space address of stage 0
  int             lane;
  const u4*       gS;   // lane's uint4 in row 0 of the tile
  const u4*       gP;
  const uint8_t*  gE;   // tile base of E (bytes)
  const u4*       gCK;  // lane's uint4 of checkpoint 0, half 0
  const uint8_t*  gCRC; // syndrome weights in this decoder's visiting order, or nullptr
  const uint16_t* qpp;
  uint32_t        slot = 0;

  __device__ WarpRing(const TdecView& v, uint8_t* smem, int tile, int lane_) : lane(lane_)
  {
    gen  = smem;
    base = smem_u32(smem);
    if (IN8) {
      gS = v.S8 + row8(v, tile, 0, lane);
      gP = (DEC2 ? v.P18 : v.P08) + row8(v, tile, 0, lane);
    } else {
      gS = v.S + vec_row(v, tile, 0, lane);
      gP = (DEC2 ? v.P1 : v.P0) + vec_row(v, tile, 0, lane);
    }
    gE   = reinterpret_cast<const uint8_t*>(v.E + e_idx(v, tile, 0, 0));
    gCK  = v.CK + ck_idx(v, tile, 0, 0, lane);
    gCRC = reinterpret_cast<const uint8_t*>(DEC2 ? v.crc_perm : v.crc_nat);
    qpp  = v.qpp_fwd;
  }

  // Enqueue the copies of window w into the next stage (one cp.async group).  PH2 adds the checkpoint, the CRC weights
  // and (DEC2) the interleaver entries q = PI(8w..8w+7), which the caller fetched one issue ahead.
  template <bool PH2>
  __device__ __forceinline__ void issue(uint32_t w, const u4& q)
  {
    const uint32_t st = base + (slot % L::NST) * L::BYTES;
    const uint32_t l16 = (uint32_t)lane * 16u;
    if (IN8) {
      cp16(st + L::OFF_P + l16, gP + (size_t)w * 32u);
      if (!DEC2) cp16(st + L::OFF_S + l16, gS + (size_t)w * 32u);
    } else {
      cp16(st + L::OFF_P + l16, gP + (size_t)(2u * w) * 32u);
      cp16(st + L::OFF_P + 512u + l16, gP + (size_t)(2u * w + 1u) * 32u);
      if (!DEC2) {
        cp16(st + L::OFF_S + l16, gS + (size_t)(2u * w) * 32u);
        cp16(st + L::OFF_S + 512u + l16, gS + (size_t)(2u * w + 1u) * 32u);
      }
    }
    if (!FIRST) {
      if (!DEC2) { // eight consecutive 128-byte rows = 1 KB
        cp16(st + L::OFF_E + l16, gE + (size_t)w * 1024u + l16);
        cp16(st + L::OFF_E + 512u + l16, gE + (size_t)w * 1024u + 512u + l16);
      } else {
#pragma unroll
        for (int t = 0; t < 8; t++) {
          cp4(st + L::OFF_E + (uint32_t)t * 128u + (uint32_t)lane * 4u, gE + (size_t)win_pi(q, t) * 128u + (uint32_t)lane * 4u);
        }
      }
    }
    if (PH2) {
      cp16(st + L::OFF_CK + l16, gCK + (size_t)(2u * w) * 32u);
      cp16(st + L::OFF_CK + 512u + l16, gCK + (size_t)(2u * w + 1u) * 32u);
      if (gCRC != nullptr && lane < 4) cp16(st + L::OFF_CRC + l16, gCRC + (size_t)w * 64u + l16);
      if (DEC2 && lane == 0) *reinterpret_cast<u4*>(gen + (slot % L::NST) * L::BYTES + L::OFF_QPP) = q;
    }
    slot++;
  }

  __device__ __forceinline__ const uint8_t* stage(uint32_t i) const { return gen + (i % L::NST) * L::BYTES; }

  __device__ __forceinline__ void read(WinRegs& r, const uint8_t* st) const
  {
    u4       s[2] = {}, p[2] = {};
    uint32_t e[8] = {};
    p[0] = *reinterpret_cast<const u4*>(st + L::OFF_P + lane * 16);
    if (!IN8) p[1] = *reinterpret_cast<const u4*>(st + L::OFF_P + 512 + lane * 16);
    if (!DEC2) {
      s[0] = *reinterpret_cast<const u4*>(st + L::OFF_S + lane * 16);
      if (!IN8) s[1] = *reinterpret_cast<const u4*>(st + L::OFF_S + 512 + lane * 16);
    }
    if (!FIRST) {
#pragma unroll
      for (int t = 0; t < 8; t++) e[t] = *reinterpret_cast<const uint32_t*>(st + L::OFF_E + t * 128 + lane * 4);
    }
    win_unpack<DEC2, FIRST, IN8>(r, s, p, e);
  }
  __device__ __forceinline__ void read_ck(uint32_t c[8], const uint8_t* st) const
  {
    const u4 c0 = *reinterpret_cast<const u4*>(st + L::OFF_CK + lane * 16);
    const u4 c1 = *reinterpret_cast<const u4*>(st + L::OFF_CK + 512 + lane * 16);
    c[0] = c0.x; c[1] = c0.y; c[2] = c0.z; c[3] = c0.w; c[4] = c1.x; c[5] = c1.y; c[6] = c1.z; c[7] = c1.w;
  }
};

__device__ __forceinline__ u4 ldg_q(const uint16_t* qpp, uint32_t w)
{
  const uint4 t = __ldg(reinterpret_cast<const uint4*>(qpp + 8u * w));
  return u4{t.x, t.y, t.z, t.w};
}

template <bool DEC2, bool FIRST, bool IN8>
__global__ void __maxnreg__(144) tdec_siso_pass_kernel(TdecView v, int pass_idx)
{
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ LaneResult xres[32];
  const int tile = blockIdx.x;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  if (tile >= v.ntiles || v.fmt[tile] != (IN8 ? 0 : 1)) return;
  // whole tile finished (early stop): nothing to do.  Both warps read the same flags, so the exit is CTA-uniform.
  CbStatus*  stp  = v.status + ((size_t)tile * TDEC_TILE_CB + 2 * (size_t)lane);
  CbStatus   s_lo = stp[0], s_hi = stp[1];
  const bool act_lo = s_lo.active != 0, act_hi = s_hi.active != 0;
  if (__ballot_sync(0xFFFFFFFFu, act_lo || act_hi) == 0u) return;

  using RG = WarpRing<DEC2, FIRST, IN8>;
  constexpr int NST = RG::L::NST;
  const int      K  = v.K;
  const uint32_t nw = (uint32_t)K / 8u, ws = (uint32_t)v.ws;
  RG             rg(v, smem + warp * ring::RING_BYTES, tile, lane);
  u4* const       ckw = v.CK + ck_idx(v, tile, 0, 0, lane);  // checkpoint w, half h at ckw[(2w+h)*32]
  uint32_t* const ew  = v.E + e_idx(v, tile, 0, lane);       // row k at ew[k*32]
  uint16_t* const hbw = v.HB + hb_idx(v, tile, 0, lane);     // window w at hbw[w*32]
  const CrcPow*   have_crc = DEC2 ? v.crc_perm : v.crc_nat;

  WinRegs    r;
  LaneResult res = {0u, 0u};
  uint32_t   M[8]; // warp F: alpha, warp B: beta (un-normalised at window boundaries)
  u4         qn = {};

  if (warp == 0) {
    // ================= warp F =================
    // ---- phase 1: forward recursion over windows [0, ws), alpha checkpoints ----
    uint32_t nxt = 0;
    if (DEC2) qn = ldg_q(v.qpp_fwd, 0);
#pragma unroll 1
    for (int i = 0; i < NST; i++) {
      if (nxt < ws) {
        const u4 q = qn;
        if (DEC2 && nxt + 1 < ws) qn = ldg_q(v.qpp_fwd, nxt + 1);
        rg.template issue<false>(nxt, q);
        nxt++;
      }
      cp_commit();
    }
    M[0] = 0;
#pragma unroll
    for (int i = 1; i < 8; i++) M[i] = NEG_INF2;
#pragma unroll 1
    for (uint32_t w = 0; w < ws; w++) {
      cp_wait<NST - 1>();
      __syncwarp();
      rg.read(r, rg.stage(w));
      ckw[(2u * w) * 32u]      = u4{M[0], M[1], M[2], M[3]};
      ckw[(2u * w + 1u) * 32u] = u4{M[4], M[5], M[6], M[7]};
      alpha_window(M, r);
      __syncwarp(); // every lane has consumed the stage before it is refilled
      if (nxt < ws) {
        const u4 q = qn;
        if (DEC2 && nxt + 1 < ws) qn = ldg_q(v.qpp_fwd, nxt + 1);
        rg.template issue<false>(nxt, q);
        nxt++;
      }
      cp_commit();
    }
  } else {
    // ================= warp B =================
    // ---- phase 1: tail + backward recursion over windows [ws, nw), beta checkpoints ----
    uint32_t nxt = nw; // windows are issued nw-1, nw-2, ..., ws
    if (DEC2) qn = ldg_q(v.qpp_fwd, nw - 1);
#pragma unroll 1
    for (int i = 0; i < NST; i++) {
      if (nxt > ws) {
        nxt--;
        const u4 q = qn;
        if (DEC2 && nxt > ws) qn = ldg_q(v.qpp_fwd, nxt - 1);
        rg.template issue<false>(nxt, q);
      }
      cp_commit();
    }
    {
      const u4 pt = IN8 ? rg.gP[(size_t)nw * 32u] : rg.gP[(size_t)(K / 4) * 32u];
      M[0]        = 0;
#pragma unroll
      for (int i = 1; i < 8; i++) M[i] = NEG_INF2;
      if (DEC2) {
        const u4 s2 = v.S2T[(size_t)tile * 32 + lane];
#pragma unroll
        for (int t = 2; t >= 0; t--) {
          const u4       pq[2] = {pt, pt};
          const uint32_t x = u4_get(s2, t), y = win_val<IN8>(pq, t);
          beta_step(M, x, y, add2(x, y));
        }
      } else {
        const u4 stl = IN8 ? rg.gS[(size_t)nw * 32u] : rg.gS[(size_t)(K / 4) * 32u];
        beta_tail<IN8>(M, stl, pt);
      }
    }
    ckw[(2u * (nw - 1)) * 32u]      = u4{M[0], M[1], M[2], M[3]};
    ckw[(2u * (nw - 1) + 1u) * 32u] = u4{M[4], M[5], M[6], M[7]};
    uint32_t i = 0;
#pragma unroll 1
    for (uint32_t w = nw - 1;; w--, i++) {
      cp_wait<NST - 1>();
      __syncwarp();
      rg.read(r, rg.stage(i));
      beta_window(M, r);
      if (w > ws) {
        ckw[(2u * (w - 1)) * 32u]      = u4{M[0], M[1], M[2], M[3]};
        ckw[(2u * (w - 1) + 1u) * 32u] = u4{M[4], M[5], M[6], M[7]};
        normalise(M);
      }
      __syncwarp();
      if (nxt > ws) {
        nxt--;
        const u4 q = qn;
        if (DEC2 && nxt > ws) qn = ldg_q(v.qpp_fwd, nxt - 1);
        rg.template issue<false>(nxt, q);
      }
      cp_commit();
      if (w == ws) break;
    }
  }

  // the checkpoints were written by the other warp
  cp_wait<0>();
  __syncthreads();

  WinOut   o;
  uint32_t c[8];
  if (warp == 0) {
    // ---- phase 2: windows ws..nw-1 upwards ----
    rg.slot      = 0;
    uint32_t nxt = ws;
    if (DEC2) qn = ldg_q(v.qpp_fwd, ws);
#pragma unroll 1
    for (int i = 0; i < NST; i++) {
      if (nxt < nw) {
        const u4 q = qn;
        if (DEC2 && nxt + 1 < nw) qn = ldg_q(v.qpp_fwd, nxt + 1);
        rg.template issue<true>(nxt, q);
        nxt++;
      }
      cp_commit();
    }
#pragma unroll 1
    for (uint32_t w = ws; w < nw; w++) {
      cp_wait<NST - 1>();
      __syncwarp();
      const uint8_t* st = rg.stage(w - ws);
      rg.read(r, st);
      rg.read_ck(c, st);
      u4 q = {};
      if (DEC2) q = *reinterpret_cast<const u4*>(st + RG::L::OFF_QPP);
      fwd_window(M, c, 8u * w + 8u < (uint32_t)K, r, have_crc ? reinterpret_cast<const CrcPow*>(st + RG::L::OFF_CRC) : nullptr, res, o);
#pragma unroll
      for (int t = 0; t < 8; t++) ew[(DEC2 ? win_pi(q, t) : 8u * w + (uint32_t)t) * 32u] = o.enew[t];
      hb_store(&hbw[w * 32u], o.bits, act_lo, act_hi);
      __syncwarp();
      if (nxt < nw) {
        const u4 qi = qn;
        if (DEC2 && nxt + 1 < nw) qn = ldg_q(v.qpp_fwd, nxt + 1);
        rg.template issue<true>(nxt, qi);
        nxt++;
      }
      cp_commit();
    }
  } else {
    // ---- phase 2: windows ws-1..0 downwards ----
    rg.slot      = 0;
    uint32_t nxt = ws;
    if (DEC2) qn = ldg_q(v.qpp_fwd, ws - 1);
#pragma unroll 1
    for (int i = 0; i < NST; i++) {
      if (nxt > 0) {
        nxt--;
        const u4 q = qn;
        if (DEC2 && nxt > 0) qn = ldg_q(v.qpp_fwd, nxt - 1);
        rg.template issue<true>(nxt, q);
      }
      cp_commit();
    }
    uint32_t i = 0;
#pragma unroll 1
    for (uint32_t w = ws - 1;; w--, i++) {
      cp_wait<NST - 1>();
      __syncwarp();
      const uint8_t* st = rg.stage(i);
      rg.read(r, st);
      rg.read_ck(c, st);
      u4 q = {};
      if (DEC2) q = *reinterpret_cast<const u4*>(st + RG::L::OFF_QPP);
      bwd_window(M, c, r, have_crc ? reinterpret_cast<const CrcPow*>(st + RG::L::OFF_CRC) : nullptr, res, o);
#pragma unroll
      for (int t = 0; t < 8; t++) ew[(DEC2 ? win_pi(q, t) : 8u * w + (uint32_t)t) * 32u] = o.enew[t];
      hb_store(&hbw[w * 32u], o.bits, act_lo, act_hi);
      __syncwarp();
      if (nxt > 0) {
        nxt--;
        const u4 qi = qn;
        if (DEC2 && nxt > 0) qn = ldg_q(v.qpp_fwd, nxt - 1);
        rg.template issue<true>(nxt, qi);
      }
      cp_commit();
      if (w == 0) break;
    }
    xres[lane] = res;
  }
  __syncthreads();
  if (warp == 0) {
    res.crc_lo16x2 ^= xres[lane].crc_lo16x2;
    res.crc_hi8x2 ^= xres[lane].crc_hi8x2;
    finish_pass(v, stp, s_lo, s_hi, act_lo, act_hi, res, pass_idx);
  }
}

template <bool IN8>
static void launch_siso_pass_fmt(const TdecView& v, int pass_idx, cudaStream_t stream)
{
  const size_t smem = 2 * ring::RING_BYTES;
  dim3         grid((unsigned)v.ntiles), block(64);
  static bool  attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(tdec_siso_pass_kernel<false, true, IN8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(tdec_siso_pass_kernel<true, false, IN8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(tdec_siso_pass_kernel<false, false, IN8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_done = true;
  }
  if (pass_idx == 0) {
    tdec_siso_pass_kernel<false, true, IN8><<<grid, block, smem, stream>>>(v, pass_idx);
  } else if (pass_idx & 1) {
    tdec_siso_pass_kernel<true, false, IN8><<<grid, block, smem, stream>>>(v, pass_idx);
  } else {
    tdec_siso_pass_kernel<false, false, IN8><<<grid, block, smem, stream>>>(v, pass_idx);
  }
}

// Two launches per pass: the tiles held in the int8 format and those held in int16 (a CTA whose tile is in the other
// format exits at once; which tiles are which is only known on the device, after the load kernel ran).
void launch_siso_pass(const TdecView& v, int pass_idx, cudaStream_t stream)
{
  launch_siso_pass_fmt<true>(v, pass_idx, stream);
  launch_siso_pass_fmt<false>(v, pass_idx, stream);
}

// ---------------------------------------------------------------------------------------------------------------
// natural -> tiled.  Block = (tile, chunk of 32 trellis rows).  Phase 1 stages the 64 blocks' 96 contiguous int16
// each through shared memory with 8-byte loads (a block's natural vector is only 8-byte aligned: (3K+12)*2 bytes);
// phase 2 emits one uint4 per thread and stream.
//   tdec_load8_kernel   always runs: writes the int8 arrays (8 rows x 2 blocks per uint4) and raises fmt[tile] when a
//                       value of the tile does not fit int8; also arms the per-block state.
//   tdec_load16_kernel  runs after it and fills the int16 arrays (4 rows x 2 blocks per uint4) of the raised tiles only.
constexpr int LOAD_ROWS = 32;
constexpr int LOAD_PITCH = 3 * LOAD_ROWS + 4; // int16 per staged block, +4 keeps 8-byte alignment and skews banks

__device__ __forceinline__ bool fits8(int16_t a)
{
  return (int16_t)(int8_t)a == a;
}

// `offsets` (optional): int16 offset of each block's vector inside llr (soft buffers scattered in a HARQ pool); without it
// the vectors are contiguous, block cb at cb*(3K+12).  ALIGNED8: every vector starts on an 8-byte boundary.
// Returns (to all threads) whether every staged value fits int8.
template <bool ALIGNED8>
__device__ __forceinline__ bool load_stage_chunk(int16_t* sm, const TdecView& v, const int16_t* __restrict__ llr,
                                                 const uint64_t* __restrict__ offsets, uint32_t ncb, int tile, int k0, int rows)
{
  const size_t nllr = 3 * (size_t)v.K + 12;
  const int    nvec = rows * 3 / 4; // 8-byte vectors per block in this chunk
  bool         bad  = false;
  for (int idx = threadIdx.x; idx < TDEC_TILE_CB * nvec; idx += blockDim.x) {
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
    bad |= !fits8(lo16(val.x)) || !fits8(hi16(val.x)) || !fits8(lo16(val.y)) || !fits8(hi16(val.y));
    *reinterpret_cast<uint2*>(&sm[c * LOAD_PITCH + 4 * q]) = val;
  }
  return __syncthreads_or(bad) == 0;
}

template <bool ALIGNED8>
__global__ void __launch_bounds__(256)
    tdec_load8_kernel(TdecView v, const int16_t* __restrict__ llr, const uint64_t* __restrict__ offsets, uint32_t ncb)
{
  __shared__ __align__(16) int16_t sm[TDEC_TILE_CB * LOAD_PITCH];
  const int    tile  = blockIdx.y;
  const int    chunk = blockIdx.x;
  const int    K     = v.K;
  const int    k0    = chunk * LOAD_ROWS;
  const size_t nllr  = 3 * (size_t)K + 12;
  const int    tid   = threadIdx.x;

  if (k0 < K) {
    const int  rows = min(LOAD_ROWS, K - k0); // multiple of 8
    const bool ok   = load_stage_chunk<ALIGNED8>(sm, v, llr, offsets, ncb, tile, k0, rows);
    if (!ok && tid == 0) atomicOr(v.fmt + tile, 1u);
    // (window, stream, lane) -> one uint4 of 8 rows x 2 blocks
    for (int idx = tid; idx < (rows / 8) * 3 * 32; idx += 256) {
      const int lane = idx & 31, s = (idx >> 5) % 3, w8 = idx / 96;
      uint32_t  w[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        uint32_t word = 0;
#pragma unroll
        for (int j = 0; j < 2; j++) {
          const int      o  = 3 * (8 * w8 + 2 * q + j) + s;
          const uint32_t lo = (uint8_t)sm[(2 * lane) * LOAD_PITCH + o], hi = (uint8_t)sm[(2 * lane + 1) * LOAD_PITCH + o];
          word |= (lo | (hi << 8)) << (16 * j);
        }
        w[q] = word;
      }
      u4* dst = s == 0 ? v.S8 : (s == 1 ? v.P08 : v.P18);
      dst[row8(v, tile, k0 / 8 + w8, lane)] = u4{w[0], w[1], w[2], w[3]};
    }
  } else {
    // the chunk past the payload carries the 12 tail values: row K/8 of S8/P08/P18 and S2T
    if (tid < 32) {
      const int lane = tid;
      uint32_t  w16[4][4];
      bool      bad = false;
#pragma unroll
      for (int s = 0; s < 4; s++) {
#pragma unroll
        for (int t = 0; t < 4; t++) {
          int16_t a = 0, b = 0;
          const uint32_t cb0 = (uint32_t)tile * TDEC_TILE_CB + 2 * lane;
          if (cb0 < ncb) a = natural_pick(llr + (offsets ? offsets[cb0] : cb0 * nllr), K, s, K + t);
          if (cb0 + 1 < ncb) b = natural_pick(llr + (offsets ? offsets[cb0 + 1] : (cb0 + 1) * nllr), K, s, K + t);
          w16[s][t] = pack2(a, b);
          if (s < 3) bad |= !fits8(a) || !fits8(b); // S2T stays int16
        }
      }
      if (__any_sync(0xFFFFFFFFu, bad) && lane == 0) atomicOr(v.fmt + tile, 1u);
#pragma unroll
      for (int s = 0; s < 3; s++) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int t = 0; t < 4; t++) {
          const uint32_t lo = w16[s][t] & 0xFFu, hi = (w16[s][t] >> 16) & 0xFFu;
          w[t >> 1] |= (lo | (hi << 8)) << (16 * (t & 1));
        }
        u4* dst = s == 0 ? v.S8 : (s == 1 ? v.P08 : v.P18);
        dst[row8(v, tile, K / 8, lane)] = u4{w[0], w[1], w[2], w[3]};
      }
      v.S2T[(size_t)tile * 32 + lane] = u4{w16[3][0], w16[3][1], w16[3][2], w16[3][3]};
      // arm the per-block state for a fresh decode (srsran_tdec_new_cb, turbodecoder.c:510-525)
      const uint32_t cb0 = (uint32_t)tile * TDEC_TILE_CB + 2 * lane;
      v.status[cb0]      = CbStatus{(uint8_t)(cb0 < ncb), 0, 0, 0};
      v.status[cb0 + 1]  = CbStatus{(uint8_t)(cb0 + 1 < ncb), 0, 0, 0};
    }
  }
}

template <bool ALIGNED8>
__global__ void __launch_bounds__(256)
    tdec_load16_kernel(TdecView v, const int16_t* __restrict__ llr, const uint64_t* __restrict__ offsets, uint32_t ncb)
{
  __shared__ __align__(16) int16_t sm[TDEC_TILE_CB * LOAD_PITCH];
  const int    tile  = blockIdx.y;
  const int    chunk = blockIdx.x;
  const int    K     = v.K;
  const int    k0    = chunk * LOAD_ROWS;
  const size_t nllr  = 3 * (size_t)K + 12;
  const int    tid   = threadIdx.x;
  if (v.fmt[tile] == 0u) return; // the tile lives in the int8 arrays

  if (k0 < K) {
    const int rows = min(LOAD_ROWS, K - k0); // multiple of 8
    load_stage_chunk<ALIGNED8>(sm, v, llr, offsets, ncb, tile, k0, rows);
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
  } else if (tid < 32) {
    const int lane = tid;
    uint32_t  w[3][4];
#pragma unroll
    for (int s = 0; s < 3; s++) {
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
  cudaMemsetAsync(v.fmt, 0, (size_t)v.ntiles * sizeof(uint32_t), stream);
  if (aligned8) {
    tdec_load8_kernel<true><<<grid, block, 0, stream>>>(v, llr_dev, offsets_dev, ncb);
    tdec_load16_kernel<true><<<grid, block, 0, stream>>>(v, llr_dev, offsets_dev, ncb);
  } else {
    tdec_load8_kernel<false><<<grid, block, 0, stream>>>(v, llr_dev, offsets_dev, ncb);
    tdec_load16_kernel<false><<<grid, block, 0, stream>>>(v, llr_dev, offsets_dev, ncb);
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
