// CUDA kernels of the batched turbo decoder (sm_100a).  The arithmetic lives in tdec_core.h; this file maps it onto
// the grid and adds the two layout-conversion kernels at the ends of a decode.
//
//   tdec_load_natural_kernel   natural [cb][3K+12] int16 (turbodecoder_gen.c:238-258 order)  ->  S/P0/P1/S2T tiles
//   tdec_siso_pass_kernel      one SISO pass (turbodecoder_iter.h:72-144) for every still-active code block
//   tdec_decide_kernel         HB (visiting order of the last pass)  ->  packed bytes, natural order
//                              (turbodecoder.c:370-378 + turbodecoder_gen.c:260-277)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <type_traits>

#include "b200_runtime.h"
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

// Every global access below is (tile base, uniform across the CTA) + (32-bit byte offset): the bases can live in
// uniform registers and the per-window address arithmetic is 32-bit.
template <bool DEC2, bool FIRST, bool IN8>
struct WarpRing {
  using L = ring::Lay<DEC2, IN8>;
  static constexpr uint32_t RING_END = L::NST * L::BYTES;
  uint8_t*       gen;  // generic pointer to stage 0
  uint32_t       base; // shared-space address of stage 0
  uint32_t       l16, l4;
  const uint8_t* tS;   // tile bases (bytes)
  const uint8_t* tP;
  uint8_t*       tE;
  uint8_t*       tCK;
  uint8_t*       gHB;  // the decision array; hb_off = byte offset of this lane's pair (its home, not this tile)
  uint32_t       hb_off;
  const uint8_t* gCRC; // syndrome weights in this decoder's visiting order, or nullptr
  uint32_t       fill = 0, drain = 0; // byte offsets of the next stage to fill / to consume

  __device__ WarpRing(const TdecView& v, const TileDesc& td, uint8_t* smem, int lane, uint32_t hb0)
  {
    gen  = smem;
    base = smem_u32(smem);
    l16  = (uint32_t)lane * 16u;
    l4   = (uint32_t)lane * 4u;
    if (IN8) {
      tS = reinterpret_cast<const uint8_t*>(td.S8);
      tP = reinterpret_cast<const uint8_t*>(DEC2 ? td.P18 : td.P08);
    } else {
      tS = reinterpret_cast<const uint8_t*>(td.S);
      tP = reinterpret_cast<const uint8_t*>(DEC2 ? td.P1 : td.P0);
    }
    tE     = reinterpret_cast<uint8_t*>(td.E);
    tCK    = reinterpret_cast<uint8_t*>(td.CK);
    gHB    = reinterpret_cast<uint8_t*>(v.HB);
    hb_off = hb0 * 2u;
    gCRC   = reinterpret_cast<const uint8_t*>(DEC2 ? td.crc_perm : td.crc_nat);
  }

  __device__ __forceinline__ void restart() { fill = drain = 0; }

  // Enqueue the copies of window w into the next stage (the caller commits the cp.async group).  PH2 adds the
  // checkpoint, the CRC weights and (DEC2) the interleaver entries q = PI(8w..8w+7), fetched one issue ahead.
  template <bool PH2>
  __device__ __forceinline__ void issue(uint32_t w, const u4& q)
  {
    const uint32_t     st = base + fill;
    constexpr uint32_t WB = IN8 ? 512u : 1024u; // bytes of S / P per window
    cp16(st + L::OFF_P + l16, tP + (w * WB + l16));
    if (!IN8) cp16(st + L::OFF_P + 512u + l16, tP + (w * WB + 512u + l16));
    if (!DEC2) {
      cp16(st + L::OFF_S + l16, tS + (w * WB + l16));
      if (!IN8) cp16(st + L::OFF_S + 512u + l16, tS + (w * WB + 512u + l16));
    }
    if (!FIRST) {
      if (!DEC2) { // eight consecutive 128-byte rows = 1 KB
        cp16(st + L::OFF_E + l16, tE + (w * 1024u + l16));
        cp16(st + L::OFF_E + 512u + l16, tE + (w * 1024u + 512u + l16));
      } else {
#pragma unroll
        for (int t = 0; t < 8; t++) cp4(st + L::OFF_E + (uint32_t)t * 128u + l4, tE + (win_pi(q, t) * 128u + l4));
      }
    }
    if (PH2) {
      cp16(st + L::OFF_CK + l16, tCK + (w * 1024u + l16));
      cp16(st + L::OFF_CK + 512u + l16, tCK + (w * 1024u + 512u + l16));
      if (gCRC != nullptr && l16 < 64u) cp16(st + L::OFF_CRC + l16, gCRC + (w * 64u + l16));
      if (DEC2 && l16 == 0u) *reinterpret_cast<u4*>(gen + fill + L::OFF_QPP) = q;
    }
    fill = (fill + L::BYTES == RING_END) ? 0u : fill + L::BYTES;
  }

  // stage holding the oldest window; call after cp_wait + __syncwarp
  __device__ __forceinline__ const uint8_t* next_stage()
  {
    const uint8_t* st = gen + drain;
    drain             = (drain + L::BYTES == RING_END) ? 0u : drain + L::BYTES;
    return st;
  }

  __device__ __forceinline__ void read(WinRegs& r, const uint8_t* st) const
  {
    u4       s[2] = {}, p[2] = {};
    uint32_t e[8] = {};
    p[0] = *reinterpret_cast<const u4*>(st + L::OFF_P + l16);
    if (!IN8) p[1] = *reinterpret_cast<const u4*>(st + L::OFF_P + 512 + l16);
    if (!DEC2) {
      s[0] = *reinterpret_cast<const u4*>(st + L::OFF_S + l16);
      if (!IN8) s[1] = *reinterpret_cast<const u4*>(st + L::OFF_S + 512 + l16);
    }
    if (!FIRST) {
#pragma unroll
      for (int t = 0; t < 8; t++) e[t] = *reinterpret_cast<const uint32_t*>(st + L::OFF_E + t * 128 + l4);
    }
    win_unpack<DEC2, FIRST, IN8>(r, s, p, e);
  }
  __device__ __forceinline__ void read_ck(uint32_t c[8], const uint8_t* st) const
  {
    const u4 c0 = *reinterpret_cast<const u4*>(st + L::OFF_CK + l16);
    const u4 c1 = *reinterpret_cast<const u4*>(st + L::OFF_CK + 512 + l16);
    c[0] = c0.x; c[1] = c0.y; c[2] = c0.z; c[3] = c0.w; c[4] = c1.x; c[5] = c1.y; c[6] = c1.z; c[7] = c1.w;
  }
  __device__ __forceinline__ void store_ck(uint32_t w, const uint32_t c[8]) const
  {
    *reinterpret_cast<u4*>(tCK + (w * 1024u + l16))        = u4{c[0], c[1], c[2], c[3]};
    *reinterpret_cast<u4*>(tCK + (w * 1024u + 512u + l16)) = u4{c[4], c[5], c[6], c[7]};
  }
  // what is subtracted from the LLR of step t to form the new extrinsic (turbodecoder_iter.h:108,118): nothing on the first pass,
  // DEC2 its systematic input (already in a register), DEC1 the a-priori value, re-read from the stage
  __device__ __forceinline__ uint32_t e_in(const uint8_t* st, const WinRegs& r, int t) const
  {
    if (FIRST) return 0u;
    if (DEC2) return r.xs[t];
    // volatile: keep the load at the step that consumes it (hoisted to the top of the window it would pin 8 registers)
    uint32_t e;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(smem_u32(st) + L::OFF_E + (uint32_t)t * 128u + l4));
    return e;
  }
  // where the new extrinsic of step t of window w goes; q = the window's interleaver entries (DEC2)
  __device__ __forceinline__ uint32_t* e_out(uint32_t w, const u4& q, int t) const
  {
    return reinterpret_cast<uint32_t*>(tE + ((DEC2 ? win_pi(q, t) : 8u * w + (uint32_t)t) * 128u + l4));
  }
  // hard decisions of window w
  __device__ __forceinline__ void store_hb(uint32_t w, uint32_t bits, bool act_lo, bool act_hi) const
  {
    hb_store(reinterpret_cast<uint16_t*>(gHB + (size_t)hb_off + w * 64u), bits, act_lo, act_hi);
  }
};

__device__ __forceinline__ u4 ldg_q(const uint16_t* qpp, uint32_t w)
{
  const uint4 t = __ldg(reinterpret_cast<const uint4*>(qpp + 8u * w));
  return u4{t.x, t.y, t.z, t.w};
}

template <bool DEC2, bool FIRST, bool IN8>
__device__ __forceinline__ void siso_pass_tile(const TdecView& v, int pass_idx, uint8_t* smem, LaneResult* xres)
{
  const int tile = blockIdx.x;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  // whole tile finished or emptied by the re-packing (early stop): nothing to do.  Both warps read the same flags, so the
  // exit is CTA-uniform.
  const LaneMap lm   = v.lanes[(size_t)tile * 32 + lane];
  const bool    held = lm.st0 != LANE_EMPTY;
  CbStatus*     stp  = v.status + (held ? lm.st0 : 0u);
  const bool act_lo = held && stp[0].active != 0, act_hi = held && stp[1].active != 0; // the records themselves are re-read at the end of the pass
  if (__ballot_sync(0xFFFFFFFFu, act_lo || act_hi) == 0u) return;

  using RG = WarpRing<DEC2, FIRST, IN8>;
  constexpr int   NST = RG::L::NST;
  const TileDesc& td  = v.tiles[tile];
  const uint32_t  K   = td.K;
  const uint32_t  nw  = K / 8u, ws = (uint32_t)tdec_split((int)K, v.split_percent);
  RG              rg(v, td, smem + warp * ring::RING_BYTES, lane, lm.hb0);
  const bool      have_crc = (DEC2 ? td.crc_perm : td.crc_nat) != nullptr;
  const uint16_t* qpp_fwd  = td.qpp_fwd;

  WinRegs    r;
  LaneResult res = {0u, 0u};
  uint32_t   M[8]; // warp F: alpha, warp B: beta (un-normalised at window boundaries)
  u4         qn = {};

  // the windows a warp visits in a phase: first, first+dir, ..., last (dir = +1 or -1)
  auto run_phase = [&](auto ph2_tag, uint32_t first, uint32_t count, int dir, auto&& body) {
    constexpr bool PH2 = decltype(ph2_tag)::value;
    rg.restart();
    uint32_t nxt = first, left = count;
    if (DEC2) qn = ldg_q(qpp_fwd, first);
    auto issue_next = [&]() {
      if (left > 0) {
        const u4 q = qn;
        if (DEC2 && left > 1) qn = ldg_q(qpp_fwd, nxt + dir);
        rg.template issue<PH2>(nxt, q);
        nxt += dir;
        left--;
      }
      cp_commit();
    };
#pragma unroll 1
    for (int i = 0; i < NST; i++) issue_next();
    uint32_t w = first;
#pragma unroll 1
    for (uint32_t n = 0; n < count; n++, w += dir) {
      cp_wait<NST - 1>();
      __syncwarp();
      body(w, rg.next_stage());
      __syncwarp(); // every lane has consumed the stage before it is refilled
      issue_next();
    }
  };

  if (warp == 0) {
    // warp F, phase 1: forward recursion over windows [0, ws), alpha checkpoints
    M[0] = 0;
#pragma unroll
    for (int i = 1; i < 8; i++) M[i] = NEG_INF2;
    run_phase(std::false_type{}, 0u, ws, +1, [&](uint32_t w, const uint8_t* st) {
      rg.read(r, st);
      rg.store_ck(w, M);
      alpha_window(M, r);
    });
  } else {
    // warp B, phase 1: tail + backward recursion over windows [ws, nw), beta checkpoints
    {
      const u4 pt = *reinterpret_cast<const u4*>(rg.tP + (nw * (IN8 ? 512u : 1024u) + rg.l16));
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
        const u4 stl = *reinterpret_cast<const u4*>(rg.tS + (nw * (IN8 ? 512u : 1024u) + rg.l16));
        beta_tail<IN8>(M, stl, pt);
      }
    }
    rg.store_ck(nw - 1, M);
    run_phase(std::false_type{}, nw - 1, nw - ws, -1, [&](uint32_t w, const uint8_t* st) {
      rg.read(r, st);
      beta_window(M, r);
      if (w > ws) {
        rg.store_ck(w - 1, M);
        normalise(M);
      }
    });
  }

  // the checkpoints were written by the other warp
  cp_wait<0>();
  __syncthreads();

  WinOut   o;
  uint32_t c[8];
  if (warp == 0) {
    // warp F, phase 2: windows ws..nw-1 upwards
    run_phase(std::true_type{}, ws, nw - ws, +1, [&](uint32_t w, const uint8_t* st) {
      rg.read(r, st);
      rg.read_ck(c, st);
      u4 q = {};
      if (DEC2) q = *reinterpret_cast<const u4*>(st + RG::L::OFF_QPP);
      fwd_window(M, c, 8u * w + 8u < K, r, have_crc ? reinterpret_cast<const CrcPow*>(st + RG::L::OFF_CRC) : nullptr, res, o,
                 [&](int t, uint32_t L) { *rg.e_out(w, q, t) = sub2(L, rg.e_in(st, r, t)); });
      rg.store_hb(w, o.bits, act_lo, act_hi);
    });
  } else {
    // warp B, phase 2: windows ws-1..0 downwards
    run_phase(std::true_type{}, ws - 1, ws, -1, [&](uint32_t w, const uint8_t* st) {
      rg.read(r, st);
      rg.read_ck(c, st);
      u4 q = {};
      if (DEC2) q = *reinterpret_cast<const u4*>(st + RG::L::OFF_QPP);
      bwd_window(M, c, r, have_crc ? reinterpret_cast<const CrcPow*>(st + RG::L::OFF_CRC) : nullptr, res, o,
                 [&](int t, uint32_t L) { *rg.e_out(w, q, t) = sub2(L, rg.e_in(st, r, t)); });
      rg.store_hb(w, o.bits, act_lo, act_hi);
    });
    xres[lane] = res;
  }
  __syncthreads();
  if (warp == 0) {
    res.crc_lo16x2 ^= xres[lane].crc_lo16x2;
    res.crc_hi8x2 ^= xres[lane].crc_hi8x2;
    finish_pass(v, have_crc, stp, stp[0], stp[1], act_lo, act_hi, res, pass_idx);
  }
}

// One launch per pass; a CTA takes the int8 or the int16 code path according to its tile's format (which tiles are
// which is only known on the device, after the load kernels ran).
template <bool DEC2, bool FIRST>
__global__ void __maxnreg__(128) tdec_siso_pass_kernel(TdecView v, int pass_idx, int unless_flagged)
{
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ LaneResult xres[32];
  if ((int)blockIdx.x >= v.ntiles) return;
  if (unless_flagged && v.ctl[TDEC_CTL_USE_LL] != 0u) return; // the low-latency kernel, launched beside this one, owns the pass
  if (v.fmt[blockIdx.x] == 0u || v.tiles[blockIdx.x].S == nullptr) {
    siso_pass_tile<DEC2, FIRST, true>(v, pass_idx, smem, xres);
  } else {
    siso_pass_tile<DEC2, FIRST, false>(v, pass_idx, smem, xres);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Low-latency SISO pass: the same arithmetic, bit for bit, for batches that cannot fill the GPU with tiles (a single
// subframe's 13 code blocks are ONE tile; the last passes of an early-stop decode concern a handful of blocks).  There a
// pass of the kernel above is bound by the latency of its two warps: each runs its recursion AND the 7x heavier window
// work (rebuild + LLR) one window after the other, ~0.45 ms per K=6144 pass however empty the GPU is.
//
// Here a tile gets a whole SM: 16 warps.  Only the two recursions are sequential, so
//   warp 0  runs the forward recursion over ALL windows and leaves alpha_{8w}   for every window  (CKA, a per-block scratch)
//   warp 1  runs the tail + backward recursion over ALL windows, leaving beta_{8w+8} (un-normalised) (CKB = the tile's CK)
// and every window w whose two checkpoints exist is an independent piece of work -- rebuild its 8 beta vectors from CKB[w],
// 8 forward steps from CKA[w] with LLR output: exactly fwd_window() of the throughput kernel -- taken by whichever warp
// is free (the two recursion warps join when they are done).  The fronts meet in the middle of the trellis, so the work
// is handed out from the middle outwards.  4 recursions per step instead of 3 and twice the checkpoint traffic: worth it
// only when tiles are scarce (srsran_b200 engine: batches of at most one tile per SM, and the tail passes of an
// early-stop decode once the running blocks have been re-packed into that few tiles).
// The extrinsic array is updated in place like above: window w's rows are written by its worker only after BOTH
// recursions have consumed window w (doneA > w, doneB <= w), and no other window reads or writes those rows.
namespace ll {
constexpr int WARPS = 16;
template <bool DEC2, bool IN8>
struct Lay {
  static constexpr uint32_t SP      = IN8 ? 512u : 1024u;
  static constexpr uint32_t OFF_P   = 0;
  static constexpr uint32_t OFF_S   = SP;
  static constexpr uint32_t OFF_E   = DEC2 ? SP : 2 * SP;
  static constexpr uint32_t OFF_CKA = OFF_E + 1024;
  static constexpr uint32_t OFF_CKB = OFF_CKA + 1024;
  static constexpr uint32_t OFF_CRC = OFF_CKB + 1024;
  static constexpr uint32_t OFF_QPP = OFF_CRC + 64;
  static constexpr uint32_t BYTES   = (OFF_QPP + 16 + 127) / 128 * 128;
  static_assert(2 * BYTES <= 2 * 5248, "two worker stages per warp");
};
// a recursion warp alone on its issue port consumes a window every few hundred cycles: it needs its inputs requested a
// whole memory latency ahead, so its ring is deep (32 KB: 10-21 stages of P | S | E) while a worker warp's two stages are small
constexpr uint32_t REC_RING    = 32768;
constexpr uint32_t WORKER_RING = 2 * 5248; // two stages of the largest worker layout (int16 tiles, DEC1)
constexpr uint32_t SMEM_BYTES  = 2 * REC_RING + (WARPS - 2) * WORKER_RING;
template <bool DEC2, bool FIRST, bool IN8>
struct RecLay {
  static constexpr uint32_t SP    = IN8 ? 512u : 1024u;
  static constexpr uint32_t OFF_P = 0;
  static constexpr uint32_t OFF_S = SP;
  static constexpr uint32_t OFF_E = DEC2 ? SP : 2 * SP;
  static constexpr uint32_t BYTES = (OFF_E + (FIRST ? 0u : 1024u) + 127) / 128 * 128;
  static constexpr int      NST   = (int)(REC_RING / BYTES) > 24 ? 24 : (int)(REC_RING / BYTES);
};
struct Shared {
  volatile int doneA; // windows [0, doneA) consumed by the forward recursion, CKA[0 .. doneA-1] stored
  volatile int doneB; // windows [doneB, nw) consumed by the backward recursion, CKB[doneB-1 .. nw-1] stored (doneB-1 >= 0)
  int          cur_up, cur_dn;
  int          tile;
  LaneResult   xres[WARPS][32];
};
} // namespace ll

template <bool DEC2, bool FIRST, bool IN8>
__device__ __forceinline__ void ll_pass_tile(const TdecView& v, int tile, int pass_idx, uint8_t* smem, ll::Shared& sh, u4* cka, int debug)
{
  const long long t_start = clock64();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const LaneMap lm   = v.lanes[(size_t)tile * 32 + lane];
  const bool    held = lm.st0 != LANE_EMPTY;
  CbStatus*     stp  = v.status + (held ? lm.st0 : 0u);
  const bool act_lo = held && stp[0].active != 0, act_hi = held && stp[1].active != 0;
  if (__ballot_sync(0xFFFFFFFFu, act_lo || act_hi) == 0u) return; // every warp sees the same lanes: uniform over the thread block

  using RG = WarpRing<DEC2, FIRST, IN8>;
  using WL = ll::Lay<DEC2, IN8>;
  const TileDesc& td  = v.tiles[tile];
  const uint32_t  K   = td.K;
  const int       nw  = (int)(K / 8u);
  const bool      have_crc = (DEC2 ? td.crc_perm : td.crc_nat) != nullptr;
  const uint16_t* qpp_fwd  = td.qpp_fwd;
  // shared memory: two deep rings for the recursion warps (re-used as their worker stages afterwards), then the workers' stages
  uint8_t*        my_smem  = warp < 2 ? smem + warp * ll::REC_RING : smem + 2 * ll::REC_RING + (warp - 2) * ll::WORKER_RING;
  RG              rg(v, td, my_smem, lane, lm.hb0);
  uint8_t*        gCKA = reinterpret_cast<uint8_t*>(cka);

  if (threadIdx.x == 0) {
    sh.doneA  = 0;
    sh.doneB  = nw;
    sh.cur_up = nw / 2;
    sh.cur_dn = nw / 2 - 1;
  }
  __syncthreads();

  WinRegs  r;
  uint32_t M[8];
  if (warp < 2) {
    // ---- the two sequential recursions, each streamed through its own deep ring -----------------------------------------
    using RL = ll::RecLay<DEC2, FIRST, IN8>;
    constexpr int      NST  = RL::NST;
    constexpr uint32_t WB   = IN8 ? 512u : 1024u;
    const uint32_t     rb   = smem_u32(my_smem);
    const uint32_t     l16r = (uint32_t)lane * 16u, l4r = (uint32_t)lane * 4u;
    const int     dir = warp == 0 ? +1 : -1;
    int           nxt = warp == 0 ? 0 : nw - 1, left = nw;
    int           fill = 0, drain = 0;
    u4            qn  = {};
    M[0] = 0;
#pragma unroll
    for (int i = 1; i < 8; i++) M[i] = NEG_INF2;
    if (warp == 1) {
      const u4 pt = *reinterpret_cast<const u4*>(rg.tP + ((uint32_t)nw * WB + rg.l16));
      if (DEC2) {
        const u4 s2 = v.S2T[(size_t)tile * 32 + lane];
#pragma unroll
        for (int t = 2; t >= 0; t--) {
          const u4       pq[2] = {pt, pt};
          const uint32_t x = u4_get(s2, t), y = win_val<IN8>(pq, t);
          beta_step(M, x, y, add2(x, y));
        }
      } else {
        const u4 stl = *reinterpret_cast<const u4*>(rg.tS + ((uint32_t)nw * WB + rg.l16));
        beta_tail<IN8>(M, stl, pt);
      }
      rg.store_ck((uint32_t)(nw - 1), M); // CKB[nw-1] = beta_K
    }
    if (DEC2 && !FIRST) qn = ldg_q(qpp_fwd, (uint32_t)nxt);
    auto issue_next = [&]() {
      if (left > 0) {
        const uint32_t st = rb + (uint32_t)fill * RL::BYTES, uw = (uint32_t)nxt;
        cp16(st + RL::OFF_P + l16r, rg.tP + (uw * WB + l16r));
        if (!IN8) cp16(st + RL::OFF_P + 512u + l16r, rg.tP + (uw * WB + 512u + l16r));
        if (!DEC2) {
          cp16(st + RL::OFF_S + l16r, rg.tS + (uw * WB + l16r));
          if (!IN8) cp16(st + RL::OFF_S + 512u + l16r, rg.tS + (uw * WB + 512u + l16r));
        }
        if (!FIRST) {
          if (!DEC2) {
            cp16(st + RL::OFF_E + l16r, rg.tE + (uw * 1024u + l16r));
            cp16(st + RL::OFF_E + 512u + l16r, rg.tE + (uw * 1024u + 512u + l16r));
          } else {
            const u4 q = qn;
            if (left > 1) qn = ldg_q(qpp_fwd, (uint32_t)(nxt + dir));
#pragma unroll
            for (int t = 0; t < 8; t++) cp4(st + RL::OFF_E + (uint32_t)t * 128u + l4r, rg.tE + (win_pi(q, t) * 128u + l4r));
          }
        }
        fill = fill + 1 == NST ? 0 : fill + 1;
        nxt += dir;
        left--;
      }
      cp_commit();
    };
#pragma unroll 1
    for (int i = 0; i < NST; i++) issue_next();
    int w = warp == 0 ? 0 : nw - 1;
#pragma unroll 1
    for (int n = 0; n < nw; n++, w += dir) {
      cp_wait<NST - 1>();
      __syncwarp();
      const uint8_t* st = my_smem + (uint32_t)drain * RL::BYTES;
      drain             = drain + 1 == NST ? 0 : drain + 1;
      {
        u4       sv[2] = {}, pv[2] = {};
        uint32_t e[8] = {};
        pv[0] = *reinterpret_cast<const u4*>(st + RL::OFF_P + l16r);
        if (!IN8) pv[1] = *reinterpret_cast<const u4*>(st + RL::OFF_P + 512 + l16r);
        if (!DEC2) {
          sv[0] = *reinterpret_cast<const u4*>(st + RL::OFF_S + l16r);
          if (!IN8) sv[1] = *reinterpret_cast<const u4*>(st + RL::OFF_S + 512 + l16r);
        }
        if (!FIRST) {
#pragma unroll
          for (int t = 0; t < 8; t++) e[t] = *reinterpret_cast<const uint32_t*>(st + RL::OFF_E + t * 128 + l4r);
        }
        win_unpack<DEC2, FIRST, IN8>(r, sv, pv, e);
      }
      if (warp == 0) {
        // CKA[w] = alpha_{8w}, same [half][lane] rows as the tile's own checkpoint array
        *reinterpret_cast<u4*>(gCKA + ((uint32_t)w * 1024u + rg.l16))        = u4{M[0], M[1], M[2], M[3]};
        *reinterpret_cast<u4*>(gCKA + ((uint32_t)w * 1024u + 512u + rg.l16)) = u4{M[4], M[5], M[6], M[7]};
        alpha_window(M, r);
      } else {
        beta_window(M, r);
        if (w > 0) {
          rg.store_ck((uint32_t)(w - 1), M);
          normalise(M);
        }
      }
      __syncwarp(); // every lane has consumed the stage before it is refilled
      issue_next();
      if ((n & 3) == 3 || n == nw - 1) { // publish the progress every four windows: stores first, then the counter
        __threadfence_block();
        if (lane == 0) {
          if (warp == 0) sh.doneA = w + 1;
          else sh.doneB = w;
        }
      }
    }
    cp_wait<0>();
    __syncwarp();
    if (debug && lane == 0 && blockIdx.x == 0) printf("[ll] pass %d tile %d warp %d: recursion over %d windows took %lld cycles\n", pass_idx, tile, warp, nw, clock64() - t_start);
  }

  // ---- window work, handed out from the middle of the trellis outwards ---------------------------------------------------
  const uint32_t sbase = smem_u32(my_smem);
  const uint32_t l16 = (uint32_t)lane * 16u, l4 = (uint32_t)lane * 4u;
  const uint8_t* gCRC = reinterpret_cast<const uint8_t*>(DEC2 ? td.crc_perm : td.crc_nat);
  LaneResult     res  = {0u, 0u};
  auto grab = [&]() -> int {
    int w = -1;
    if (lane == 0) {
      if (warp & 1) {
        w = atomicAdd(&sh.cur_up, 1);
        if (w >= nw) {
          w = atomicSub(&sh.cur_dn, 1);
          if (w < 0) w = -1;
        }
      } else {
        w = atomicSub(&sh.cur_dn, 1);
        if (w < 0) {
          w = atomicAdd(&sh.cur_up, 1);
          if (w >= nw) w = -1;
        }
      }
    }
    return __shfl_sync(0xFFFFFFFFu, w, 0);
  };
  // waits until both recursions are past window w, then enqueues everything the window needs into stage `s`
  auto fetch = [&](int w, int s) {
    while (!(sh.doneA > w && sh.doneB <= w)) __nanosleep(200);
    __threadfence_block(); // the checkpoints were written (and fenced) by the recursion warps before they raised the counters
    __syncwarp();
    const uint32_t     st = sbase + (uint32_t)s * WL::BYTES;
    constexpr uint32_t WB = IN8 ? 512u : 1024u;
    const uint32_t     uw = (uint32_t)w;
    cp16(st + WL::OFF_P + l16, rg.tP + (uw * WB + l16));
    if (!IN8) cp16(st + WL::OFF_P + 512u + l16, rg.tP + (uw * WB + 512u + l16));
    if (!DEC2) {
      cp16(st + WL::OFF_S + l16, rg.tS + (uw * WB + l16));
      if (!IN8) cp16(st + WL::OFF_S + 512u + l16, rg.tS + (uw * WB + 512u + l16));
    }
    u4 q = {};
    if (DEC2) q = ldg_q(qpp_fwd, uw);
    if (!FIRST) {
      if (!DEC2) {
        cp16(st + WL::OFF_E + l16, rg.tE + (uw * 1024u + l16));
        cp16(st + WL::OFF_E + 512u + l16, rg.tE + (uw * 1024u + 512u + l16));
      } else {
#pragma unroll
        for (int t = 0; t < 8; t++) cp4(st + WL::OFF_E + (uint32_t)t * 128u + l4, rg.tE + (win_pi(q, t) * 128u + l4));
      }
    }
    cp16(st + WL::OFF_CKA + l16, gCKA + (uw * 1024u + l16));
    cp16(st + WL::OFF_CKA + 512u + l16, gCKA + (uw * 1024u + 512u + l16));
    cp16(st + WL::OFF_CKB + l16, rg.tCK + (uw * 1024u + l16));
    cp16(st + WL::OFF_CKB + 512u + l16, rg.tCK + (uw * 1024u + 512u + l16));
    if (gCRC != nullptr && l16 < 64u) cp16(st + WL::OFF_CRC + l16, gCRC + (uw * 64u + l16));
    if (DEC2 && lane == 0) *reinterpret_cast<u4*>(my_smem + (uint32_t)s * WL::BYTES + WL::OFF_QPP) = q;
    cp_commit();
  };
  // Warps share an issue port four by four (warp id mod 4).  The recursions are the critical path of the pass, so the worker
  // warps that sit on the recursion warps' ports stay out of their way until the recursion there is finished; the workers on
  // the other two ports start as soon as the fronts have crossed.
  if (warp >= 2 && (warp & 3) == 0) {
    while (sh.doneA < nw) __nanosleep(400);
  } else if (warp >= 2 && (warp & 3) == 1) {
    while (sh.doneB > 0) __nanosleep(400);
  }
  int cur = grab(), s = 0;
  if (cur >= 0) fetch(cur, 0);
#pragma unroll 1
  while (cur >= 0) {
    const int nxt = grab();
    if (nxt >= 0) fetch(nxt, s ^ 1); // the next window's data travels while this one is worked on
    if (nxt >= 0) cp_wait<1>();
    else cp_wait<0>();
    __syncwarp();
    const uint8_t* st = my_smem + (uint32_t)s * WL::BYTES;
    {
      u4       sv[2] = {}, pv[2] = {};
      uint32_t e[8] = {};
      pv[0] = *reinterpret_cast<const u4*>(st + WL::OFF_P + l16);
      if (!IN8) pv[1] = *reinterpret_cast<const u4*>(st + WL::OFF_P + 512 + l16);
      if (!DEC2) {
        sv[0] = *reinterpret_cast<const u4*>(st + WL::OFF_S + l16);
        if (!IN8) sv[1] = *reinterpret_cast<const u4*>(st + WL::OFF_S + 512 + l16);
      }
      if (!FIRST) {
#pragma unroll
        for (int t = 0; t < 8; t++) e[t] = *reinterpret_cast<const uint32_t*>(st + WL::OFF_E + t * 128 + l4);
      }
      win_unpack<DEC2, FIRST, IN8>(r, sv, pv, e);
    }
    uint32_t c[8];
    {
      const u4 a0 = *reinterpret_cast<const u4*>(st + WL::OFF_CKA + l16), a1 = *reinterpret_cast<const u4*>(st + WL::OFF_CKA + 512 + l16);
      const u4 b0 = *reinterpret_cast<const u4*>(st + WL::OFF_CKB + l16), b1 = *reinterpret_cast<const u4*>(st + WL::OFF_CKB + 512 + l16);
      M[0] = a0.x; M[1] = a0.y; M[2] = a0.z; M[3] = a0.w; M[4] = a1.x; M[5] = a1.y; M[6] = a1.z; M[7] = a1.w;
      c[0] = b0.x; c[1] = b0.y; c[2] = b0.z; c[3] = b0.w; c[4] = b1.x; c[5] = b1.y; c[6] = b1.z; c[7] = b1.w;
    }
    u4 q = {};
    if (DEC2) q = *reinterpret_cast<const u4*>(st + WL::OFF_QPP);
    WinOut         o;
    const uint32_t uw = (uint32_t)cur;
    fwd_window(M, c, 8u * uw + 8u < K, r, have_crc ? reinterpret_cast<const CrcPow*>(st + WL::OFF_CRC) : nullptr, res, o,
               [&](int t, uint32_t L) {
                 uint32_t ein = 0u;
                 if (!FIRST) {
                   if (DEC2) ein = r.xs[t];
                   else asm volatile("ld.shared.u32 %0, [%1];" : "=r"(ein) : "r"(smem_u32(st) + WL::OFF_E + (uint32_t)t * 128u + l4));
                 }
                 *rg.e_out(uw, q, t) = sub2(L, ein);
               });
    rg.store_hb(uw, o.bits, act_lo, act_hi);
    __syncwarp(); // every lane is done with this stage before it is refilled
    cur = nxt;
    s ^= 1;
  }
  sh.xres[warp][lane] = res;
  if (debug && lane == 0 && blockIdx.x == 0) printf("[ll] pass %d tile %d warp %d: done after %lld cycles\n", pass_idx, tile, warp, clock64() - t_start);
  __syncthreads();
  if (warp == 0) {
    LaneResult tot = {0u, 0u};
#pragma unroll
    for (int i = 0; i < ll::WARPS; i++) {
      tot.crc_lo16x2 ^= sh.xres[i][lane].crc_lo16x2;
      tot.crc_hi8x2 ^= sh.xres[i][lane].crc_hi8x2;
    }
    finish_pass(v, have_crc, stp, stp[0], stp[1], act_lo, act_hi, tot, pass_idx);
  }
  __syncthreads(); // the shared control block is reused by the next tile
}

// Thread blocks take tiles from a queue (TDEC_CTL_LL_QUEUE): a batch of mixed lengths is ordered longest first and the
// running tiles of an early-stop decode sit anywhere.  only_if_flagged: run only when the device decided that the pass
// belongs to this kernel (TDEC_CTL_USE_LL); the throughput kernel, launched beside it, then returns at once.
template <bool DEC2, bool FIRST>
__global__ void __launch_bounds__(ll::WARPS * 32, 1) tdec_siso_pass_ll_kernel(TdecView v, int pass_idx, int only_if_flagged)
{
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ ll::Shared sh;
  if ((only_if_flagged & 1) && v.ctl[TDEC_CTL_USE_LL] == 0u) return;
  u4* cka = v.ll_ck + (size_t)blockIdx.x * v.ll_ck_slot;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) sh.tile = (int)atomicAdd(v.ctl + TDEC_CTL_LL_QUEUE, 1u);
    __syncthreads();
    const int tile = sh.tile;
    if (tile >= v.ntiles) return;
    if (v.fmt[tile] == 0u || v.tiles[tile].S == nullptr) {
      ll_pass_tile<DEC2, FIRST, true>(v, tile, pass_idx, smem, sh, cka, only_if_flagged & 2);
    } else {
      ll_pass_tile<DEC2, FIRST, false>(v, tile, pass_idx, smem, sh, cka, only_if_flagged & 2);
    }
  }
}

static void siso_set_attributes()
{
  const size_t smem = 2 * ring::RING_BYTES;
  static std::atomic<uint64_t> attr_done{0}; // function attributes are per device
  if (once_per_device(attr_done)) {
    cudaFuncSetAttribute(tdec_siso_pass_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(tdec_siso_pass_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(tdec_siso_pass_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // 7 CTAs per SM (one wave for the 1024 tiles of a 65,536-block batch) need the full shared-memory carve-out
    cudaFuncSetAttribute(tdec_siso_pass_kernel<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(tdec_siso_pass_kernel<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(tdec_siso_pass_kernel<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    const int ll_smem = (int)ll::SMEM_BYTES;
    cudaFuncSetAttribute(tdec_siso_pass_ll_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ll_smem);
    cudaFuncSetAttribute(tdec_siso_pass_ll_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ll_smem);
    cudaFuncSetAttribute(tdec_siso_pass_ll_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ll_smem);
  }
}

int siso_resident_tiles_per_sm()
{
  siso_set_attributes();
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, tdec_siso_pass_kernel<false, false>, 64, 2 * ring::RING_BYTES) != cudaSuccess) {
    return -1;
  }
  return n;
}

// mode: SISO_THROUGHPUT (two warps per tile, every tile its own thread block), SISO_LOW_LATENCY (16 warps per tile, one
// thread block per SM) or SISO_AUTO (both are launched and the device's TDEC_CTL_USE_LL flag says which one works)
void launch_siso_pass(const TdecView& v, int pass_idx, int mode, int sm_count, cudaStream_t stream)
{
  siso_set_attributes();
  if (mode != SISO_LOW_LATENCY) {
    const size_t smem = 2 * ring::RING_BYTES;
    dim3         grid((unsigned)v.ntiles), block(64);
    const int    unless = mode == SISO_AUTO;
    if (pass_idx == 0) {
      tdec_siso_pass_kernel<false, true><<<grid, block, smem, stream>>>(v, pass_idx, unless);
    } else if (pass_idx & 1) {
      tdec_siso_pass_kernel<true, false><<<grid, block, smem, stream>>>(v, pass_idx, unless);
    } else {
      tdec_siso_pass_kernel<false, false><<<grid, block, smem, stream>>>(v, pass_idx, unless);
    }
  }
  if (mode != SISO_THROUGHPUT) {
    const size_t smem = ll::SMEM_BYTES;
    dim3         grid((unsigned)std::min(v.ntiles, sm_count)), block(ll::WARPS * 32);
    static const bool ll_debug = getenv("SRSLTE_B200_TDEC_LL_DEBUG") != nullptr; // cycle stamps of the first thread block (printf)
    const int    only = (mode == SISO_AUTO ? 1 : 0) | (ll_debug ? 2 : 0);
    cudaMemsetAsync(v.ctl + TDEC_CTL_LL_QUEUE, 0, sizeof(uint32_t), stream);
    if (pass_idx == 0) {
      tdec_siso_pass_ll_kernel<false, true><<<grid, block, smem, stream>>>(v, pass_idx, only);
    } else if (pass_idx & 1) {
      tdec_siso_pass_ll_kernel<true, false><<<grid, block, smem, stream>>>(v, pass_idx, only);
    } else {
      tdec_siso_pass_ll_kernel<false, false><<<grid, block, smem, stream>>>(v, pass_idx, only);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// natural -> tiled.  Block = (tile, chunk of 32 trellis rows).  Phase 1 stages the 64 blocks' 96 contiguous int16
// each through shared memory with 8-byte loads (a block's natural vector is only 8-byte aligned: (3K+12)*2 bytes);
// phase 2 emits one uint4 per thread and stream.
//   tdec_load8_kernel   always runs: writes the int8 arrays (8 rows x 2 blocks per uint4) and raises fmt[tile] when a
//                       value of the tile does not fit int8; also arms the per-block state.
//   tdec_load16_kernel  runs after it and fills the int16 arrays (4 rows x 2 blocks per uint4) of the raised tiles only.
constexpr int    LOAD_ROWS = 64;                 // trellis rows per CTA: 384 contiguous bytes of each block's natural vector
constexpr int    RAW_PITCH = LOAD_ROWS * 6 + 8;  // staging buffer: block c's chunk as loaded, at byte c * RAW_PITCH
constexpr size_t LOAD_SMEM = (size_t)TDEC_TILE_CB * RAW_PITCH;

__device__ __forceinline__ bool fits8(int16_t a)
{
  return (int16_t)(int8_t)a == a;
}

// A value of the tile does not fit int8: the tile goes to the int16 arrays -- or, when the workspace was carved without them
// (int16 on demand), the batch is flagged so that the host repeats it with them.
__device__ __forceinline__ void raise_int16(const TdecView& v, const TileDesc& td, int tile)
{
  atomicOr(v.fmt + tile, 1u);
  if (td.S == nullptr) atomicOr(v.err, 1u);
}

// Source of block c (0..63) of a tile: `offsets` (optional) holds the int16 offset of every block's vector inside llr,
// indexed by the block's position in the batch (soft buffers scattered in a HARQ pool); without it the vectors of a
// tile are contiguous from llr + td.llr_off.  nullptr past the tile's last block.
__device__ __forceinline__ const int16_t* block_src(const TileDesc& td, const int16_t* __restrict__ llr, const uint64_t* __restrict__ offsets, int c)
{
  if ((uint32_t)c >= td.nblk) return nullptr;
  return llr + (offsets ? offsets[td.cb0 + (uint32_t)c] : td.llr_off + (uint64_t)c * (3ull * td.K + 12ull));
}

// ALIGNED8: every vector starts on an 8-byte boundary.  Returns (to all threads) whether every staged value fits int8.
template <bool ALIGNED8>
__device__ __forceinline__ bool load_stage_chunk(uint8_t* sm, const TileDesc& td, const int16_t* __restrict__ llr,
                                                 const uint64_t* __restrict__ offsets, int k0, int rows)
{
  const int nvec = rows * 3 / 4; // 8-byte vectors per block in this chunk (<= 48)
  // warp w stages blocks w, w+8, ...; lane q takes the q-th 8-byte piece of the block's chunk (no divisions).  Two
  // blocks per round keep 6 independent 8-byte loads per thread in flight.
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  uint32_t  acc  = 0; // v ^ (v << 1) has bits 15..8 clear <=> bits 15..7 of v are equal <=> v fits int8 (per half)
  constexpr int NB = 2;
  for (int c0 = wid; c0 < TDEC_TILE_CB; c0 += NB * nwarp) {
    constexpr int NH = (LOAD_ROWS * 3 / 4 + 31) / 32; // 8-byte pieces per lane and block
    uint2 val[NB][NH];
#pragma unroll
    for (int u = 0; u < NB; u++) {
      const int      c    = c0 + u * nwarp;
      const int16_t* base = c < TDEC_TILE_CB ? block_src(td, llr, offsets, c) : nullptr;
#pragma unroll
      for (int h = 0; h < NH; h++) {
        const int q = lane + 32 * h;
        val[u][h]   = make_uint2(0u, 0u);
        if (base != nullptr && q < nvec) {
          const int16_t* src = base + 3 * (size_t)k0 + 4 * (size_t)q;
          if (ALIGNED8) {
            val[u][h] = __ldcs(reinterpret_cast<const uint2*>(src));
          } else {
            val[u][h].x = (uint32_t)(uint16_t)src[0] | ((uint32_t)(uint16_t)src[1] << 16);
            val[u][h].y = (uint32_t)(uint16_t)src[2] | ((uint32_t)(uint16_t)src[3] << 16);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < NB; u++) {
      const int c = c0 + u * nwarp;
#pragma unroll
      for (int h = 0; h < NH; h++) {
        const int q = lane + 32 * h;
        if (c < TDEC_TILE_CB && q < nvec) {
          acc |= (val[u][h].x ^ (val[u][h].x << 1)) | (val[u][h].y ^ (val[u][h].y << 1));
          *reinterpret_cast<uint2*>(sm + c * RAW_PITCH + 8 * q) = val[u][h];
        }
      }
    }
  }
  const bool bad = (acc & 0xFF00FF00u) != 0u;
  return __syncthreads_or(bad) == 0;
}

// The 12 tail values of a tile's blocks (row K/8 of S8/P08/P18 and S2T), the per-block state of a fresh decode
// (srsran_tdec_new_cb, turbodecoder.c:510-525) and the lane map (every pair starts at home).  One warp.
__device__ __forceinline__ void load_tail_and_arm(const TdecView& v, const TileDesc& td, int tile, const int16_t* __restrict__ llr,
                                                  const uint64_t* __restrict__ offsets, int lane)
{
  const int K = (int)td.K;
  uint32_t  w16[4][4];
  bool      bad = false;
  const int16_t* src0 = block_src(td, llr, offsets, 2 * lane);
  const int16_t* src1 = block_src(td, llr, offsets, 2 * lane + 1);
#pragma unroll
  for (int s_ = 0; s_ < 4; s_++) {
#pragma unroll
    for (int t = 0; t < 4; t++) {
      int16_t a0 = 0, b0 = 0;
      if (src0) a0 = natural_pick(src0, K, s_, K + t);
      if (src1) b0 = natural_pick(src1, K, s_, K + t);
      w16[s_][t] = pack2(a0, b0);
      if (s_ < 3) bad |= !fits8(a0) || !fits8(b0); // S2T stays int16
    }
  }
  if (__any_sync(0xFFFFFFFFu, bad) && lane == 0) raise_int16(v, td, tile);
#pragma unroll
  for (int s_ = 0; s_ < 3; s_++) {
    uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int t = 0; t < 4; t++) {
      const uint32_t lo = w16[s_][t] & 0xFFu, hi = (w16[s_][t] >> 16) & 0xFFu;
      w[t >> 1] |= (lo | (hi << 8)) << (16 * (t & 1));
    }
    u4* dst = s_ == 0 ? td.S8 : (s_ == 1 ? td.P08 : td.P18);
    dst[row8(K / 8, lane)] = u4{w[0], w[1], w[2], w[3]};
  }
  v.S2T[(size_t)tile * 32 + lane] = u4{w16[3][0], w16[3][1], w16[3][2], w16[3][3]};
  const LaneMap home  = lane_home(td, tile, lane);
  v.lanes[(size_t)tile * 32 + lane] = home;
  v.status[home.st0]     = CbStatus{(uint8_t)(src0 != nullptr), 0, 0, 0};
  v.status[home.st0 + 1] = CbStatus{(uint8_t)(src1 != nullptr), 0, 0, 0};
}

template <bool ALIGNED8>
__global__ void __launch_bounds__(256)
    tdec_load8_kernel(TdecView v, const int16_t* __restrict__ llr, const uint64_t* __restrict__ offsets)
{
  extern __shared__ __align__(16) uint8_t sm[];
  const int       tile  = blockIdx.y;
  const int       chunk = blockIdx.x;
  const TileDesc& td    = v.tiles[tile];
  const int       K     = (int)td.K;
  const int       k0    = chunk * LOAD_ROWS;
  const int       tid   = threadIdx.x;
  const int       nchunk = (K + LOAD_ROWS - 1) / LOAD_ROWS;
  if (chunk > nchunk) return; // a shorter tile of a mixed batch

  if (k0 < K) {
    const int  rows = min(LOAD_ROWS, K - k0); // multiple of 8
    const bool ok   = load_stage_chunk<ALIGNED8>(sm, td, llr, offsets, k0, rows);
    if (!ok && tid == 0) raise_int16(v, td, tile);
    // warp = window of the chunk; a lane reads the 48 bytes of its two blocks, pairs and packs them with byte permutes
    // and writes its 16 bytes of each int8 tile row
    const int lane = tid & 31;
    for (int w8 = tid >> 5; w8 < rows / 8; w8 += 8) {
      uint32_t a[12], b[12];
#pragma unroll
      for (int j = 0; j < 6; j++) {
        const uint2 x = *reinterpret_cast<const uint2*>(sm + (2 * lane) * RAW_PITCH + w8 * 48 + 8 * j);
        const uint2 y = *reinterpret_cast<const uint2*>(sm + (2 * lane + 1) * RAW_PITCH + w8 * 48 + 8 * j);
        a[2 * j] = x.x; a[2 * j + 1] = x.y; b[2 * j] = y.x; b[2 * j + 1] = y.y;
      }
#pragma unroll
      for (int s = 0; s < 3; s++) {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          // rows 2q, 2q+1 of stream s: value index i = 3*row + s inside the window's 24 values of a block
          const int      i0 = 3 * (2 * q) + s, i1 = 3 * (2 * q + 1) + s;
          const uint32_t p0 = __byte_perm(a[i0 >> 1], b[i0 >> 1], (i0 & 1) ? 0x7632 : 0x5410); // (a, b) as an int16 pair
          const uint32_t p1 = __byte_perm(a[i1 >> 1], b[i1 >> 1], (i1 & 1) ? 0x7632 : 0x5410);
          w[q]              = __byte_perm(p0, p1, 0x6420);                                     // their low bytes
        }
        u4* dst = s == 0 ? td.S8 : (s == 1 ? td.P08 : td.P18);
        dst[row8(k0 / 8 + w8, lane)] = u4{w[0], w[1], w[2], w[3]};
      }
    }
  } else if (tid < 32) {
    // the chunk past the payload carries the 12 tail values
    load_tail_and_arm(v, td, tile, llr, offsets, tid);
  }
}

template <bool ALIGNED8>
__global__ void __launch_bounds__(256)
    tdec_load16_kernel(TdecView v, const int16_t* __restrict__ llr, const uint64_t* __restrict__ offsets, int max_chunks)
{
  extern __shared__ __align__(16) uint8_t sm[];
  const int tid = threadIdx.x;
  // few CTAs walking all (tile, chunk) items: in the common case no tile is raised and this costs microseconds
  for (int item = blockIdx.x; item < v.ntiles * max_chunks; item += gridDim.x) {
  const int tile = item / max_chunks, chunk = item % max_chunks;
  if (v.fmt[tile] == 0u) continue; // the tile lives in the int8 arrays
  const TileDesc& td = v.tiles[tile];
  if (td.S == nullptr) continue;   // no int16 arrays in this workspace: the batch is flagged (v.err) and repeated by the host
  const int       K  = (int)td.K;
  const int       k0 = chunk * LOAD_ROWS;
  if (chunk > (K + LOAD_ROWS - 1) / LOAD_ROWS) continue;
  __syncthreads();                 // the previous item's readers of sm are done

  if (k0 < K) {
    const int rows = min(LOAD_ROWS, K - k0); // multiple of 8
    load_stage_chunk<ALIGNED8>(sm, td, llr, offsets, k0, rows);
    const int lane = tid & 31;
    for (int r4 = tid >> 5; r4 < rows / 4; r4 += 8) {
      uint32_t w[3][4];
#pragma unroll
      for (int t = 0; t < 4; t++) {
#pragma unroll
        for (int s = 0; s < 3; s++) {
          const int o = 3 * (4 * r4 + t) + s;
          w[s][t]     = pack2(reinterpret_cast<const int16_t*>(sm + (2 * lane) * RAW_PITCH)[o],
                              reinterpret_cast<const int16_t*>(sm + (2 * lane + 1) * RAW_PITCH)[o]);
        }
      }
      const uint32_t row = vec_row(k0 / 4 + r4, lane);
      td.S[row]          = u4{w[0][0], w[0][1], w[0][2], w[0][3]};
      td.P0[row]         = u4{w[1][0], w[1][1], w[1][2], w[1][3]};
      td.P1[row]         = u4{w[2][0], w[2][1], w[2][2], w[2][3]};
    }
  } else if (tid < 32) {
    const int lane = tid;
    uint32_t  w[3][4];
    const int16_t* src0 = block_src(td, llr, offsets, 2 * lane);
    const int16_t* src1 = block_src(td, llr, offsets, 2 * lane + 1);
#pragma unroll
    for (int s = 0; s < 3; s++) {
#pragma unroll
      for (int t = 0; t < 4; t++) {
        int16_t a = 0, b = 0;
        if (src0) a = natural_pick(src0, K, s, K + t);
        if (src1) b = natural_pick(src1, K, s, K + t);
        w[s][t] = pack2(a, b);
      }
    }
    const uint32_t row = vec_row(K / 4, lane);
    td.S[row]          = u4{w[0][0], w[0][1], w[0][2], w[0][3]};
    td.P0[row]         = u4{w[1][0], w[1][1], w[1][2], w[1][3]};
    td.P1[row]         = u4{w[2][0], w[2][1], w[2][2], w[2][3]};
  }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Streaming form of tdec_load8_kernel for 8-byte aligned block vectors (the common case): a CTA owns (tile, segment of
// consecutive 64-row chunks) and double-buffers the chunks with cp.async, so the global-memory latency of chunk c+1 is
// hidden behind the packing of chunk c (the one-chunk-per-CTA kernel above spends 2/3 of its time waiting on its loads:
// profiles/README.md).  Staging layout: block c of the tile at row (c & 1) * 32 + (c >> 1), 400 bytes per row, so that a lane
// finds its two blocks at rows lane and lane + 32 and the 16-byte reads of a quarter warp fall into 32 different banks.
constexpr int    LS_PITCH  = 400;                               // bytes per staged block row (384 used)
constexpr size_t LS_BUF    = (size_t)TDEC_TILE_CB * LS_PITCH;   // one chunk of a tile
constexpr size_t LS_SMEM   = 2 * LS_BUF + TDEC_TILE_CB * sizeof(uint64_t);
constexpr int    LS_CHUNKS = 12;                                // chunks per CTA (K=6144: 96 chunks -> 8 segments per tile)

__device__ __forceinline__ void cp_async8_raw(uint32_t dst, const void* src)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(__cvta_generic_to_global(src)) : "memory");
}

__global__ void __launch_bounds__(256)
    tdec_load8_stream_kernel(TdecView v, const int16_t* __restrict__ llr, const uint64_t* __restrict__ offsets)
{
  extern __shared__ __align__(16) uint8_t sm[];
  const uint8_t** sbase = reinterpret_cast<const uint8_t**>(sm + 2 * LS_BUF); // byte address of every block's vector
  const int       tile = blockIdx.y, seg = blockIdx.x;
  const TileDesc& td   = v.tiles[tile];
  const int       K    = (int)td.K;
  const int       tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int       nchunk = (K + LOAD_ROWS - 1) / LOAD_ROWS;
  const int       nseg   = (nchunk + LS_CHUNKS - 1) / LS_CHUNKS;
  if (seg >= nseg) return; // a shorter tile of a mixed batch
  const int       c_lo = seg * LS_CHUNKS, c_hi = min(nchunk, c_lo + LS_CHUNKS);
  if (tid < TDEC_TILE_CB) sbase[tid] = reinterpret_cast<const uint8_t*>(block_src(td, llr, offsets, tid));
  // rows of blocks past the end of the batch stay zero in both buffers
  for (int i = tid; i < (int)(2 * LS_BUF / 16); i += 256) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  const uint32_t sm_u32 = smem_u32(sm);

  // warp w stages blocks w, w+8, ..., w+56; lane q takes the 8-byte pieces q and q+32 of the block's chunk (48 per full chunk)
  auto prefetch = [&](int chunk, int buf) {
    const int k0   = chunk * LOAD_ROWS;
    const int nvec = min(LOAD_ROWS, K - k0) * 3 / 4;
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int      c   = wid + 8 * u;
      const uint8_t* src = sbase[c];
      if (src != nullptr) {
        const uint32_t dst = sm_u32 + (uint32_t)buf * (uint32_t)LS_BUF + (uint32_t)(((c & 1) * 32 + (c >> 1)) * LS_PITCH);
        src += 6 * (size_t)k0;
        if (lane < nvec) cp_async8_raw(dst + 8u * lane, src + 8 * lane);
        if (lane + 32 < nvec) cp_async8_raw(dst + 8u * (lane + 32), src + 8 * (lane + 32));
      }
    }
    cp_commit();
  };

  uint32_t acc = 0; // v ^ (v << 1) has bits 15..8 clear <=> v fits int8 (per int16 half)
  if (c_lo < c_hi) prefetch(c_lo, 0);
  for (int chunk = c_lo; chunk < c_hi; chunk++) {
    const int buf = (chunk - c_lo) & 1;
    cp_wait<0>();
    __syncthreads(); // this chunk has landed; everybody is done with the other buffer
    if (chunk + 1 < c_hi) prefetch(chunk + 1, buf ^ 1);
    const int      k0   = chunk * LOAD_ROWS;
    const int      rows = min(LOAD_ROWS, K - k0); // multiple of 8
    const uint8_t* bsm  = sm + (size_t)buf * LS_BUF;
    // warp = window of the chunk; a lane reads the 48 bytes of its two blocks, pairs and packs them with byte permutes and
    // writes its 16 bytes of each int8 tile row
    for (int w8 = wid; w8 < rows / 8; w8 += 8) {
      uint32_t a[12], b[12];
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const uint4 x = *reinterpret_cast<const uint4*>(bsm + lane * LS_PITCH + w8 * 48 + 16 * j);
        const uint4 y = *reinterpret_cast<const uint4*>(bsm + (lane + 32) * LS_PITCH + w8 * 48 + 16 * j);
        a[4 * j] = x.x; a[4 * j + 1] = x.y; a[4 * j + 2] = x.z; a[4 * j + 3] = x.w;
        b[4 * j] = y.x; b[4 * j + 1] = y.y; b[4 * j + 2] = y.z; b[4 * j + 3] = y.w;
      }
#pragma unroll
      for (int j = 0; j < 12; j++) acc |= (a[j] ^ (a[j] << 1)) | (b[j] ^ (b[j] << 1));
#pragma unroll
      for (int s_ = 0; s_ < 3; s_++) {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          // rows 2q, 2q+1 of stream s_: value index i = 3*row + s_ inside the window's 24 values of a block
          const int      i0 = 3 * (2 * q) + s_, i1 = 3 * (2 * q + 1) + s_;
          const uint32_t p0 = __byte_perm(a[i0 >> 1], b[i0 >> 1], (i0 & 1) ? 0x7632 : 0x5410); // (a, b) as an int16 pair
          const uint32_t p1 = __byte_perm(a[i1 >> 1], b[i1 >> 1], (i1 & 1) ? 0x7632 : 0x5410);
          w[q]              = __byte_perm(p0, p1, 0x6420);                                     // their low bytes
        }
        u4* dst = s_ == 0 ? td.S8 : (s_ == 1 ? td.P08 : td.P18);
        dst[row8(k0 / 8 + w8, lane)] = u4{w[0], w[1], w[2], w[3]};
      }
    }
  }
  if (__syncthreads_or((acc & 0xFF00FF00u) != 0u) && tid == 0) raise_int16(v, td, tile);

  // the tile's last segment also writes the tail row and arms the block state
  if (seg == nseg - 1 && tid < 32) load_tail_and_arm(v, td, tile, llr, offsets, lane);
}


void launch_load_natural(const TdecView& v,
                         int             max_K,
                         const int16_t*  llr_dev,
                         const uint64_t* offsets_dev,
                         bool            aligned8,
                         cudaStream_t    stream,
                         bool            int8_tiles_done)
{
  const int chunks = (max_K + LOAD_ROWS - 1) / LOAD_ROWS + 1; // +1: the tail chunk
  dim3      grid((unsigned)chunks, (unsigned)v.ntiles), block(256);
  const long items = (long)chunks * v.ntiles;
  dim3       grid16((unsigned)(items < 148 * 8 ? items : 148 * 8));
  static std::atomic<uint64_t> attr_done{0};
  if (once_per_device(attr_done)) { // 8 CTAs of 25 KB per SM
    cudaFuncSetAttribute(tdec_load8_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(tdec_load8_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(tdec_load8_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LOAD_SMEM);
    cudaFuncSetAttribute(tdec_load8_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LOAD_SMEM);
    cudaFuncSetAttribute(tdec_load16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LOAD_SMEM);
    cudaFuncSetAttribute(tdec_load16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LOAD_SMEM);
    cudaFuncSetAttribute(tdec_load8_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LS_SMEM);
    cudaFuncSetAttribute(tdec_load8_stream_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
  if (int8_tiles_done) { // the fused de-matching kernel wrote the int8 tiles and raised the format flags: only the raised tiles are left
    if (aligned8) tdec_load16_kernel<true><<<grid16, block, LOAD_SMEM, stream>>>(v, llr_dev, offsets_dev, chunks);
    else tdec_load16_kernel<false><<<grid16, block, LOAD_SMEM, stream>>>(v, llr_dev, offsets_dev, chunks);
    return;
  }
  cudaMemsetAsync(v.fmt, 0, (size_t)v.ntiles * sizeof(uint32_t), stream);
  cudaMemsetAsync(v.err, 0, sizeof(uint32_t), stream);
  if (aligned8) {
    const int nchunk = (max_K + LOAD_ROWS - 1) / LOAD_ROWS;
    dim3      gs((unsigned)((nchunk + LS_CHUNKS - 1) / LS_CHUNKS), (unsigned)v.ntiles);
    tdec_load8_stream_kernel<<<gs, block, LS_SMEM, stream>>>(v, llr_dev, offsets_dev);
    tdec_load16_kernel<true><<<grid16, block, LOAD_SMEM, stream>>>(v, llr_dev, offsets_dev, chunks);
  } else {
    tdec_load8_kernel<false><<<grid, block, LOAD_SMEM, stream>>>(v, llr_dev, offsets_dev);
    tdec_load16_kernel<false><<<grid16, block, LOAD_SMEM, stream>>>(v, llr_dev, offsets_dev, chunks);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// HB -> bytes.  One CTA per tile: the tile's decisions (K/8 x 64 bytes) are staged in shared memory, so the gather
// through the inverse interleaver that a block ending on a DEC2 pass needs (eight 1-bit lookups per output byte)
// never leaves the SM; the 64 x K/8 output bytes are assembled in shared memory too and leave as one contiguous run.
// Tiles are HOME tiles here: status and HB never move when the lanes are re-packed between passes.
__global__ void __launch_bounds__(256) tdec_decide_kernel(TdecView v,
                                                          uint8_t* __restrict__ out,
                                                          uint8_t* __restrict__ crc_ok,
                                                          uint8_t* __restrict__ npass,
                                                          uint8_t* __restrict__ npass_run)
{
  extern __shared__ __align__(16) uint8_t dsm[];
  const int       tile  = blockIdx.x;
  const TileDesc& td    = v.tiles[tile];
  const int       K     = (int)td.K;
  const int       nb    = K / 8; // bytes per block = windows per block
  const int       pitch = nb + 4;
  uint16_t*       hb    = reinterpret_cast<uint16_t*>(dsm);           // [nb][32]
  uint8_t*        so    = dsm + (((size_t)nb * 64 + 15) / 16) * 16;   // [64][pitch]
  const int       tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint4*    src = reinterpret_cast<const uint4*>(v.HB + (size_t)td.hb_row0 * 32);
  const uint16_t* __restrict__ qpp_rev = td.qpp_rev;
  for (int i = tid; i < nb * 4; i += 256) reinterpret_cast<uint4*>(hb)[i] = src[i];
  __syncthreads();

  // a warp takes 4 output bytes (32 decisions) at a time; lane = lane of the tile = the block pair (2 lane, 2 lane + 1) whose
  // decisions share a 16-bit HB word (low byte = even block), so one shared-memory read serves two blocks
  const CbStatus st_lo = v.status[(size_t)tile * TDEC_TILE_CB + 2 * lane], st_hi = v.status[(size_t)tile * TDEC_TILE_CB + 2 * lane + 1];
  // last pass was DEC2: HB is in its visiting order
  const bool     perm_lo = st_lo.npass_run > 0 && ((st_lo.npass_run - 1) & 1), perm_hi = st_hi.npass_run > 0 && ((st_hi.npass_run - 1) & 1);
  const bool     any_perm = __any_sync(0xFFFFFFFFu, perm_lo || perm_hi);
  const uint32_t keep_nat = (perm_lo ? 0u : 0x00FFu) | (perm_hi ? 0u : 0xFF00u); // halves that take the natural-order word
  for (int j0 = wid * 4; j0 < nb; j0 += 32) {
    const int      nj  = min(4, nb - j0);
    const uint32_t rev = (8 * j0 + lane < K) ? qpp_rev[8 * j0 + lane] : 0u; // visiting index of bit 8*j0+lane
    for (int j = 0; j < nj; j++) {
      const uint32_t nat = hb[(j0 + j) * 32 + lane];
      uint32_t       w2  = nat;
      if (any_perm) {
        uint32_t acc = 0; // bit 0 / bit 8: the low / high block's decision, shifted in MSB first
#pragma unroll
        for (int t = 0; t < 8; t++) {
          const uint32_t i = __shfl_sync(0xFFFFFFFFu, rev, 8 * j + t);
          const uint32_t w = hb[(i >> 3) * 32 + lane];
          acc              = (acc << 1) | ((w >> (7 - (i & 7))) & 0x0101u);
        }
        w2 = (nat & keep_nat) | (acc & ~keep_nat);
      }
      so[(2 * lane) * pitch + j0 + j]     = (uint8_t)(w2 & 0xFFu);
      so[(2 * lane + 1) * pitch + j0 + j] = (uint8_t)((w2 >> 8) & 0xFFu);
    }
  }
  __syncthreads();
  const uint32_t nblk = td.nblk;
  uint8_t*       dst  = out + td.out_off;
  if ((nb & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    const int q = nb / 16;
    for (uint32_t i = tid; i < nblk * q; i += 256) {
      const uint32_t c = i / q, k = i % q;
      const uint32_t* p = reinterpret_cast<const uint32_t*>(so + c * pitch + 16 * k);
      reinterpret_cast<uint4*>(dst + (size_t)c * nb)[k] = make_uint4(p[0], p[1], p[2], p[3]);
    }
  } else {
    for (uint32_t i = tid; i < nblk * nb; i += 256) dst[i] = so[(i / nb) * pitch + (i % nb)];
  }
  if (tid < TDEC_TILE_CB && (uint32_t)tid < nblk) {
    const uint32_t cb = td.cb0 + tid;
    const CbStatus s  = v.status[(size_t)tile * TDEC_TILE_CB + tid];
    if (crc_ok) crc_ok[cb] = s.crc_ok;
    // the caller's loop counter (sch.c:431-432): pass at which the CRC matched, else the passes spent
    if (npass) npass[cb] = s.crc_ok ? s.npass_crc : s.npass_run;
    if (npass_run) npass_run[cb] = s.npass_run;
  }
}

void launch_decide(const TdecView& v,
                   int             max_K,
                   uint8_t*        out_dev,
                   uint8_t*        crc_ok_dev,
                   uint8_t*        npass_dev,
                   uint8_t*        npass_run_dev,
                   cudaStream_t    stream)
{
  const int    nb   = max_K / 8;
  const size_t smem = (((size_t)nb * 64 + 15) / 16) * 16 + (size_t)TDEC_TILE_CB * (nb + 4);
  static std::atomic<uint64_t> attr_done{0};
  if (once_per_device(attr_done)) {
    cudaFuncSetAttribute(tdec_decide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  }
  tdec_decide_kernel<<<(unsigned)v.ntiles, 256, smem, stream>>>(v, out_dev, crc_ok_dev, npass_dev, npass_run_dev);
}

// ---------------------------------------------------------------------------------------------------------------
// Block-granular early stop (tdec_core.h: compact_*).  After a pass, per group of equal-K tiles: find the lanes that
// still run, and if they fit markedly fewer tiles, move the running lanes of the group's last tiles into the free lane
// slots of its first ones.  Everything is decided on the device; the host enqueues the two kernels after every pass of
// an early-stop decode and the next pass launches over all tiles as before (emptied tiles exit at once).
// step 1, a warp per tile (all groups at once): which lane slots hold a pair that still runs
__global__ void __launch_bounds__(256) tdec_compact_scan_kernel(TdecView v, uint32_t* __restrict__ mask)
{
  const uint32_t tile = blockIdx.x * 8u + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
  if (tile >= (uint32_t)v.ntiles) return;
  const uint32_t m = __ballot_sync(0xFFFFFFFFu, lane_running(v, tile * 32u + lane));
  if (lane == 0) mask[tile] = m;
}

// steps 2 and 3, one thread block per group: the same decisions as compact_plan_group / compact_emit_tile (tdec_core.h, run
// by the CPU emulation), with the sums and the running counts computed by the block instead of one thread.
__global__ void __launch_bounds__(256) tdec_compact_plan_kernel(TdecView v, const TileGroup* __restrict__ groups, const uint32_t* __restrict__ mask,
                                                                uint32_t* __restrict__ pref, GroupPlan* __restrict__ plans,
                                                                MoveRec* __restrict__ moves, uint32_t* __restrict__ move_counter,
                                                                uint32_t move_cap, uint32_t min_gain_tiles, uint32_t ll_max_tiles)
{
  __shared__ uint32_t  part[256];
  __shared__ uint32_t  red_run[8], red_last[8], red_int8[8];
  __shared__ GroupPlan plan;
  const TileGroup g   = groups[blockIdx.x];
  const uint32_t  tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
  // every thread owns a contiguous run of the group's tiles
  const uint32_t per = (g.ntiles + 255u) / 256u, t_lo = min(g.ntiles, tid * per), t_hi = min(g.ntiles, t_lo + per);
  uint32_t       run = 0, last = 0, int8_only = 1;
  for (uint32_t t = t_lo; t < t_hi; t++) {
    const uint32_t n = popc32(mask[g.first_tile + t]);
    run += n;
    if (n) last = t + 1;
    int8_only &= (v.fmt[g.first_tile + t] == 0u) ? 1u : 0u;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    run += __shfl_xor_sync(0xFFFFFFFFu, run, o);
    last = max(last, __shfl_xor_sync(0xFFFFFFFFu, last, o));
    int8_only &= __shfl_xor_sync(0xFFFFFFFFu, int8_only, o);
  }
  if (lane == 0) {
    red_run[wid]  = run;
    red_last[wid] = last;
    red_int8[wid] = int8_only;
  }
  __syncthreads();
  uint32_t running = 0, tiles_now = 0, all_int8 = 1;
#pragma unroll
  for (int w = 0; w < 8; w++) {
    running += red_run[w];
    tiles_now = max(tiles_now, red_last[w]);
    all_int8 &= red_int8[w];
  }
  const uint32_t need = (running + 31u) / 32u;
  const bool     go   = all_int8 && tiles_now >= need + min_gain_tiles && 5u * need <= 3u * tiles_now && tiles_now > need;
  // Tiles that still hold running lanes after this round, summed over the groups by whichever block finishes last: few
  // enough of them and the next pass goes to the low-latency kernel (TDEC_CTL_USE_LL).
  auto account = [&](uint32_t tiles_left) {
    atomicAdd(v.ctl + TDEC_CTL_TILES_LEFT, tiles_left);
    __threadfence();
    if (atomicAdd(v.ctl + TDEC_CTL_PLAN_DONE, 1u) == gridDim.x - 1u) {
      const uint32_t total        = atomicExch(v.ctl + TDEC_CTL_TILES_LEFT, 0u);
      v.ctl[TDEC_CTL_PLAN_DONE]   = 0u;
      v.ctl[TDEC_CTL_USE_LL]      = (total <= ll_max_tiles) ? 1u : 0u;
    }
  };
  if (!go) {
    if (tid == 0) {
      plans[blockIdx.x] = GroupPlan{need, 0, 0, 0};
      account(tiles_now);
    }
    return;
  }
  // running counts: free lane slots over the receiver tiles [0, need), running lanes over the donor tiles [need, tiles_now)
  uint32_t sum = 0;
  for (uint32_t t = t_lo; t < t_hi; t++) {
    const uint32_t n = popc32(mask[g.first_tile + t]);
    sum += t < need ? 32u - n : (t < tiles_now ? n : 0u);
  }
  part[tid] = sum;
  __syncthreads();
  if (tid == 0) {
    uint32_t acc = 0;
    for (int i = 0; i < 256; i++) {
      const uint32_t x = part[i];
      part[i]          = acc;
      acc += x;
    }
  }
  __syncthreads();
  // the free slots of the receivers come first in the running count: donors count from the receivers' total
  uint32_t nfree_thread = 0; // receivers' total = count in front of the thread that owns tile `need`
  {
    const uint32_t owner = min(255u, need / per);
    nfree_thread         = part[owner];
    for (uint32_t t = min(g.ntiles, owner * per); t < need; t++) nfree_thread += 32u - popc32(mask[g.first_tile + t]);
  }
  const uint32_t nfree = nfree_thread;
  uint32_t       acc   = part[tid];
  for (uint32_t t = t_lo; t < t_hi; t++) {
    const uint32_t n = popc32(mask[g.first_tile + t]);
    pref[g.first_tile + t] = t < need ? acc : acc - nfree;
    acc += t < need ? 32u - n : (t < tiles_now ? n : 0u);
  }
  if (tid == 0) {
    const uint32_t nrun = running - (32u * need - nfree); // running lanes outside the receivers
    GroupPlan      p    = GroupPlan{need, 0, 0, 0};
    if (nrun > 0 && nrun <= nfree) {
      const uint32_t base = atomicAdd(move_counter, nrun);
      if (base + nrun <= move_cap) {
        p.base  = base;
        p.moves = nrun;
        p.go    = 1;
      }
    }
    plan              = p;
    plans[blockIdx.x] = p;
    account(p.go ? need : tiles_now);
  }
  __syncthreads();
  if (!plan.go) return;
  for (uint32_t t = t_lo; t < t_hi; t++) {
    if (t < tiles_now) compact_emit_tile(g, t, mask, pref, plan, moves);
  }
}

// The data follows in GATHER form, one thread block per receiver tile: gsrc[slot] = the lane slot whose pair moves into
// `slot` (LANE_EMPTY: the slot keeps what it has).  A warp walks the tile's rows; lane l reads its new value from its source
// (lanes that share a donor tile share its 128-byte row) and the row leaves as one full, coalesced store.  (The first version
// copied column by column, 4 bytes per row and request: three times the time of a SISO pass's worth of sectors.)
__global__ void __launch_bounds__(256) tdec_compact_gsrc_kernel(TdecView v, const MoveRec* __restrict__ moves, const uint32_t* __restrict__ move_counter,
                                                                uint32_t* __restrict__ gsrc)
{
  const uint32_t n = *move_counter;
  for (uint32_t m = blockIdx.x * blockDim.x + threadIdx.x; m < n; m += gridDim.x * blockDim.x) gsrc[moves[m].dst] = moves[m].src;
}

constexpr int COMPACT_SPLIT = 4; // thread blocks per receiver tile (rows are dealt out among them)

__global__ void __launch_bounds__(256) tdec_compact_move_kernel(TdecView v, const TileGroup* __restrict__ groups, const GroupPlan* __restrict__ plans,
                                                                const uint32_t* __restrict__ gsrc, const uint32_t* __restrict__ move_counter)
{
  if (*move_counter == 0u) return;
  const uint32_t  tile = blockIdx.x, part = blockIdx.y, lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
  const TileDesc& b    = v.tiles[tile];
  const GroupPlan pl   = plans[b.group];
  if (!pl.go || tile - groups[b.group].first_tile >= pl.receivers) return; // only receiver tiles of a re-packed group take lanes
  const uint32_t src = gsrc[tile * 32u + lane];
  if (__ballot_sync(0xFFFFFFFFu, src != LANE_EMPTY) == 0u) return;
  const bool      mv  = src != LANE_EMPTY;
  const uint32_t  ts  = mv ? (src >> 5) : tile, ls = mv ? (src & 31u) : lane;
  const TileDesc& a   = v.tiles[ts];
  const uint32_t  K   = b.K, r8 = K / 8u + 1u;
  const uint32_t  w0  = part * 8u + wid; // this warp's first row; rows advance by 8 * COMPACT_SPLIT
  constexpr uint32_t STEP = 8u * COMPACT_SPLIT;
  if (mv) { // 16 bytes per lane and window row of each input stream; lanes that stay keep their rows untouched
    const u4 *s0 = a.S8 + ls, *s1 = a.P08 + ls, *s2 = a.P18 + ls;
    u4 *      d0 = b.S8 + lane, *d1 = b.P08 + lane, *d2 = b.P18 + lane;
    uint32_t  w  = w0;
    for (; w + STEP < r8; w += 2u * STEP) {
      const u4 x0 = s0[w * 32u], y0 = s1[w * 32u], z0 = s2[w * 32u];
      const u4 x1 = s0[(w + STEP) * 32u], y1 = s1[(w + STEP) * 32u], z1 = s2[(w + STEP) * 32u];
      d0[w * 32u] = x0; d1[w * 32u] = y0; d2[w * 32u] = z0;
      d0[(w + STEP) * 32u] = x1; d1[(w + STEP) * 32u] = y1; d2[(w + STEP) * 32u] = z1;
    }
    for (; w < r8; w += STEP) {
      const u4 x = s0[w * 32u], y = s1[w * 32u], z = s2[w * 32u];
      d0[w * 32u] = x; d1[w * 32u] = y; d2[w * 32u] = z;
    }
  }
  // E: every lane takes part (a lane that stays re-writes its own value), so each row leaves as one full 128-byte store
  const uint32_t* es = a.E + ls;
  uint32_t*       ed = b.E + lane;
  uint32_t        k  = w0;
  for (; k + 7u * STEP < K; k += 8u * STEP) { // eight independent rows in flight per warp
    uint32_t e[8];
#pragma unroll
    for (int u = 0; u < 8; u++) e[u] = es[(k + (uint32_t)u * STEP) * 32u];
#pragma unroll
    for (int u = 0; u < 8; u++) ed[(k + (uint32_t)u * STEP) * 32u] = e[u];
  }
  for (; k < K; k += STEP) ed[k * 32u] = es[k * 32u];
  if (part == 0 && wid == 0 && mv) { // nothing in this kernel reads the lane map or S2T of another slot
    v.S2T[(size_t)tile * 32 + lane] = v.S2T[(size_t)ts * 32 + ls];
    compact_rename(v, MoveRec{src, tile * 32u + lane});
  }
}

void launch_compact(const TdecView& v, const TileGroup* groups_dev, uint32_t ngroups, uint32_t* mask_dev, uint32_t* pref_dev,
                    GroupPlan* plans_dev, MoveRec* moves_dev, uint32_t* move_counter_dev, uint32_t* gsrc_dev, uint32_t min_gain_tiles,
                    uint32_t ll_max_tiles, int sm_count, cudaStream_t stream)
{
  cudaMemsetAsync(move_counter_dev, 0, sizeof(uint32_t), stream);
  cudaMemsetAsync(gsrc_dev, 0xFF, (size_t)v.ntiles * 32 * sizeof(uint32_t), stream); // LANE_EMPTY: every slot keeps what it has
  tdec_compact_scan_kernel<<<(unsigned)((v.ntiles + 7) / 8), 256, 0, stream>>>(v, mask_dev);
  tdec_compact_plan_kernel<<<ngroups, 256, 0, stream>>>(v, groups_dev, mask_dev, pref_dev, plans_dev, moves_dev, move_counter_dev,
                                                        (uint32_t)v.ntiles * 32u, min_gain_tiles, ll_max_tiles);
  tdec_compact_gsrc_kernel<<<(unsigned)std::min<long>(((long)v.ntiles * 32 + 255) / 256, (long)sm_count * 4), 256, 0, stream>>>(v, moves_dev, move_counter_dev, gsrc_dev);
  tdec_compact_move_kernel<<<dim3((unsigned)v.ntiles, COMPACT_SPLIT), 256, 0, stream>>>(v, groups_dev, plans_dev, gsrc_dev, move_counter_dev);
}

// ---------------------------------------------------------------------------------------------------------------
// int8 container -> int16.  16 values per thread and step when both pointers allow 16-byte accesses.
__global__ void __launch_bounds__(256) tdec_widen_i8_kernel(const int8_t* __restrict__ in, int16_t* __restrict__ out, size_t n)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0) {
    const size_t nv = n / 16;
    for (size_t i = t0; i < nv; i += stride) {
      const uint4 q = __ldcs(reinterpret_cast<const uint4*>(in) + i);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
      uint32_t       o[8];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        o[2 * k]     = sext8x2(w[k], 0);
        o[2 * k + 1] = sext8x2(w[k], 1);
      }
      uint4* dst = reinterpret_cast<uint4*>(out) + 2 * i;
      __stcs(dst, make_uint4(o[0], o[1], o[2], o[3]));
      __stcs(dst + 1, make_uint4(o[4], o[5], o[6], o[7]));
    }
    for (size_t i = nv * 16 + t0; i < n; i += stride) out[i] = (int16_t)in[i];
  } else {
    for (size_t i = t0; i < n; i += stride) out[i] = (int16_t)in[i];
  }
}

void launch_widen_i8(const int8_t* in, int16_t* out, size_t n, cudaStream_t stream)
{
  if (n == 0) return;
  size_t blocks = (n / 16 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 16) blocks = 148 * 16;
  tdec_widen_i8_kernel<<<(unsigned)blocks, 256, 0, stream>>>(in, out, n);
}

} // namespace b200
