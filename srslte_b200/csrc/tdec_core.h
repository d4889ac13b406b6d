// Batched max-log-MAP turbo decoder: memory layout and the per-lane SISO pass.
//
// What it computes (bit-exactly): the reference's GENERIC int16 decoder
//   lib/src/phy/fec/turbo/turbodecoder_gen.c:58-236   (map_gen_beta / map_gen_alpha, wrap-around int16)
//   lib/include/srsran/phy/fec/turbo/turbodecoder_iter.h:72-144  (pass schedule, extrinsic exchange, QPP permute)
//   lib/src/phy/fec/turbo/turbodecoder.c:370-378       (which vector the hard decision is taken on)
//   lib/src/phy/phch/sch.c:425-454                     (CRC after every pass, stop flag, pass counter)
//
// How it is mapped to the GPU (nothing like the reference's sub-block SIMD windows):
//   * One CUDA thread decodes TWO code blocks, one in each int16 half of its 32-bit registers (packed16.h), with the
//     exact sequential recursion of the generic decoder.  Parallelism comes from the batch: a warp owns a "tile" of
//     64 code blocks of equal K.
//   * Everything is stored k-major / code-block-minor so that one trellis step of a warp is one 128-byte row:
//       S, P0, P1 : uint4 [tile][(K+4)/4][32 lanes]   4 consecutive trellis steps of a lane's block pair per 16 B
//                   (rows K..K+2 hold the three tail steps; S2T holds encoder 2's systematic tail)
//       E         : u32   [tile][K][32 lanes]         the one extrinsic array, natural bit order, updated in place
//       CK        : uint4 [tile][K/8][2][32 lanes]    un-normalised backward metrics every 8th step (scratch)
//       HB        : u16   [tile][K/8][32 lanes]       hard decisions of the last pass run, 8 per block per entry
//   * A pass is: backward sweep over the whole block keeping only every 8th metric vector (CK), then a forward
//     sweep that, per window of 8 steps, rebuilds the 8 backward vectors in registers from its checkpoint and runs
//     the forward recursion + LLR output.  No beta array (98 KB/block in the reference, turbodecoder_gen.c:206) is
//     ever materialised.  Checkpoints hold the value BEFORE the every-4th-step normalisation, like the reference's
//     beta[] does (turbodecoder_gen.c:98-110), so the recomputed vectors are the very same int16 values.
//   * The two constituent decoders share E in place:
//       DEC1 (even pass): a-priori = E[j];              x = S[j] + E[j];  E[j]     <- L1[j] - E[j]
//       DEC2 (odd pass) : x = E[PI(i)] (no a-priori);                      E[PI(i)] <- L2[i] - x
//     which is turbodecoder_iter.h:104-128 with app1/app2/ext1 folded into one array: ext1 - app1 interleaved is
//     DEC2's systematic input, and ext2 de-interleaved minus that same value is DEC1's next a-priori.
//   * The CRC the caller's loop checks after every pass (sch.c:437-452) is accumulated on the fly as a syndrome:
//     sum over decided-one positions j of x^(K-1-j) mod g(x); zero <=> srsran_crc_checksum_byte()==0.  This works
//     in DEC2's permuted visiting order too, so no per-pass de-interleave of decisions is needed.
#pragma once
#include "packed16.h"

namespace b200 {

constexpr int      TDEC_TILE_CB = 64;  // code blocks per warp tile
constexpr int      TDEC_WIN     = 8;   // checkpoint spacing / register window
constexpr uint32_t NEG_INF2     = 0xD8F0D8F0u; // -10000 in both halves (turbodecoder_gen.c:37)

struct alignas(16) u4 {
  uint32_t x, y, z, w;
};

// per code block, 4 bytes
struct CbStatus {
  uint8_t active;    // 1 while more passes are wanted
  uint8_t crc_ok;    // CRC syndrome was zero after pass npass_crc
  uint8_t npass_crc; // 1-based pass at which the CRC first matched (0 = never)
  uint8_t npass_run; // passes actually executed (the decision in HB belongs to pass npass_run-1)
};

// CRC power table entry: x^m mod g, low 16 bits and high 8 bits, each duplicated into both halves
struct alignas(8) CrcPow {
  uint32_t lo16x2;
  uint32_t hi8x2;
};

struct TdecView {
  int K;      // code block length (one of the 188 LTE sizes, multiple of 8)
  int ntiles; // tiles of 64 blocks
  u4*             S;
  u4*             P0;
  u4*             P1;
  u4*             S2T; // [tile][32] : x,y,z = systematic tail of encoder 2 (app2[K..K+2])
  uint32_t*       E;
  u4*             CK;
  uint16_t*       HB;
  CbStatus*       status; // [ntiles*64]
  const uint16_t* qpp_fwd; // PI(i), K entries (tc_interl_lte.c:89-93)
  // CRC syndrome weights in VISITING order, nullptr = no CRC:
  const CrcPow*   crc_nat;  // [j] = x^(K-1-j) mod g        (DEC1 visits natural position j)
  const CrcPow*   crc_perm; // [i] = x^(K-1-PI(i)) mod g    (DEC2 visits natural position PI(i) at step i)
  int             early_stop; // stop a block at its first CRC match (sch.c:446-449)
  int             max_pass;
};

B200_HD size_t vec_row(const TdecView& v, int tile, int k4, int lane)
{
  return ((size_t)tile * (size_t)((v.K + 4) / 4) + (size_t)k4) * 32 + (size_t)lane;
}
B200_HD size_t e_idx(const TdecView& v, int tile, int k, int lane)
{
  return ((size_t)tile * (size_t)v.K + (size_t)k) * 32 + (size_t)lane;
}
B200_HD size_t ck_idx(const TdecView& v, int tile, int w, int half, int lane)
{
  return (((size_t)tile * (size_t)(v.K / 8) + (size_t)w) * 2 + (size_t)half) * 32 + (size_t)lane;
}
B200_HD size_t hb_idx(const TdecView& v, int tile, int w, int lane)
{
  return ((size_t)tile * (size_t)(v.K / 8) + (size_t)w) * 32 + (size_t)lane;
}

// One backward step (turbodecoder_gen.c:71-103 without the store): B <- beta_k from beta_{k+1}
B200_HD void beta_step(uint32_t B[8], uint32_t x, uint32_t y, uint32_t xy)
{
  uint32_t t2 = add2(B[5], y), t3 = add2(B[5], x), t4 = add2(B[6], x), t5 = add2(B[6], y);
  uint32_t n0 = addmax2(B[4], xy, B[0]);
  uint32_t n1 = addmax2(B[0], xy, B[4]);
  uint32_t n2 = addmax2(B[1], x, t2);
  uint32_t n3 = addmax2(B[1], y, t3);
  uint32_t n4 = addmax2(B[2], y, t4);
  uint32_t n5 = addmax2(B[2], x, t5);
  uint32_t n6 = addmax2(B[3], xy, B[7]);
  uint32_t n7 = addmax2(B[7], xy, B[3]);
  B[0] = n0; B[1] = n1; B[2] = n2; B[3] = n3; B[4] = n4; B[5] = n5; B[6] = n6; B[7] = n7;
}

// subtract state 0 from every state (turbodecoder_gen.c:105-110,186-191)
B200_HD void normalise(uint32_t M[8])
{
  uint32_t n = neg2(M[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) M[i] = add2(M[i], n);
  M[0] = 0;
}

// One forward step (turbodecoder_gen.c:139-184): returns L = max over 1-branches - max over 0-branches and
// advances A.  b = beta_k for the step's own k.
B200_HD uint32_t alpha_step(uint32_t A[8], const uint32_t b[8], uint32_t x, uint32_t y, uint32_t xy)
{
  uint32_t m0 = A[0], m1 = add2(A[3], y), m2 = add2(A[4], y), m3 = A[7];
  uint32_t m4 = A[1], m5 = add2(A[2], y), m6 = add2(A[5], y), m7 = A[6];
  uint32_t n0 = add2(A[1], xy), n1 = add2(A[2], x), n2 = add2(A[5], x), n3 = add2(A[6], xy);
  uint32_t n4 = add2(A[0], xy), n5 = add2(A[3], x), n6 = add2(A[4], x), n7 = add2(A[7], xy);

  // two half-chains per side keep the dependency depth at 5 instead of 8
  uint32_t z0 = add2(m0, b[0]);
  uint32_t z1 = add2(m4, b[4]);
  uint32_t o0 = add2(n0, b[0]);
  uint32_t o1 = add2(n4, b[4]);
  z0 = addmax2(m1, b[1], z0);
  z1 = addmax2(m5, b[5], z1);
  o0 = addmax2(n1, b[1], o0);
  o1 = addmax2(n5, b[5], o1);
  z0 = addmax2(m2, b[2], z0);
  z1 = addmax2(m6, b[6], z1);
  o0 = addmax2(n2, b[2], o0);
  o1 = addmax2(n6, b[6], o1);
  z0 = addmax2(m3, b[3], z0);
  z1 = addmax2(m7, b[7], z1);
  o0 = addmax2(n3, b[3], o0);
  o1 = addmax2(n7, b[7], o1);
  uint32_t zero_side = max2(z0, z1);
  uint32_t one_side  = max2(o0, o1);

  A[0] = max2(m0, n0); A[1] = max2(m1, n1); A[2] = max2(m2, n2); A[3] = max2(m3, n3);
  A[4] = max2(m4, n4); A[5] = max2(m5, n5); A[6] = max2(m6, n6); A[7] = max2(m7, n7);

  return sub2(one_side, zero_side);
}

// Per-lane base pointers of one tile: every access below is base + small 32-bit offset, so the address math per
// window is a couple of integer ops instead of 64-bit index products.
struct LanePtrs {
  const u4* S;   // lane's uint4 in row 0 of the tile; row r at S[r * 32]
  const u4* P;   // parity stream of the running constituent decoder
  uint32_t* E;   // lane's word in E row 0; row k at E[k * 32]
  u4*       CK;  // window w, half h at CK[(2 * w + h) * 32]
  uint16_t* HB;  // window w at HB[w * 32]
  const uint16_t* qpp;
};

template <bool DEC2>
B200_HD LanePtrs lane_ptrs(const TdecView& v, int tile, int lane)
{
  LanePtrs p;
  p.S   = v.S + vec_row(v, tile, 0, lane);
  p.P   = (DEC2 ? v.P1 : v.P0) + vec_row(v, tile, 0, lane);
  p.E   = v.E + e_idx(v, tile, 0, lane);
  p.CK  = v.CK + ck_idx(v, tile, 0, 0, lane);
  p.HB  = v.HB + hb_idx(v, tile, 0, lane);
  p.qpp = v.qpp_fwd;
  return p;
}

// Inputs of one window of 8 trellis steps for one lane
template <bool DEC2>
struct WinIn {
  u4       s[2]; // DEC1 only: systematic
  u4       p[2]; // parity of this constituent decoder
  uint32_t e[8]; // DEC1: E[j] (a-priori), DEC2: E[PI(i)] (systematic input)
  u4       q;    // DEC2 only: the eight interleaver entries PI(8w..8w+7), two per word
};

B200_HD uint32_t u4_get(const u4& q, int i)
{
  return i == 0 ? q.x : (i == 1 ? q.y : (i == 2 ? q.z : q.w));
}

// PI(8w+t) out of the packed table words
B200_HD uint32_t win_pi(const u4& q, int t)
{
  const uint32_t word = u4_get(q, t >> 1);
  return (t & 1) ? (word >> 16) : (word & 0xFFFFu);
}

template <bool DEC2, bool FIRST>
B200_HD void load_window(WinIn<DEC2>& in, const LanePtrs& p, uint32_t w)
{
  in.p[0] = p.P[(2u * w) * 32u];
  in.p[1] = p.P[(2u * w + 1u) * 32u];
  if (!DEC2) {
    in.s[0] = p.S[(2u * w) * 32u];
    in.s[1] = p.S[(2u * w + 1u) * 32u];
    if (FIRST) {
#pragma unroll
      for (int t = 0; t < 8; t++) in.e[t] = 0;
    } else {
#pragma unroll
      for (int t = 0; t < 8; t++) in.e[t] = p.E[(8u * w + (uint32_t)t) * 32u];
    }
  } else {
    // eight warp-uniform interleaver entries: one 16-byte load
    in.q = *reinterpret_cast<const u4*>(p.qpp + 8u * w);
#pragma unroll
    for (int t = 0; t < 8; t++) in.e[t] = p.E[win_pi(in.q, t) * 32u];
  }
}

// x (systematic + a-priori), y (parity) of step t inside the window
template <bool DEC2>
B200_HD void win_xy(const WinIn<DEC2>& in, int t, uint32_t& x, uint32_t& y)
{
  y = u4_get(in.p[t >> 2], t & 3);
  if (DEC2) {
    x = in.e[t];
  } else {
    x = add2(u4_get(in.s[t >> 2], t & 3), in.e[t]);
  }
}

B200_HD void store_ck(const LanePtrs& p, uint32_t w, const uint32_t B[8])
{
  u4 c0 = {B[0], B[1], B[2], B[3]}, c1 = {B[4], B[5], B[6], B[7]};
  p.CK[(2u * w) * 32u]      = c0;
  p.CK[(2u * w + 1u) * 32u] = c1;
}

// ---- backward sweep: leaves CK[w] = un-normalised beta_{8(w+1)} for every window ------------------------------
// PF windows are kept in flight in registers; a buffer is refilled (for window w-PF) right after window w consumed
// it, so a load has PF-1 windows of arithmetic to land.
template <bool DEC2, bool FIRST, int PF>
B200_HD void beta_sweep_lane(const TdecView& v, int tile, int lane)
{
  const int      K  = v.K;
  const int      nw = K / 8;
  const LanePtrs p  = lane_ptrs<DEC2>(v, tile, lane);
  uint32_t       B[8];
  B[0] = 0;
#pragma unroll
  for (int i = 1; i < 8; i++) B[i] = NEG_INF2;

  WinIn<DEC2> buf[PF];
#pragma unroll
  for (int u = 0; u < PF; u++) {
    if (nw - 1 - u >= 0) load_window<DEC2, FIRST>(buf[u], p, (uint32_t)(nw - 1 - u));
  }

  // three tail steps k = K+2, K+1, K: no a-priori, no normalisation (turbodecoder_gen.c:73-75,105)
  {
    u4 pt = p.P[(uint32_t)(K / 4) * 32u];
    u4 st = DEC2 ? v.S2T[(size_t)tile * 32 + lane] : p.S[(uint32_t)(K / 4) * 32u];
#pragma unroll
    for (int t = 2; t >= 0; t--) {
      uint32_t x = u4_get(st, t), y = u4_get(pt, t);
      beta_step(B, x, y, add2(x, y));
    }
  }
  store_ck(p, (uint32_t)(nw - 1), B);

  for (int wb = nw - 1; wb >= 0; wb -= PF) {
#pragma unroll
    for (int u = 0; u < PF; u++) {
      const int w = wb - u;
      if (w >= 0) {
#pragma unroll
        for (int t = 7; t >= 0; t--) {
          uint32_t x, y;
          win_xy<DEC2>(buf[u], t, x, y);
          beta_step(B, x, y, add2(x, y));
          if (t == 0 && w > 0) store_ck(p, (uint32_t)(w - 1), B);
          if ((t & 3) == 0) normalise(B); // k = 8w+t, always < K here
        }
        if (w - PF >= 0) load_window<DEC2, FIRST>(buf[u], p, (uint32_t)(w - PF));
      }
    }
  }
}

// ---- forward sweep ------------------------------------------------------------------------------------------------
struct LaneResult {
  uint32_t crc_lo16x2; // syndrome bits 0..15 of both blocks
  uint32_t crc_hi8x2;  // syndrome bits 16..23 of both blocks
};

// One window: rebuild its 8 backward vectors from checkpoint (c0,c1), then 8 forward steps with LLR output.
template <bool DEC2>
B200_HD void alpha_window(const TdecView&    v,
                          const LanePtrs&    p,
                          uint32_t           w,
                          const WinIn<DEC2>& cur,
                          const u4&          c0,
                          const u4&          c1,
                          const CrcPow*      cw, // the 8 syndrome weights of this window's steps, or nullptr
                          uint32_t           A[8],
                          LaneResult&        res,
                          bool               act_lo,
                          bool               act_hi)
{
  const uint32_t K = (uint32_t)v.K;
  uint32_t       xs[8], ys[8];
#pragma unroll
  for (int t = 0; t < 8; t++) win_xy<DEC2>(cur, t, xs[t], ys[t]);

  // bw[t] is the vector the forward step at position 8w+t needs (beta_{8w+t+1})
  uint32_t bw[8][8];
  uint32_t B[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
  for (int i = 0; i < 8; i++) bw[7][i] = B[i];
  if (8u * w + 8u < K) normalise(B); // beta_{8w+8} sits on a multiple of 4; the last one (k==K) is not normalised
#pragma unroll
  for (int t = 7; t >= 1; t--) {
    beta_step(B, xs[t], ys[t], add2(xs[t], ys[t]));
#pragma unroll
    for (int i = 0; i < 8; i++) bw[t - 1][i] = B[i];
    if (t == 4) normalise(B);
  }

  uint32_t bits = 0; // two 8-bit shift registers: bits 0..7 low block, 16..23 high block
#pragma unroll
  for (int t = 0; t < 8; t++) {
    uint32_t L = alpha_step(A, bw[t], xs[t], ys[t], add2(xs[t], ys[t]));
    if ((t & 3) == 3) normalise(A); // forward index k = 8w+t+1 (turbodecoder_gen.c:186)
    const uint32_t pos = DEC2 ? win_pi(cur.q, t) : 8u * w + (uint32_t)t;
    // extrinsic hand-over, in place (turbodecoder_iter.h:108,118-127)
    p.E[pos * 32u] = sub2(L, cur.e[t]);
    uint32_t one   = pos2(L); // turbodecoder_gen.c:266: LLR > 0 -> 1
    bits           = (bits << 1) | one;
    if (cw) {
      const CrcPow   c    = cw[t];
      const uint32_t mask = one * 0xFFFFu;
      res.crc_lo16x2 ^= (c.lo16x2 & mask);
      res.crc_hi8x2 ^= (c.hi8x2 & mask);
    }
  }
  // hard decisions of this window: low byte = low block, high byte = high block, MSB = first step
  const uint16_t hb = (uint16_t)((bits & 0xFFu) | ((bits >> 8) & 0xFF00u));
  if (act_lo && act_hi) {
    p.HB[w * 32u] = hb;
  } else {
    const uint16_t old = p.HB[w * 32u];
    const uint16_t m   = (uint16_t)((act_lo ? 0x00FFu : 0u) | (act_hi ? 0xFF00u : 0u));
    p.HB[w * 32u]      = (uint16_t)((hb & m) | (old & ~m));
  }
}

template <bool DEC2, bool FIRST>
B200_HD LaneResult alpha_sweep_lane(const TdecView& v, int tile, int lane, bool act_lo, bool act_hi)
{
  const uint32_t nw = (uint32_t)v.K / 8u;
  const LanePtrs p  = lane_ptrs<DEC2>(v, tile, lane);
  const CrcPow*  cbase = DEC2 ? v.crc_perm : v.crc_nat;
  uint32_t       A[8];
  LaneResult     res = {0u, 0u};
  A[0]               = 0;
#pragma unroll
  for (int i = 1; i < 8; i++) A[i] = NEG_INF2;

  // two register buffers, windows alternate between them (no copies); the other buffer's loads are issued before
  // the current window's ~450 instructions of arithmetic
  WinIn<DEC2> inA, inB;
  u4          ckA0, ckA1, ckB0, ckB1;
  load_window<DEC2, FIRST>(inA, p, 0u);
  ckA0 = p.CK[0];
  ckA1 = p.CK[32];
  for (uint32_t w = 0; w < nw; w += 2) {
    if (w + 1 < nw) {
      load_window<DEC2, FIRST>(inB, p, w + 1);
      ckB0 = p.CK[(2u * (w + 1)) * 32u];
      ckB1 = p.CK[(2u * (w + 1) + 1u) * 32u];
    }
    alpha_window<DEC2>(v, p, w, inA, ckA0, ckA1, cbase ? cbase + 8u * w : nullptr, A, res, act_lo, act_hi);
    if (w + 1 < nw) {
      if (w + 2 < nw) {
        load_window<DEC2, FIRST>(inA, p, w + 2);
        ckA0 = p.CK[(2u * (w + 2)) * 32u];
        ckA1 = p.CK[(2u * (w + 2) + 1u) * 32u];
      }
      alpha_window<DEC2>(v, p, w + 1, inB, ckB0, ckB1, cbase ? cbase + 8u * (w + 1) : nullptr, A, res, act_lo, act_hi);
    }
  }
  return res;
}

// ---- end of a pass: CRC verdict, pass counters, stop flag (sch.c:431-452) ---------------------------------------
B200_HD void finish_pass(const TdecView&   v,
                         CbStatus*         st,
                         CbStatus          s_lo,
                         CbStatus          s_hi,
                         bool              act_lo,
                         bool              act_hi,
                         const LaneResult& r,
                         int               pass_idx)
{
  const bool last = (pass_idx + 1 >= v.max_pass);
  if (act_lo) {
    bool ok        = v.crc_nat && ((r.crc_lo16x2 & 0xFFFFu) == 0) && ((r.crc_hi8x2 & 0xFFu) == 0);
    s_lo.npass_run = (uint8_t)(pass_idx + 1);
    if (ok && !s_lo.crc_ok) {
      s_lo.crc_ok    = 1;
      s_lo.npass_crc = (uint8_t)(pass_idx + 1);
    }
    if (last || (ok && v.early_stop)) s_lo.active = 0;
    st[0] = s_lo;
  }
  if (act_hi) {
    bool ok        = v.crc_nat && ((r.crc_lo16x2 >> 16) == 0) && (((r.crc_hi8x2 >> 16) & 0xFFu) == 0);
    s_hi.npass_run = (uint8_t)(pass_idx + 1);
    if (ok && !s_hi.crc_ok) {
      s_hi.crc_ok    = 1;
      s_hi.npass_crc = (uint8_t)(pass_idx + 1);
    }
    if (last || (ok && v.early_stop)) s_hi.active = 0;
    st[1] = s_hi;
  }
}

// ---- one full pass for one lane (two code blocks), direct-from-global variant ------------------------------------
// Used by the CPU emulation (and kept as the reference structure of the pass); the GPU kernel in tdec_kernels.cu runs
// the same steps but feeds the windows through a TMA-filled shared-memory ring.
template <bool DEC2, bool FIRST, int PF>
B200_HD void siso_pass_lane(const TdecView& v, int tile, int lane, int pass_idx)
{
  CbStatus* st     = v.status + ((size_t)tile * TDEC_TILE_CB + 2 * (size_t)lane);
  CbStatus  s_lo   = st[0];
  CbStatus  s_hi   = st[1];
  bool      act_lo = s_lo.active != 0, act_hi = s_hi.active != 0;
  if (!act_lo && !act_hi) {
    return;
  }
  beta_sweep_lane<DEC2, FIRST, PF>(v, tile, lane);
  LaneResult r = alpha_sweep_lane<DEC2, FIRST>(v, tile, lane, act_lo, act_hi);
  finish_pass(v, st, s_lo, s_hi, act_lo, act_hi, r, pass_idx);
}

} // namespace b200

// ---------------------------------------------------------------------------------------------------------------
// Layout conversion at the two ends of a decode, written per element so the CUDA kernels (tdec_kernels.cu) and the
// host emulation (tests) share one definition.
namespace b200 {

// Natural decoder input of one block: 3K+12 int16, in[3i+j] = stream j of bit i, then 12 tail values
// (turbodecoder_gen.c:238-258).  Returns the value that belongs at trellis row k (0..K+3) of stream
// `which` (0 = S, 1 = P0, 2 = P1, 3 = S2T row k-K) for that block.
B200_HD int16_t natural_pick(const int16_t* in, int K, int which, int k)
{
  if (k < K) {
    return which < 3 ? in[3 * k + which] : (int16_t)0;
  }
  const int t = k - K;
  if (t > 2) {
    return 0;
  }
  switch (which) {
    case 0:
      return in[3 * K + 2 * t];
    case 1:
      return in[3 * K + 2 * t + 1];
    case 2:
      return in[3 * K + 6 + 2 * t + 1];
    default:
      return in[3 * K + 6 + 2 * t];
  }
}

// Decided byte jb (bits 8jb..8jb+7, MSB first, natural order) of code block cb after its last pass.
// rev = inverse QPP table (tc_interl_lte.c:93); after an odd pass HB is in DEC2's visiting order and bit j sits at
// visiting index rev[j] (the reference instead de-interleaves the whole LLR vector, turbodecoder_iter.h:127).
B200_HD uint8_t decide_byte(const TdecView& v, const uint16_t* rev, int cb, int jb)
{
  const int      tile = cb / TDEC_TILE_CB, lane = (cb % TDEC_TILE_CB) >> 1, half = cb & 1;
  const CbStatus st   = v.status[cb];
  const bool     perm = st.npass_run > 0 && ((st.npass_run - 1) & 1);
  if (!perm) {
    uint16_t hb = v.HB[hb_idx(v, tile, jb, lane)];
    return (uint8_t)(half ? (hb >> 8) : (hb & 0xFF));
  }
  uint32_t byte = 0;
  for (int t = 0; t < 8; t++) {
    int      i  = rev[8 * jb + t];
    uint16_t hb = v.HB[hb_idx(v, tile, i >> 3, lane)];
    uint32_t b  = half ? (hb >> 8) : (hb & 0xFF);
    byte        = (byte << 1) | ((b >> (7 - (i & 7))) & 1u);
  }
  return (uint8_t)byte;
}

} // namespace b200
