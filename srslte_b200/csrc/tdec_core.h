// Batched max-log-MAP turbo decoder: memory layout and the per-lane SISO arithmetic.
//
// What it computes (bit-exactly): the reference's GENERIC int16 decoder
//   lib/src/phy/fec/turbo/turbodecoder_gen.c:58-236   (map_gen_beta / map_gen_alpha, wrap-around int16)
//   lib/include/srsran/phy/fec/turbo/turbodecoder_iter.h:72-144  (pass schedule, extrinsic exchange, QPP permute)
//   lib/src/phy/fec/turbo/turbodecoder.c:370-378       (which vector the hard decision is taken on)
//   lib/src/phy/phch/sch.c:425-454                     (CRC after every pass, stop flag, pass counter)
//
// How it is mapped to the GPU (nothing like the reference's sub-block SIMD windows):
//   * One CUDA thread decodes TWO code blocks, one in each int16 half of its 32-bit registers (packed16.h), with the
//     exact sequential recursions of the generic decoder.  A "tile" is 64 code blocks of equal K = 32 lanes.
//   * TWO warps work on a tile, from both ends of the trellis (the recursions are exact, so this is the only
//     intra-block parallelism there is).  With ws = the split window:
//       phase 1   warp F: forward (alpha) recursion over windows [0, ws), leaving alpha checkpoints  CK[w] = alpha_{8w}
//                 warp B: tail + backward (beta) recursion over windows [ws, nw), leaving            CK[w] = beta_{8w+8}
//       phase 2   warp F: windows ws..nw-1 upwards: rebuild the window's 8 beta vectors in registers from CK[w], then
//                         8 forward steps with LLR output (fwd_window)
//                 warp B: windows ws-1..0 downwards: rebuild the window's 8 alpha vectors from CK[w], then 8 backward
//                         steps with LLR output (bwd_window)
//     No beta array (98 KB/block in the reference, turbodecoder_gen.c:206) is ever materialised.  beta checkpoints hold
//     the value BEFORE the every-4th-step normalisation, like the reference's beta[] does (turbodecoder_gen.c:98-110);
//     alpha checkpoints sit on multiples of 8 and are therefore freshly normalised (turbodecoder_gen.c:186-191).
//   * Everything is stored k-major / code-block-minor so that one trellis step of a warp is one contiguous row.  A batch
//     may mix code block lengths: every tile has its own K and its own arrays (TileDesc), tiles are ordered by
//     descending K so that one launch per pass covers the whole batch, longest tiles first:
//       int16 inputs  S, P0, P1 : uint4 [(K+4)/4][32 lanes]  4 consecutive steps of a lane's block pair per 16 B
//       int8 inputs   S8,P08,P18: uint4 [K/8+1][32 lanes]    8 consecutive steps x 2 blocks per 16 B
//                   (used for a tile when every channel LLR of its 64 blocks fits int8: half the bytes per sweep;
//                    fmt[tile] says which set is valid; rows past K hold the three tail steps; S2T = encoder 2's
//                    systematic tail, always int16)
//       E         : u32   [K][32 lanes]         the one extrinsic array, natural bit order, updated in place
//       CK        : uint4 [K/8][2][32 lanes]    checkpoints (scratch), alpha below the split, beta above
//       HB        : u16   [K/8][32 lanes]       hard decisions of the last pass run, 8 per block per entry
//   * Block-granular early stop: a lane slot (tile, lane) HOLDS a block pair, named by LaneMap {st0, hb0} = where the
//     pair's status records and decision words live (its HOME, the slot it was loaded into).  Between passes the lanes
//     that still run are re-packed into fewer tiles (compact_* below move a lane's S8/P08/P18/S2T/E columns and its
//     LaneMap into a free slot of an earlier tile of the same K and CRC kind); status and HB stay at home, so the
//     decision kernel and the caller never see the move, and tiles left without a running lane exit at once.
//   * The two constituent decoders share E in place:
//       DEC1 (even pass): a-priori = E[j];              x = S[j] + E[j];  E[j]     <- L1[j] - E[j]
//       DEC2 (odd pass) : x = E[PI(i)] (no a-priori);                      E[PI(i)] <- L2[i] - x
//     which is turbodecoder_iter.h:104-128 with app1/app2/ext1 folded into one array: ext1 - app1 interleaved is
//     DEC2's systematic input, and ext2 de-interleaved minus that same value is DEC1's next a-priori.
//   * The CRC the caller's loop checks after every pass (sch.c:437-452) is accumulated on the fly as a syndrome:
//     sum over decided-one positions j of x^(K-1-j) mod g(x); zero <=> srsran_crc_checksum_byte()==0.  It is a XOR,
//     so it works in DEC2's permuted visiting order and in either sweep direction.
#pragma once
#include "packed16.h"

namespace b200 {

constexpr int      TDEC_TILE_CB = 64;  // code blocks per tile
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

constexpr uint32_t LANE_EMPTY = 0xFFFFFFFFu;

// TdecView::ctl
constexpr int TDEC_CTL_USE_LL      = 0; // 1: so few tiles still run that the next pass goes to the low-latency kernel
constexpr int TDEC_CTL_LL_QUEUE    = 1; // low-latency pass: next tile to hand out
constexpr int TDEC_CTL_PLAN_DONE   = 2; // re-packing: groups planned so far (the last one decides TDEC_CTL_USE_LL)
constexpr int TDEC_CTL_TILES_LEFT  = 3; // re-packing: tiles that still hold running lanes, summed over the groups
constexpr int TDEC_CTL_WORDS       = 8;

// which block pair a lane slot holds
struct LaneMap {
  uint32_t st0; // index of the pair's first record in status[] (LANE_EMPTY: the slot holds nothing)
  uint32_t hb0; // index of the pair's window-0 decision word in HB[] (window w at hb0 + 32 w)
};

// per tile of 64 code blocks of equal K
struct alignas(16) TileDesc {
  u4*             S8;   // int8 rows  [K/8+1][32]
  u4*             P08;
  u4*             P18;
  u4*             S;    // int16 rows [(K+4)/4][32], valid when fmt[tile] != 0
  u4*             P0;
  u4*             P1;
  uint32_t*       E;    // [K][32]
  u4*             CK;   // [K/8][2][32]
  const uint16_t* qpp_fwd; // PI(i), K entries (tc_interl_lte.c:89-93)
  const uint16_t* qpp_rev; // inverse table
  // CRC syndrome weights in VISITING order, nullptr = no CRC:
  const CrcPow*   crc_nat;  // [j] = x^(K-1-j) mod g        (DEC1 visits natural position j)
  const CrcPow*   crc_perm; // [i] = x^(K-1-PI(i)) mod g    (DEC2 visits natural position PI(i) at step i)
  uint64_t        llr_off;  // int16 offset of the natural input vector of the tile's first block (contiguous inputs)
  uint64_t        out_off;  // byte offset of the tile's first block in the decision output
  uint32_t        K;        // code block length (one of the 188 LTE sizes, multiple of 8)
  uint32_t        hb_row0;  // first row (of 32 u16) of the tile's decisions in HB
  uint32_t        cb0;      // index of the tile's first code block in the per-block arrays (offset list, crc_ok, npass)
  uint32_t        nblk;     // code blocks in this tile (1..64)
  uint32_t        group;    // (K, CRC kind) group the tile belongs to: lanes are only re-packed inside a group
  uint32_t        pad[3];
};

// consecutive tiles of equal K and CRC kind
struct TileGroup {
  uint32_t first_tile, ntiles;
};

struct MoveRec {
  uint32_t src, dst; // lane slots: tile * 32 + lane
};

struct TdecView {
  int              ntiles; // tiles of 64 blocks
  int              split_percent; // share of a tile's trellis windows below the split (tdec_split)
  const TileDesc*  tiles;  // [ntiles]
  uint32_t*        fmt;    // [ntiles] 0: the int8 arrays hold the tile, 1: the int16 arrays do
  uint32_t*        err;    // bit 0: a tile needed the int16 arrays but the workspace was carved without them (TileDesc::S == nullptr)
  uint32_t*        ctl;    // device-side control words, see TDEC_CTL_*
  u4*              ll_ck;  // low-latency pass: alpha checkpoints of the tile a thread block works on, one slot per thread block
  uint32_t         ll_ck_slot; // u4 elements per slot (max K / 8 * 64)
  u4*              S2T;    // [ntiles*32] : x,y,z = systematic tail of encoder 2 (app2[K..K+2]) of the pair in that slot
  LaneMap*         lanes;  // [ntiles*32]
  uint16_t*        HB;
  CbStatus*        status; // [ntiles*64], index = home tile * 64 + block of the tile
  int              early_stop; // stop a block at its first CRC match (sch.c:446-449)
  int              max_pass;
};

// Split of the trellis between the two warps.  Per step warp F spends ~15 instructions below the split and ~56 above,
// warp B ~62 below (its alpha rebuild cannot share the branch sums with the LLR) and ~15 above: measured best at ws = 0.51 nw since warp B shares its backward terms with the LLR (sweeps in profiles/README.md).
B200_HD int tdec_split(int K, int percent)
{
  const int nw = K / 8;
  int       ws = (nw * percent + 50) / 100;
  if (ws < 1) ws = 1;
  if (ws > nw - 1) ws = nw - 1;
  return ws;
}

// rows of a tile's arrays (all offsets fit 32 bits: a tile's largest array, E, is K * 128 bytes < 1 MB)
B200_HD uint32_t vec_row(int k4, int lane)
{
  return (uint32_t)k4 * 32u + (uint32_t)lane;
}
B200_HD uint32_t row8(int w, int lane)
{
  return (uint32_t)w * 32u + (uint32_t)lane;
}
B200_HD uint32_t e_idx(int k, int lane)
{
  return (uint32_t)k * 32u + (uint32_t)lane;
}
B200_HD uint32_t ck_idx(int w, int half, int lane)
{
  return ((uint32_t)w * 2u + (uint32_t)half) * 32u + (uint32_t)lane;
}
// decision word of window w of the pair whose home is hb0
B200_HD size_t hb_idx(uint32_t hb0, int w)
{
  return (size_t)hb0 + (size_t)w * 32u;
}
// home of the pair loaded into (tile, lane)
B200_HD LaneMap lane_home(const TileDesc& td, int tile, int lane)
{
  return LaneMap{(uint32_t)tile * (uint32_t)TDEC_TILE_CB + 2u * (uint32_t)lane, td.hb_row0 * 32u + (uint32_t)lane};
}

B200_HD uint32_t u4_get(const u4& q, int i)
{
  return i == 0 ? q.x : (i == 1 ? q.y : (i == 2 ? q.z : q.w));
}

// two int8 (bytes 0,1 or 2,3 of word) sign-extended into an int16x2
B200_HD uint32_t sext8x2(uint32_t word, int odd)
{
#if defined(__CUDA_ARCH__)
  uint32_t r;
  if (odd) {
    asm("prmt.b32 %0, %1, 0, 0xB3A2;" : "=r"(r) : "r"(word));
  } else {
    asm("prmt.b32 %0, %1, 0, 0x9180;" : "=r"(r) : "r"(word));
  }
  return r;
#else
  const int8_t lo = (int8_t)(uint8_t)(word >> (odd ? 16 : 0)), hi = (int8_t)(uint8_t)(word >> (odd ? 24 : 8));
  return pack2((int16_t)lo, (int16_t)hi);
#endif
}

// value of step t (0..7) of a window out of its raw words: int16 format = two uint4 (4 steps each), int8 = one uint4
template <bool IN8>
B200_HD uint32_t win_val(const u4 q[2], int t)
{
  if (IN8) {
    return sext8x2(u4_get(q[0], t >> 1), t & 1);
  }
  return u4_get(q[t >> 2], t & 3);
}

// PI(8w+t) out of the packed table words
B200_HD uint32_t win_pi(const u4& q, int t)
{
  const uint32_t word = u4_get(q, t >> 1);
  return (t & 1) ? (word >> 16) : (word & 0xFFFFu);
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

// One forward step without output (turbodecoder_gen.c:139-184, the state update only): A <- alpha_{k+1} from alpha_k.
// Same sixteen branch sums as alpha_step below, folded into add+max pairs (max is commutative, so the result is the
// same int16 value).
B200_HD void alpha_update(uint32_t A[8], uint32_t x, uint32_t y, uint32_t xy)
{
  uint32_t t1 = add2(A[3], y), t2 = add2(A[4], y), t5 = add2(A[2], y), t6 = add2(A[5], y);
  uint32_t n0 = addmax2(A[1], xy, A[0]);
  uint32_t n1 = addmax2(A[2], x, t1);
  uint32_t n2 = addmax2(A[5], x, t2);
  uint32_t n3 = addmax2(A[6], xy, A[7]);
  uint32_t n4 = addmax2(A[0], xy, A[1]);
  uint32_t n5 = addmax2(A[3], x, t5);
  uint32_t n6 = addmax2(A[4], x, t6);
  uint32_t n7 = addmax2(A[7], xy, A[6]);
  A[0] = n0; A[1] = n1; A[2] = n2; A[3] = n3; A[4] = n4; A[5] = n5; A[6] = n6; A[7] = n7;
}

// subtract state 0 from every state (turbodecoder_gen.c:105-110,186-191)
B200_HD void normalise(uint32_t M[8])
{
  uint32_t n = neg2(M[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) M[i] = add2(M[i], n);
  M[0] = 0;
}

// LLR of one step (turbodecoder_gen.c:139-184): L = max over 1-branches - max over 0-branches of
// alpha_k[s] + gamma + beta_{k+1}[s'].  b = the stored (un-normalised) beta_{k+1}.  UPDATE also advances A.
template <bool UPDATE>
B200_HD uint32_t llr_step(uint32_t A[8], const uint32_t b[8], uint32_t x, uint32_t y, uint32_t xy)
{
  uint32_t m0 = A[0], m1 = add2(A[3], y), m2 = add2(A[4], y), m3 = A[7];
  uint32_t m4 = A[1], m5 = add2(A[2], y), m6 = add2(A[5], y), m7 = A[6];
  uint32_t n0 = add2(A[1], xy), n1 = add2(A[2], x), n2 = add2(A[5], x), n3 = add2(A[6], xy);
  uint32_t n4 = add2(A[0], xy), n5 = add2(A[3], x), n6 = add2(A[4], x), n7 = add2(A[7], xy);

  // two half-chains per side keep the dependency depth at 5 instead of 8.  The packed add and the packed add+max issue
  // on different half-rate pipes; warp F (UPDATE) is add-heavy, so its four chain heads are written as max(a+b, -32768),
  // which is the same value on the other pipe.
  constexpr uint32_t MIN2 = 0x80008000u;
  uint32_t z0 = UPDATE ? addmax2(m0, b[0], MIN2) : add2(m0, b[0]);
  uint32_t z1 = UPDATE ? addmax2(m4, b[4], MIN2) : add2(m4, b[4]);
  uint32_t o0 = UPDATE ? addmax2(n0, b[0], MIN2) : add2(n0, b[0]);
  uint32_t o1 = UPDATE ? addmax2(n4, b[4], MIN2) : add2(n4, b[4]);
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

  if (UPDATE) {
    A[0] = max2(m0, n0); A[1] = max2(m1, n1); A[2] = max2(m2, n2); A[3] = max2(m3, n3);
    A[4] = max2(m4, n4); A[5] = max2(m5, n5); A[6] = max2(m6, n6); A[7] = max2(m7, n7);
  }
  return sub2(one_side, zero_side);
}

// LLR of one step AND the backward update in one go, for warp B's steps that need no normalisation in between: the sixteen
// terms beta_{k+1}[s'] + gamma of the backward recursion (turbodecoder_gen.c:76-92) are also, grouped by the PREVIOUS state s,
// the second summands of the sixteen LLR candidates alpha_k[s] + (gamma + beta_{k+1}[s']) (:160-176).  int16 addition wraps and
// is associative, so every candidate and every new beta is the same 16-bit value the reference computes with
// (alpha + gamma) + beta; 39 packed operations instead of the 43 of llr_step<false> + beta_step.
//   previous state s -> (next state, gamma) on bit 0: 0->(0,0) 1->(4,0) 2->(5,y) 3->(1,y) 4->(2,y) 5->(6,y) 6->(7,0) 7->(3,0)
//                                         on bit 1: 0->(4,xy) 1->(0,xy) 2->(1,x) 3->(5,x) 4->(6,x) 5->(2,x) 6->(3,xy) 7->(7,xy)
B200_HD uint32_t llr_beta_step(const uint32_t A[8], uint32_t U[8], uint32_t x, uint32_t y, uint32_t xy)
{
  const uint32_t z0 = U[0], z1 = U[4], z2 = add2(U[5], y), z3 = add2(U[1], y);
  const uint32_t z4 = add2(U[2], y), z5 = add2(U[6], y), z6 = U[7], z7 = U[3];
  const uint32_t o0 = add2(U[4], xy), o1 = add2(U[0], xy), o2 = add2(U[1], x), o3 = add2(U[5], x);
  const uint32_t o4 = add2(U[6], x), o5 = add2(U[2], x), o6 = add2(U[3], xy), o7 = add2(U[7], xy);
  uint32_t a0 = add2(A[0], z0), a1 = add2(A[4], z4), b0 = add2(A[0], o0), b1 = add2(A[4], o4);
  a0 = addmax2(A[1], z1, a0);
  a1 = addmax2(A[5], z5, a1);
  b0 = addmax2(A[1], o1, b0);
  b1 = addmax2(A[5], o5, b1);
  a0 = addmax2(A[2], z2, a0);
  a1 = addmax2(A[6], z6, a1);
  b0 = addmax2(A[2], o2, b0);
  b1 = addmax2(A[6], o6, b1);
  a0 = addmax2(A[3], z3, a0);
  a1 = addmax2(A[7], z7, a1);
  b0 = addmax2(A[3], o3, b0);
  b1 = addmax2(A[7], o7, b1);
  U[0] = max2(z0, o0); U[1] = max2(z1, o1); U[2] = max2(z2, o2); U[3] = max2(z3, o3);
  U[4] = max2(z4, o4); U[5] = max2(z5, o5); U[6] = max2(z6, o6); U[7] = max2(z7, o7);
  return sub2(max2(b0, b1), max2(a0, a1));
}

// ---- one window of 8 trellis steps in registers ---------------------------------------------------------------
struct WinRegs {
  uint32_t xs[8]; // systematic (+ a-priori) input of the constituent decoder
  uint32_t ys[8]; // its parity input
  uint32_t es[8]; // what is subtracted from the LLR to form the new extrinsic: DEC1 the a-priori, DEC2 xs itself
};

// s, p: raw words of the window (s unused for DEC2), e: the eight E words (DEC1: E[8w+t], DEC2: E[PI(8w+t)])
template <bool DEC2, bool FIRST, bool IN8>
B200_HD void win_unpack(WinRegs& r, const u4 s[2], const u4 p[2], const uint32_t e[8])
{
#pragma unroll
  for (int t = 0; t < 8; t++) {
    r.ys[t] = win_val<IN8>(p, t);
    if (DEC2) {
      r.es[t] = e[t];
      r.xs[t] = e[t];
    } else if (FIRST) {
      r.es[t] = 0;
      r.xs[t] = win_val<IN8>(s, t);
    } else {
      r.es[t] = e[t];
      r.xs[t] = add2(win_val<IN8>(s, t), e[t]);
    }
  }
}

struct LaneResult {
  uint32_t crc_lo16x2; // syndrome bits 0..15 of both blocks
  uint32_t crc_hi8x2;  // syndrome bits 16..23 of both blocks
};

struct WinOut {
  uint32_t enew[8]; // new extrinsic of each step (turbodecoder_iter.h:108,118-127)
  uint32_t bits;    // decisions (turbodecoder_gen.c:266: LLR > 0 -> 1): bit 7-t = step t, low block in bits 0..7, high in 16..23
};

// sink(t, L) receives the a-posteriori LLR of step t and disposes of the new extrinsic L - e (turbodecoder_iter.h:108,118-127):
// the GPU kernel re-reads e from its shared-memory stage and stores the difference straight away, so that neither e nor the
// result occupies a register beyond the step; the plain overloads below keep both in WinRegs / WinOut.
template <class Sink>
B200_HD void win_emit(WinOut& o, LaneResult& res, const CrcPow* cw, int t, uint32_t L, Sink&& sink)
{
  sink(t, L);
  const uint32_t one = pos2(L);
#if defined(__CUDA_ARCH__)
  // bits += one * 2^(7-t) as an integer multiply-add: it issues on the FMA pipe, which idles while the packed-halfword
  // instructions keep the ALU pipe busy (a shift + OR would go there as well); the bit positions never collide, so + is |
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(o.bits) : "r"(one), "r"(1u << (7 - t)));
#else
  o.bits |= one << (7 - t);
#endif
  if (cw) {
    const CrcPow   c    = cw[t];
    const uint32_t mask = one * 0xFFFFu;
    res.crc_lo16x2 ^= (c.lo16x2 & mask);
    res.crc_hi8x2 ^= (c.hi8x2 & mask);
  }
}
struct KeepInWinOut {
  WinOut&        o;
  const WinRegs& r;
  B200_HD void   operator()(int t, uint32_t L) const { o.enew[t] = sub2(L, r.es[t]); }
};

// Warp F, phase 2.  A = alpha_{8w} on entry, alpha_{8w+8} on exit.  ck = beta_{8w+8} un-normalised; norm_ck = (8w+8 < K):
// the recursion continues from the normalised value except at the very end of the block (turbodecoder_gen.c:105).
template <class Sink>
B200_HD void fwd_window(uint32_t A[8], const uint32_t ck[8], bool norm_ck, const WinRegs& r, const CrcPow* cw, LaneResult& res, WinOut& o,
                        Sink&& sink)
{
  // bw[t] is the vector the forward step at position 8w+t needs (beta_{8w+t+1})
  uint32_t bw[8][8];
  uint32_t B[8];
#pragma unroll
  for (int i = 0; i < 8; i++) B[i] = bw[7][i] = ck[i];
  if (norm_ck) normalise(B);
#pragma unroll
  for (int t = 7; t >= 1; t--) {
    beta_step(B, r.xs[t], r.ys[t], add2(r.xs[t], r.ys[t]));
#pragma unroll
    for (int i = 0; i < 8; i++) bw[t - 1][i] = B[i];
    if (t == 4) normalise(B);
  }
  o.bits = 0;
#pragma unroll
  for (int t = 0; t < 8; t++) {
    const uint32_t L = llr_step<true>(A, bw[t], r.xs[t], r.ys[t], add2(r.xs[t], r.ys[t]));
    if ((t & 3) == 3) normalise(A); // forward index k = 8w+t+1 (turbodecoder_gen.c:186)
    win_emit(o, res, cw, t, L, sink);
  }
}
B200_HD void fwd_window(uint32_t A[8], const uint32_t ck[8], bool norm_ck, const WinRegs& r, const CrcPow* cw, LaneResult& res, WinOut& o)
{
  fwd_window(A, ck, norm_ck, r, cw, res, o, KeepInWinOut{o, r});
}

// Warp B, phase 2.  U = beta_{8w+8} un-normalised on entry (8w+8 < K always: this warp works below the split),
// beta_{8w} un-normalised on exit.  ack = alpha_{8w}.
template <class Sink>
B200_HD void bwd_window(uint32_t U[8], const uint32_t ack[8], const WinRegs& r, const CrcPow* cw, LaneResult& res, WinOut& o, Sink&& sink)
{
  // aw[t] = alpha_{8w+t}, the vector the LLR of step 8w+t needs
  uint32_t aw[8][8];
  uint32_t A[8];
#pragma unroll
  for (int i = 0; i < 8; i++) A[i] = aw[0][i] = ack[i];
#pragma unroll
  for (int t = 0; t < 7; t++) {
    alpha_update(A, r.xs[t], r.ys[t], add2(r.xs[t], r.ys[t]));
    if (t == 3) normalise(A);
#pragma unroll
    for (int i = 0; i < 8; i++) aw[t + 1][i] = A[i];
  }
  o.bits = 0;
#pragma unroll
  for (int t = 7; t >= 0; t--) {
    const uint32_t xy = add2(r.xs[t], r.ys[t]);
    uint32_t       L;
    if ((t & 3) == 3) { // U is beta_{8w+t+1}, index a multiple of 4 (and < K): the LLR takes it as stored, the recursion normalised
      L = llr_step<false>(aw[t], U, r.xs[t], r.ys[t], xy);
      normalise(U);
      beta_step(U, r.xs[t], r.ys[t], xy);
    } else {
      L = llr_beta_step(aw[t], U, r.xs[t], r.ys[t], xy);
    }
    win_emit(o, res, cw, t, L, sink);
  }
}
B200_HD void bwd_window(uint32_t U[8], const uint32_t ack[8], const WinRegs& r, const CrcPow* cw, LaneResult& res, WinOut& o)
{
  bwd_window(U, ack, r, cw, res, o, KeepInWinOut{o, r});
}

// Warp F, phase 1: 8 forward steps, no output
B200_HD void alpha_window(uint32_t A[8], const WinRegs& r)
{
#pragma unroll
  for (int t = 0; t < 8; t++) {
    alpha_update(A, r.xs[t], r.ys[t], add2(r.xs[t], r.ys[t]));
    if ((t & 3) == 3) normalise(A);
  }
}

// Warp B, phase 1: 8 backward steps.  On exit B = beta_{8w} un-normalised (the caller stores it as CK[w-1] and then
// normalises, or hands it to phase 2 as is when w is the split window).
B200_HD void beta_window(uint32_t B[8], const WinRegs& r)
{
#pragma unroll
  for (int t = 7; t >= 0; t--) {
    beta_step(B, r.xs[t], r.ys[t], add2(r.xs[t], r.ys[t]));
    if (t == 4) normalise(B); // k = 8w+4
  }
}

// hard decisions of a window: low byte = low block, high byte = high block, MSB = first step
B200_HD uint16_t hb_word(uint32_t bits)
{
  return (uint16_t)((bits & 0xFFu) | ((bits >> 8) & 0xFF00u));
}
B200_HD void hb_store(uint16_t* dst, uint32_t bits, bool act_lo, bool act_hi)
{
  const uint16_t hb = hb_word(bits);
  if (act_lo && act_hi) {
    *dst = hb;
  } else { // a block that already stopped keeps its decisions: byte stores, nothing is read back
    uint8_t* b = reinterpret_cast<uint8_t*>(dst); // little endian: byte 0 = low block
    if (act_lo) b[0] = (uint8_t)(hb & 0xFFu);
    if (act_hi) b[1] = (uint8_t)(hb >> 8);
  }
}

// ---- end of a pass: CRC verdict, pass counters, stop flag (sch.c:431-452) ---------------------------------------
B200_HD void finish_pass(const TdecView&   v,
                         bool              have_crc,
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
    bool ok        = have_crc && ((r.crc_lo16x2 & 0xFFFFu) == 0) && ((r.crc_hi8x2 & 0xFFu) == 0);
    s_lo.npass_run = (uint8_t)(pass_idx + 1);
    if (ok && !s_lo.crc_ok) {
      s_lo.crc_ok    = 1;
      s_lo.npass_crc = (uint8_t)(pass_idx + 1);
    }
    if (last || (ok && v.early_stop)) s_lo.active = 0;
    st[0] = s_lo;
  }
  if (act_hi) {
    bool ok        = have_crc && ((r.crc_lo16x2 >> 16) == 0) && (((r.crc_hi8x2 >> 16) & 0xFFu) == 0);
    s_hi.npass_run = (uint8_t)(pass_idx + 1);
    if (ok && !s_hi.crc_ok) {
      s_hi.crc_ok    = 1;
      s_hi.npass_crc = (uint8_t)(pass_idx + 1);
    }
    if (last || (ok && v.early_stop)) s_hi.active = 0;
    st[1] = s_hi;
  }
}

// ---- one full pass for one lane (two code blocks) straight from global memory --------------------------------------
// The CPU emulation (tests/emu) runs this; it is also the readable statement of the schedule.  The GPU kernel in
// tdec_kernels.cu runs the same window functions, two warps at a time, fed through shared-memory rings.
template <bool DEC2, bool FIRST, bool IN8>
B200_HD void load_window_direct(const TileDesc& td, int lane, int w, WinRegs& r, uint32_t pos[8])
{
  u4       s[2] = {}, p[2] = {};
  uint32_t e[8] = {};
  if (IN8) {
    p[0] = (DEC2 ? td.P18 : td.P08)[row8(w, lane)];
    if (!DEC2) s[0] = td.S8[row8(w, lane)];
  } else {
    const u4* P = DEC2 ? td.P1 : td.P0;
    p[0]        = P[vec_row(2 * w, lane)];
    p[1]        = P[vec_row(2 * w + 1, lane)];
    if (!DEC2) {
      s[0] = td.S[vec_row(2 * w, lane)];
      s[1] = td.S[vec_row(2 * w + 1, lane)];
    }
  }
  for (int t = 0; t < 8; t++) {
    pos[t] = DEC2 ? (uint32_t)td.qpp_fwd[8 * w + t] : (uint32_t)(8 * w + t);
    if (!FIRST) e[t] = td.E[e_idx((int)pos[t], lane)];
  }
  win_unpack<DEC2, FIRST, IN8>(r, s, p, e);
}

B200_HD void ck_load(const TileDesc& td, int lane, int w, uint32_t c[8])
{
  const u4 c0 = td.CK[ck_idx(w, 0, lane)], c1 = td.CK[ck_idx(w, 1, lane)];
  c[0] = c0.x; c[1] = c0.y; c[2] = c0.z; c[3] = c0.w; c[4] = c1.x; c[5] = c1.y; c[6] = c1.z; c[7] = c1.w;
}
B200_HD void ck_store(const TileDesc& td, int lane, int w, const uint32_t c[8])
{
  td.CK[ck_idx(w, 0, lane)] = u4{c[0], c[1], c[2], c[3]};
  td.CK[ck_idx(w, 1, lane)] = u4{c[4], c[5], c[6], c[7]};
}

// the three tail steps k = K+2, K+1, K: no a-priori, no normalisation (turbodecoder_gen.c:73-75,105)
template <bool IN8>
B200_HD void beta_tail(uint32_t B[8], const u4& st, const u4& pt)
{
  B[0] = 0;
#pragma unroll
  for (int i = 1; i < 8; i++) B[i] = NEG_INF2;
#pragma unroll
  for (int t = 2; t >= 0; t--) {
    const u4       sq[2] = {st, st}, pq[2] = {pt, pt};
    const uint32_t x = win_val<IN8>(sq, t), y = win_val<IN8>(pq, t);
    beta_step(B, x, y, add2(x, y));
  }
}

template <bool DEC2, bool FIRST, bool IN8>
B200_HD void siso_pass_lane(const TdecView& v, int tile, int lane, int pass_idx)
{
  const LaneMap lm = v.lanes[(size_t)tile * 32 + lane];
  if (lm.st0 == LANE_EMPTY) {
    return;
  }
  const TileDesc& td   = v.tiles[tile];
  CbStatus*       st   = v.status + lm.st0;
  CbStatus        s_lo = st[0];
  CbStatus        s_hi = st[1];
  bool            act_lo = s_lo.active != 0, act_hi = s_hi.active != 0;
  if (!act_lo && !act_hi) {
    return;
  }
  const int     K = (int)td.K, nw = K / 8, ws = tdec_split(K, v.split_percent);
  const CrcPow* cbase = DEC2 ? td.crc_perm : td.crc_nat;
  WinRegs       r;
  uint32_t      pos[8];

  // phase 1, warp F
  uint32_t A[8];
  A[0] = 0;
  for (int i = 1; i < 8; i++) A[i] = NEG_INF2;
  for (int w = 0; w < ws; w++) {
    ck_store(td, lane, w, A);
    load_window_direct<DEC2, FIRST, IN8>(td, lane, w, r, pos);
    alpha_window(A, r);
  }
  // phase 1, warp B
  uint32_t B[8];
  {
    // encoder 2's systematic tail is always int16 (S2T); widen an int8 parity tail to match by using IN8 on both
    u4 pt, stl;
    if (IN8) {
      pt  = (DEC2 ? td.P18 : td.P08)[row8(nw, lane)];
      stl = td.S8[row8(nw, lane)];
    } else {
      pt  = (DEC2 ? td.P1 : td.P0)[vec_row(K / 4, lane)];
      stl = td.S[vec_row(K / 4, lane)];
    }
    if (DEC2) {
      const u4 s2 = v.S2T[(size_t)tile * 32 + lane];
      B[0]        = 0;
      for (int i = 1; i < 8; i++) B[i] = NEG_INF2;
      for (int t = 2; t >= 0; t--) {
        const u4       pq[2] = {pt, pt};
        const uint32_t x = u4_get(s2, t), y = win_val<IN8>(pq, t);
        beta_step(B, x, y, add2(x, y));
      }
    } else {
      beta_tail<IN8>(B, stl, pt);
    }
  }
  ck_store(td, lane, nw - 1, B);
  for (int w = nw - 1; w >= ws; w--) {
    load_window_direct<DEC2, FIRST, IN8>(td, lane, w, r, pos);
    beta_window(B, r);
    if (w > ws) {
      ck_store(td, lane, w - 1, B);
      normalise(B);
    }
  }
  // phase 2, warp F
  LaneResult res = {0u, 0u};
  WinOut     o;
  uint32_t   c[8];
  for (int w = ws; w < nw; w++) {
    load_window_direct<DEC2, FIRST, IN8>(td, lane, w, r, pos);
    ck_load(td, lane, w, c);
    fwd_window(A, c, 8 * w + 8 < K, r, cbase ? cbase + 8 * w : nullptr, res, o);
    for (int t = 0; t < 8; t++) td.E[e_idx((int)pos[t], lane)] = o.enew[t];
    hb_store(&v.HB[hb_idx(lm.hb0, w)], o.bits, act_lo, act_hi);
  }
  // phase 2, warp B
  for (int w = ws - 1; w >= 0; w--) {
    load_window_direct<DEC2, FIRST, IN8>(td, lane, w, r, pos);
    ck_load(td, lane, w, c);
    bwd_window(B, c, r, cbase ? cbase + 8 * w : nullptr, res, o);
    for (int t = 0; t < 8; t++) td.E[e_idx((int)pos[t], lane)] = o.enew[t];
    hb_store(&v.HB[hb_idx(lm.hb0, w)], o.bits, act_lo, act_hi);
  }
  finish_pass(v, cbase != nullptr, st, s_lo, s_hi, act_lo, act_hi, res, pass_idx);
}

// ---- block-granular early stop: re-packing the lanes that still run into fewer tiles -------------------------------
// One group (equal K and CRC kind, consecutive tiles) is planned by one thread block; the steps below are written for
// `nthr` cooperating threads with `sync()` between them, so the CPU emulation runs the very same code with nthr = 1.
// Scratch (per tile, global memory): mask[t] = lanes of tile t holding a pair that still runs, pref[t] = running count
// (free lanes in receiver tiles / running lanes in donor tiles).  plan[g] = {receivers, first move, moves, go}.
struct GroupPlan {
  uint32_t receivers; // the group's running lanes fit its first `receivers` tiles
  uint32_t base;      // first entry of this group in the move list
  uint32_t moves;
  uint32_t go;
};

B200_HD bool lane_running(const TdecView& v, uint32_t slot)
{
  const LaneMap lm = v.lanes[slot];
  if (lm.st0 == LANE_EMPTY) return false;
  return (v.status[lm.st0].active | v.status[lm.st0 + 1].active) != 0;
}

// step 1: per-tile masks.  Returns false for a tile whose inputs are kept in int16 (such a group is left alone).
B200_HD bool compact_scan_tile(const TdecView& v, uint32_t tile, uint32_t* mask)
{
  uint32_t m = 0;
  for (uint32_t l = 0; l < 32; l++) {
    if (lane_running(v, tile * 32 + l)) m |= 1u << l;
  }
  mask[tile] = m;
  return v.fmt[tile] == 0u;
}

B200_HD uint32_t popc32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
  return (uint32_t)__popc(x);
#else
  return (uint32_t)__builtin_popcount(x);
#endif
}

// step 2 (one thread): decide whether re-packing pays and lay out the moves.  `min_gain_tiles`: tiles that must be
// freed for the copy to be worth its traffic (a lane's columns are ~60 KB at K=6144).
B200_HD void compact_plan_group(const TileGroup& g, const uint32_t* mask, uint32_t* pref, bool all_int8, uint32_t min_gain_tiles,
                                uint32_t* move_counter, uint32_t move_cap, GroupPlan& plan)
{
  uint32_t running = 0, tiles_now = 0;
  for (uint32_t t = 0; t < g.ntiles; t++) {
    const uint32_t n = popc32(mask[g.first_tile + t]);
    running += n;
    if (n) tiles_now = t + 1;
  }
  const uint32_t need = (running + 31) / 32;
  plan = GroupPlan{need, 0, 0, 0};
  if (!all_int8 || tiles_now < need + min_gain_tiles || 5 * need > 3 * tiles_now) return; // pays when the tiles shrink to 60 % or less
  uint32_t nfree = 0, nrun = 0;
  for (uint32_t t = 0; t < need; t++) {
    pref[g.first_tile + t] = nfree;
    nfree += 32 - popc32(mask[g.first_tile + t]);
  }
  for (uint32_t t = need; t < tiles_now; t++) {
    pref[g.first_tile + t] = nrun;
    nrun += popc32(mask[g.first_tile + t]);
  }
  for (uint32_t t = tiles_now; t < g.ntiles; t++) pref[g.first_tile + t] = nrun;
  if (nrun == 0 || nrun > nfree) return; // nrun <= nfree always holds (running <= 32 need); belt and braces
#if defined(__CUDA_ARCH__)
  const uint32_t base = atomicAdd(move_counter, nrun);
#else
  const uint32_t base = *move_counter;
  *move_counter += nrun;
#endif
  if (base + nrun > move_cap) return; // the list is sized for every lane, cannot happen
  plan.base  = base;
  plan.moves = nrun;
  plan.go    = 1;
}

// step 3: tile t of the group writes its side of the moves (receiver: destinations, donor: sources)
B200_HD void compact_emit_tile(const TileGroup& g, uint32_t t, const uint32_t* mask, const uint32_t* pref, const GroupPlan& plan, MoveRec* moves)
{
  if (!plan.go) return;
  const uint32_t tile = g.first_tile + t;
  uint32_t       k    = pref[tile];
  if (t < plan.receivers) {
    uint32_t fr = ~mask[tile];
    for (uint32_t l = 0; l < 32 && k < plan.moves; l++) {
      if ((fr >> l) & 1u) moves[plan.base + k++].dst = tile * 32 + l;
    }
  } else {
    const uint32_t m = mask[tile];
    for (uint32_t l = 0; l < 32; l++) {
      if ((m >> l) & 1u) moves[plan.base + k++].src = tile * 32 + l;
    }
  }
}

// step 4: the moved pair's name follows its data; the source slot holds nothing afterwards
B200_HD void compact_rename(const TdecView& v, const MoveRec& mv)
{
  v.lanes[mv.dst]     = v.lanes[mv.src];
  v.lanes[mv.src].st0 = LANE_EMPTY;
}

// the data of one move, element i of n = compact_move_elems(K): a lane's column of S8, P08, P18, E and its S2T entry
B200_HD uint32_t compact_move_elems(uint32_t K)
{
  return 3u * (K / 8u + 1u) + K + 1u;
}
B200_HD void compact_move_elem(const TdecView& v, const MoveRec& mv, uint32_t i)
{
  const uint32_t  ts = mv.src >> 5, ls = mv.src & 31u, td_ = mv.dst >> 5, ld = mv.dst & 31u;
  const TileDesc& a = v.tiles[ts];
  const TileDesc& b = v.tiles[td_];
  const uint32_t  r8 = a.K / 8u + 1u;
  if (i < 3u * r8) {
    const uint32_t s = i / r8, w = i % r8;
    const u4*      src = s == 0 ? a.S8 : (s == 1 ? a.P08 : a.P18);
    u4*            dst = s == 0 ? b.S8 : (s == 1 ? b.P08 : b.P18);
    dst[row8((int)w, (int)ld)] = src[row8((int)w, (int)ls)];
  } else if (i < 3u * r8 + a.K) {
    const uint32_t k = i - 3u * r8;
    b.E[e_idx((int)k, (int)ld)] = a.E[e_idx((int)k, (int)ls)];
  } else {
    v.S2T[(size_t)td_ * 32 + ld] = v.S2T[(size_t)ts * 32 + ls];
  }
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

// Decided byte jb (bits 8jb..8jb+7, MSB first, natural order) of block c (0..63) of tile `tile` after its last pass.
// rev = inverse QPP table (tc_interl_lte.c:93); after an odd pass HB is in DEC2's visiting order and bit j sits at
// visiting index rev[j] (the reference instead de-interleaves the whole LLR vector, turbodecoder_iter.h:127).
// Status and HB are addressed by the block's HOME slot, wherever its data was moved in between.
B200_HD uint8_t decide_byte(const TdecView& v, int tile, int c, int jb)
{
  const TileDesc& td   = v.tiles[tile];
  const int       lane = c >> 1, half = c & 1;
  const LaneMap   home = lane_home(td, tile, lane);
  const CbStatus  st   = v.status[home.st0 + half];
  const bool      perm = st.npass_run > 0 && ((st.npass_run - 1) & 1);
  if (!perm) {
    uint16_t hb = v.HB[hb_idx(home.hb0, jb)];
    return (uint8_t)(half ? (hb >> 8) : (hb & 0xFF));
  }
  uint32_t byte = 0;
  for (int t = 0; t < 8; t++) {
    int      i  = td.qpp_rev[8 * jb + t];
    uint16_t hb = v.HB[hb_idx(home.hb0, i >> 3)];
    uint32_t b  = half ? (hb >> 8) : (hb & 0xFF);
    byte        = (byte << 1) | ((b >> (7 - (i & 7))) & 1u);
  }
  return (uint8_t)byte;
}

} // namespace b200
