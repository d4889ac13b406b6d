// CPU emulation of the CUDA turbo-decoder data path: runs the SAME per-lane code the kernels run
// (srslte_b200/csrc/tdec_core.h, compiled here by g++ with the host versions of the packed int16 ops) lane by lane.
// Lets `pytest -m "not gpu"` check layout, recursion, extrinsic exchange, CRC syndrome and early-stop logic against
// the oracle on a machine without a GPU.  Test-only; never shipped.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "lte_tables.h"
#include "tdec_core.h"

using namespace b200;

// n_groups groups of equal-K blocks (llr, out, crc_ok, ...: group after group, as srsran_b200_tdec_run_mixed takes them).
// compact != 0 runs the lane re-packing of tdec_core.h (compact_*) after every pass, the way the device code does;
// *moves_done (optional) returns how many lanes were moved in total.
extern "C" int emu_tdec_mixed(const int16_t*  llr,
                              uint32_t        n_groups,
                              const uint32_t* Ks,
                              const uint32_t* ncbs,
                              uint32_t        max_pass,
                              int             crc_kind, /* 0 = CRC24B, 1 = CRC24A, 2 = none */
                              int             early_stop,
                              int             force_int16, /* never use the int8 tile format */
                              int             split_percent,
                              int             compact,
                              uint8_t*        out,
                              uint8_t*        crc_ok,
                              uint8_t*        npass_crc,
                              uint8_t*        npass_run,
                              uint32_t*       moves_done)
{
  struct GroupMem {
    std::vector<uint16_t> fwd, rev;
    std::vector<CrcPow>   cnat, cperm;
  };
  struct TileMem {
    std::vector<u4>       S, P0, P1, S8, P08, P18, CK;
    std::vector<uint32_t> E;
  };
  // tiles in order of descending K, like TdecEngine::prepare
  std::vector<uint32_t> order(n_groups);
  for (uint32_t i = 0; i < n_groups; i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return Ks[a] > Ks[b]; });
  std::vector<uint64_t> g_llr(n_groups), g_out(n_groups);
  std::vector<uint32_t> g_cb0(n_groups);
  uint64_t              lo = 0, oo = 0;
  uint32_t              c0 = 0;
  for (uint32_t i = 0; i < n_groups; i++) {
    if (cb_index_exact(Ks[i]) < 0) return -1;
    g_llr[i] = lo; g_out[i] = oo; g_cb0[i] = c0;
    lo += (uint64_t)ncbs[i] * (3ull * Ks[i] + 12); oo += (uint64_t)ncbs[i] * (Ks[i] / 8); c0 += ncbs[i];
  }
  size_t ntiles = 0, hb_rows = 0;
  for (uint32_t i = 0; i < n_groups; i++) {
    ntiles += (ncbs[i] + TDEC_TILE_CB - 1) / TDEC_TILE_CB;
    hb_rows += (size_t)((ncbs[i] + TDEC_TILE_CB - 1) / TDEC_TILE_CB) * (Ks[i] / 8);
  }
  const u4 junk = {0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu}; // the unused format must never be read
  std::vector<GroupMem>  gm(n_groups);
  std::vector<TileMem>   tm(ntiles);
  std::vector<TileDesc>  tiles(ntiles);
  std::vector<TileGroup> groups;
  std::vector<uint32_t>  fmt(ntiles, 0), mask(ntiles, 0), pref(ntiles, 0);
  std::vector<u4>        S2T(ntiles * 32);
  std::vector<LaneMap>   lanes(ntiles * 32);
  std::vector<uint16_t>  HB(hb_rows * 32, 0);
  std::vector<CbStatus>  st(ntiles * TDEC_TILE_CB);
  std::vector<MoveRec>   moves(ntiles * 32);
  uint32_t tile = 0, hb_row = 0;
  for (uint32_t oi = 0; oi < n_groups; oi++) {
    const uint32_t g = order[oi], K = Ks[g], nt = (ncbs[g] + TDEC_TILE_CB - 1) / TDEC_TILE_CB;
    if (nt == 0) { groups.push_back(TileGroup{tile, 0}); continue; }
    const int cbi = cb_index_exact(K);
    qpp_tables(cbi, gm[g].fwd, gm[g].rev);
    if (crc_kind != 2) crc_visit_tables(crc_kind == 1 ? CRC24A_POLY : CRC24B_POLY, cbi, gm[g].cnat, gm[g].cperm);
    groups.push_back(TileGroup{tile, nt});
    for (uint32_t t = 0; t < nt; t++, tile++) {
      TileMem& m = tm[tile];
      m.S.assign((size_t)((K + 4) / 4) * 32, junk); m.P0 = m.S; m.P1 = m.S;
      m.S8.assign((size_t)(K / 8 + 1) * 32, junk); m.P08 = m.S8; m.P18 = m.S8;
      m.CK.resize((size_t)(K / 8) * 2 * 32);
      m.E.assign((size_t)K * 32, 0xDEADBEEF); // garbage on purpose: pass 0 must not read it
      TileDesc& d = tiles[tile];
      d.S8 = m.S8.data(); d.P08 = m.P08.data(); d.P18 = m.P18.data(); d.S = m.S.data(); d.P0 = m.P0.data(); d.P1 = m.P1.data();
      d.E = m.E.data(); d.CK = m.CK.data(); d.qpp_fwd = gm[g].fwd.data(); d.qpp_rev = gm[g].rev.data();
      d.crc_nat = crc_kind != 2 ? gm[g].cnat.data() : nullptr; d.crc_perm = crc_kind != 2 ? gm[g].cperm.data() : nullptr;
      d.llr_off = g_llr[g] + (uint64_t)t * TDEC_TILE_CB * (3ull * K + 12); d.out_off = g_out[g] + (uint64_t)t * TDEC_TILE_CB * (K / 8);
      d.K = K; d.hb_row0 = hb_row; d.cb0 = g_cb0[g] + t * TDEC_TILE_CB; d.nblk = std::min<uint32_t>(TDEC_TILE_CB, ncbs[g] - t * TDEC_TILE_CB);
      d.group = oi;
      hb_row += K / 8;
    }
  }
  TdecView v;
  v.ntiles = (int)ntiles; v.split_percent = split_percent; v.tiles = tiles.data(); v.fmt = fmt.data(); v.S2T = S2T.data();
  v.lanes = lanes.data(); v.HB = HB.data(); v.status = st.data(); v.early_stop = early_stop; v.max_pass = (int)max_pass;

  // natural -> tiled (what the load kernels do)
  std::vector<int16_t> zeros(3 * 6144 + 12, 0);
  for (int tl = 0; tl < (int)ntiles; tl++) {
    const TileDesc& d = tiles[tl];
    const uint32_t  K = d.K;
    const size_t    nllr = 3 * (size_t)K + 12;
    auto src = [&](uint32_t c) { return c < d.nblk ? llr + d.llr_off + (size_t)c * nllr : (const int16_t*)nullptr; };
    bool fits = !force_int16;
    for (uint32_t c = 0; c < d.nblk && fits; c++) {
      const int16_t* a = src(c);
      for (size_t i = 0; i < nllr; i++) {
        if (a[i] < -128 || a[i] > 127) {
          // encoder 2's systematic tail lives in S2T (int16) and may be anything
          const size_t t = i - 3 * (size_t)K;
          if (i >= 3 * (size_t)K + 6 && (t & 1) == 0) continue;
          fits = false;
          break;
        }
      }
    }
    fmt[tl] = fits ? 0 : 1;
    for (int lane = 0; lane < 32; lane++) {
      const int16_t* a = src(2 * lane) ? src(2 * lane) : zeros.data();
      const int16_t* b = src(2 * lane + 1) ? src(2 * lane + 1) : zeros.data();
      if (!fits) {
        for (int k4 = 0; k4 < (int)(K + 4) / 4; k4++) {
          uint32_t w[3][4];
          for (int s = 0; s < 3; s++)
            for (int t = 0; t < 4; t++) w[s][t] = pack2(natural_pick(a, K, s, 4 * k4 + t), natural_pick(b, K, s, 4 * k4 + t));
          d.S[vec_row(k4, lane)]  = u4{w[0][0], w[0][1], w[0][2], w[0][3]};
          d.P0[vec_row(k4, lane)] = u4{w[1][0], w[1][1], w[1][2], w[1][3]};
          d.P1[vec_row(k4, lane)] = u4{w[2][0], w[2][1], w[2][2], w[2][3]};
        }
      } else {
        for (int w8 = 0; w8 <= (int)K / 8; w8++) {
          uint32_t w[3][4];
          for (int s = 0; s < 3; s++)
            for (int q = 0; q < 4; q++) {
              uint32_t word = 0;
              for (int j = 0; j < 2; j++) {
                const int k = 8 * w8 + 2 * q + j;
                const uint8_t lo8 = (uint8_t)(int8_t)natural_pick(a, K, s, k), hi8 = (uint8_t)(int8_t)natural_pick(b, K, s, k);
                word |= ((uint32_t)lo8 | ((uint32_t)hi8 << 8)) << (16 * j);
              }
              w[s][q] = word;
            }
          d.S8[row8(w8, lane)]  = u4{w[0][0], w[0][1], w[0][2], w[0][3]};
          d.P08[row8(w8, lane)] = u4{w[1][0], w[1][1], w[1][2], w[1][3]};
          d.P18[row8(w8, lane)] = u4{w[2][0], w[2][1], w[2][2], w[2][3]};
        }
      }
      uint32_t t3[4];
      for (int t = 0; t < 4; t++) t3[t] = pack2(natural_pick(a, K, 3, K + t), natural_pick(b, K, 3, K + t));
      S2T[(size_t)tl * 32 + lane] = u4{t3[0], t3[1], t3[2], t3[3]};
      const LaneMap home = lane_home(d, tl, lane);
      lanes[(size_t)tl * 32 + lane] = home;
      st[home.st0]     = CbStatus{(uint8_t)(src(2 * lane) != nullptr), 0, 0, 0};
      st[home.st0 + 1] = CbStatus{(uint8_t)(src(2 * lane + 1) != nullptr), 0, 0, 0};
    }
  }
  uint32_t total_moves = 0;
  for (uint32_t p = 0; p < max_pass; p++) {
    for (int tl = 0; tl < (int)ntiles; tl++) {
      for (int lane = 0; lane < 32; lane++) {
        if (fmt[tl] == 0) {
          if (p == 0) {
            siso_pass_lane<false, true, true>(v, tl, lane, (int)p);
          } else if (p & 1) {
            siso_pass_lane<true, false, true>(v, tl, lane, (int)p);
          } else {
            siso_pass_lane<false, false, true>(v, tl, lane, (int)p);
          }
        } else {
          if (p == 0) {
            siso_pass_lane<false, true, false>(v, tl, lane, (int)p);
          } else if (p & 1) {
            siso_pass_lane<true, false, false>(v, tl, lane, (int)p);
          } else {
            siso_pass_lane<false, false, false>(v, tl, lane, (int)p);
          }
        }
      }
    }
    if (compact && early_stop && p + 1 < max_pass) {
      // the steps of tdec_compact_plan_kernel / tdec_compact_move_kernel, one "thread"
      uint32_t counter = 0;
      for (const TileGroup& g : groups) {
        bool all8 = true;
        for (uint32_t t = 0; t < g.ntiles; t++) all8 = compact_scan_tile(v, g.first_tile + t, mask.data()) && all8;
        GroupPlan plan;
        compact_plan_group(g, mask.data(), pref.data(), all8, (uint32_t)(compact - 1), &counter, (uint32_t)moves.size(), plan);
        for (uint32_t t = 0; t < g.ntiles; t++) compact_emit_tile(g, t, mask.data(), pref.data(), plan, moves.data());
      }
      for (uint32_t m = 0; m < counter; m++) {
        const uint32_t ne = compact_move_elems(tiles[moves[m].src >> 5].K);
        for (uint32_t i = 0; i < ne; i++) compact_move_elem(v, moves[m], i);
        compact_rename(v, moves[m]);
      }
      total_moves += counter;
    }
  }
  if (moves_done) *moves_done = total_moves;
  for (int tl = 0; tl < (int)ntiles; tl++) {
    const TileDesc& d = tiles[tl];
    for (uint32_t c = 0; c < d.nblk; c++) {
      for (uint32_t jb = 0; jb < d.K / 8; jb++) out[d.out_off + (size_t)c * (d.K / 8) + jb] = decide_byte(v, tl, (int)c, (int)jb);
      const CbStatus s = st[(size_t)tl * TDEC_TILE_CB + c];
      crc_ok[d.cb0 + c]    = s.crc_ok;
      npass_crc[d.cb0 + c] = s.npass_crc;
      npass_run[d.cb0 + c] = s.npass_run;
    }
  }
  return 0;
}

extern "C" int emu_tdec_batch2(const int16_t* llr, uint32_t ncb, uint32_t K, uint32_t max_pass, int crc_kind, int early_stop,
                               int force_int16, int split_percent, uint8_t* out, uint8_t* crc_ok, uint8_t* npass_crc, uint8_t* npass_run)
{
  if (ncb == 0) return cb_index_exact(K) < 0 ? -1 : 0;
  return emu_tdec_mixed(llr, 1, &K, &ncb, max_pass, crc_kind, early_stop, force_int16, split_percent, 0, out, crc_ok, npass_crc, npass_run,
                        nullptr);
}

extern "C" int emu_tdec_batch(const int16_t* llr, uint32_t ncb, uint32_t K, uint32_t max_pass, int crc_kind, int early_stop,
                              uint8_t* out, uint8_t* crc_ok, uint8_t* npass_crc, uint8_t* npass_run)
{
  return emu_tdec_batch2(llr, ncb, K, max_pass, crc_kind, early_stop, 0, 47, out, crc_ok, npass_crc, npass_run);
}
