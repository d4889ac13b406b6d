// CPU emulation of the CUDA turbo-decoder data path: runs the SAME per-lane code the kernels run
// (srslte_b200/csrc/tdec_core.h, compiled here by g++ with the host versions of the packed int16 ops) lane by lane.
// Lets `pytest -m "not gpu"` check layout, recursion, extrinsic exchange, CRC syndrome and early-stop logic against
// the oracle on a machine without a GPU.  Test-only; never shipped.
#include <stdint.h>
#include <string.h>
#include <vector>

#include "lte_tables.h"
#include "tdec_core.h"

using namespace b200;

extern "C" int emu_tdec_batch2(const int16_t* llr,
                               uint32_t       ncb,
                               uint32_t       K,
                               uint32_t       max_pass,
                               int            crc_kind, /* 0 = CRC24B, 1 = CRC24A, 2 = none */
                               int            early_stop,
                               int            force_int16, /* never use the int8 tile format */
                               int            split_percent,
                               uint8_t*       out,
                               uint8_t*       crc_ok,
                               uint8_t*       npass_crc,
                               uint8_t*       npass_run)
{
  int cbi = cb_index_exact(K);
  if (cbi < 0) return -1;
  const int    ntiles = (int)((ncb + TDEC_TILE_CB - 1) / TDEC_TILE_CB);
  const size_t vrows  = (size_t)ntiles * ((K + 4) / 4) * 32;
  const size_t rows8  = (size_t)ntiles * (K / 8 + 1) * 32;
  const u4     junk   = {0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu}; // the unused format must never be read
  std::vector<u4>       S(vrows, junk), P0(vrows, junk), P1(vrows, junk), S2T((size_t)ntiles * 32), CK((size_t)ntiles * (K / 8) * 2 * 32);
  std::vector<u4>       S8(rows8, junk), P08(rows8, junk), P18(rows8, junk);
  std::vector<uint32_t> fmt(ntiles, 0);
  std::vector<uint32_t> E((size_t)ntiles * K * 32, 0xDEADBEEF); // garbage on purpose: pass 0 must not read it
  std::vector<uint16_t> HB((size_t)ntiles * (K / 8) * 32, 0);
  std::vector<CbStatus> st((size_t)ntiles * TDEC_TILE_CB);
  std::vector<uint16_t> fwd, rev;
  std::vector<CrcPow>   cnat, cperm;
  qpp_tables(cbi, fwd, rev);
  if (crc_kind != 2) crc_visit_tables(crc_kind == 1 ? CRC24A_POLY : CRC24B_POLY, cbi, cnat, cperm);

  TdecView v;
  v.K = (int)K; v.ntiles = ntiles; v.ws = tdec_split((int)K, split_percent);
  v.S = S.data(); v.P0 = P0.data(); v.P1 = P1.data(); v.S2T = S2T.data();
  v.S8 = S8.data(); v.P08 = P08.data(); v.P18 = P18.data(); v.fmt = fmt.data();
  v.E = E.data(); v.CK = CK.data(); v.HB = HB.data(); v.status = st.data();
  v.qpp_fwd = fwd.data(); v.crc_nat = crc_kind != 2 ? cnat.data() : nullptr; v.crc_perm = crc_kind != 2 ? cperm.data() : nullptr;
  v.early_stop = early_stop; v.max_pass = (int)max_pass;

  const size_t nllr = 3 * (size_t)K + 12;
  std::vector<int16_t> zeros(nllr, 0);
  for (int tile = 0; tile < ntiles; tile++) {
    bool fits = !force_int16;
    for (uint32_t c = 0; c < TDEC_TILE_CB && fits; c++) {
      const uint32_t cb = (uint32_t)tile * TDEC_TILE_CB + c;
      if (cb >= ncb) break;
      for (size_t i = 0; i < nllr; i++) {
        if (llr[cb * nllr + i] < -128 || llr[cb * nllr + i] > 127) {
          // encoder 2's systematic tail lives in S2T (int16) and may be anything
          const size_t t = i - 3 * (size_t)K;
          if (i >= 3 * (size_t)K + 6 && (t & 1) == 0) continue;
          fits = false;
          break;
        }
      }
    }
    fmt[tile] = fits ? 0 : 1;
    for (int lane = 0; lane < 32; lane++) {
      uint32_t       cb0 = (uint32_t)tile * TDEC_TILE_CB + 2 * lane, cb1 = cb0 + 1;
      const int16_t* a   = cb0 < ncb ? llr + cb0 * nllr : zeros.data();
      const int16_t* b   = cb1 < ncb ? llr + cb1 * nllr : zeros.data();
      if (!fits) {
        for (int k4 = 0; k4 < (int)(K + 4) / 4; k4++) {
          uint32_t w[3][4];
          for (int s = 0; s < 3; s++)
            for (int t = 0; t < 4; t++) w[s][t] = pack2(natural_pick(a, K, s, 4 * k4 + t), natural_pick(b, K, s, 4 * k4 + t));
          S[vec_row(v, tile, k4, lane)]  = u4{w[0][0], w[0][1], w[0][2], w[0][3]};
          P0[vec_row(v, tile, k4, lane)] = u4{w[1][0], w[1][1], w[1][2], w[1][3]};
          P1[vec_row(v, tile, k4, lane)] = u4{w[2][0], w[2][1], w[2][2], w[2][3]};
        }
      } else {
        for (int w8 = 0; w8 <= (int)K / 8; w8++) {
          uint32_t w[3][4];
          for (int s = 0; s < 3; s++)
            for (int q = 0; q < 4; q++) {
              uint32_t word = 0;
              for (int j = 0; j < 2; j++) {
                const int k = 8 * w8 + 2 * q + j;
                const uint8_t lo = (uint8_t)(int8_t)natural_pick(a, K, s, k), hi = (uint8_t)(int8_t)natural_pick(b, K, s, k);
                word |= ((uint32_t)lo | ((uint32_t)hi << 8)) << (16 * j);
              }
              w[s][q] = word;
            }
          S8[row8(v, tile, w8, lane)]  = u4{w[0][0], w[0][1], w[0][2], w[0][3]};
          P08[row8(v, tile, w8, lane)] = u4{w[1][0], w[1][1], w[1][2], w[1][3]};
          P18[row8(v, tile, w8, lane)] = u4{w[2][0], w[2][1], w[2][2], w[2][3]};
        }
      }
      uint32_t t3[4];
      for (int t = 0; t < 4; t++) t3[t] = pack2(natural_pick(a, K, 3, K + t), natural_pick(b, K, 3, K + t));
      S2T[(size_t)tile * 32 + lane] = u4{t3[0], t3[1], t3[2], t3[3]};
      st[cb0] = CbStatus{(uint8_t)(cb0 < ncb), 0, 0, 0};
      st[cb1] = CbStatus{(uint8_t)(cb1 < ncb), 0, 0, 0};
    }
  }
  for (uint32_t p = 0; p < max_pass; p++) {
    for (int tile = 0; tile < ntiles; tile++) {
      for (int lane = 0; lane < 32; lane++) {
        if (fmt[tile] == 0) {
          if (p == 0) {
            siso_pass_lane<false, true, true>(v, tile, lane, (int)p);
          } else if (p & 1) {
            siso_pass_lane<true, false, true>(v, tile, lane, (int)p);
          } else {
            siso_pass_lane<false, false, true>(v, tile, lane, (int)p);
          }
        } else {
          if (p == 0) {
            siso_pass_lane<false, true, false>(v, tile, lane, (int)p);
          } else if (p & 1) {
            siso_pass_lane<true, false, false>(v, tile, lane, (int)p);
          } else {
            siso_pass_lane<false, false, false>(v, tile, lane, (int)p);
          }
        }
      }
    }
  }
  for (uint32_t cb = 0; cb < ncb; cb++) {
    for (uint32_t jb = 0; jb < K / 8; jb++) out[(size_t)cb * (K / 8) + jb] = decide_byte(v, rev.data(), (int)cb, (int)jb);
    crc_ok[cb]    = st[cb].crc_ok;
    npass_crc[cb] = st[cb].npass_crc;
    npass_run[cb] = st[cb].npass_run;
  }
  return 0;
}

extern "C" int emu_tdec_batch(const int16_t* llr, uint32_t ncb, uint32_t K, uint32_t max_pass, int crc_kind, int early_stop,
                              uint8_t* out, uint8_t* crc_ok, uint8_t* npass_crc, uint8_t* npass_run)
{
  return emu_tdec_batch2(llr, ncb, K, max_pass, crc_kind, early_stop, 0, 47, out, crc_ok, npass_crc, npass_run);
}
