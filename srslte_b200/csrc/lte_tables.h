// Host-side generators for the integer tables the kernels consume.  Header-only, plain C++ (no CUDA), so the
// CUDA library and the CPU-only emulation tests share one definition.
//
// Restated from the reference (file:line relative to the reference root):
//   code-block size table + index lookup ......... lib/src/phy/fec/cbsegm.c:32-43,119-151
//   QPP interleaver PI(i) = (f1 i + f2 i^2) mod K .. lib/src/phy/fec/turbo/tc_interl_lte.c:39-59,69-94
//   rate-matching de-interleaver ................. lib/src/phy/fec/turbo/rm_turbo.c:70-71,175-248
//   CRC24A / CRC24B polynomials .................. lib/include/srsran/phy/common/phy_common.h:72-73, crc.c:30-46
#pragma once
#include <stdint.h>
#include <vector>

#include "tdec_core.h"

namespace b200 {

constexpr int NOF_CB_SIZES = 188;
constexpr int MAX_CB_LEN   = 6144;

struct QppRow {
  uint16_t K, f1, f2;
};

// TS 36.212 Table 5.1.3-3
static const QppRow g_qpp_rows[NOF_CB_SIZES] = {
#include "qpp_table.inc"
};

inline int cb_size(uint32_t idx)
{
  return idx < (uint32_t)NOF_CB_SIZES ? (int)g_qpp_rows[idx].K : -1;
}

// index of the first size >= K, -1 past the end (cbsegm.c:119-130)
inline int cb_index(uint32_t K)
{
  int lo = 0, hi = NOF_CB_SIZES;
  while (lo < hi) {
    int mid = (lo + hi) / 2;
    if (g_qpp_rows[mid].K < K) {
      lo = mid + 1;
    } else {
      hi = mid;
    }
  }
  return lo == NOF_CB_SIZES ? -1 : lo;
}

// exact-size lookup
inline int cb_index_exact(uint32_t K)
{
  int i = cb_index(K);
  return (i >= 0 && g_qpp_rows[i].K == K) ? i : -1;
}

// 36.212 5.1.2 code block segmentation with the reference's conventions (cbsegm.c:48-117): B = tbs + 24,
// Z = 6144, C = ceil(B / (Z - 24)) when B > Z, K1 = smallest table size >= B'/C, K2 the next smaller size.
struct CbSegm {
  uint32_t F, C, K1, K2, K1_idx, K2_idx, C1, C2, tbs;
};

inline int cb_segmentation(uint32_t tbs, CbSegm& s)
{
  s = CbSegm{0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (tbs == 0) {
    return 0;
  }
  const uint32_t B = tbs + 24;
  uint32_t       Bp;
  s.tbs = tbs;
  if (B <= (uint32_t)MAX_CB_LEN) {
    s.C = 1;
    Bp  = B;
  } else {
    s.C = (B + (MAX_CB_LEN - 24) - 1) / (MAX_CB_LEN - 24);
    Bp  = B + 24 * s.C;
  }
  const int i1 = cb_index((Bp - 1) / s.C + 1);
  if (i1 < 0) {
    return -1;
  }
  s.K1     = g_qpp_rows[i1].K;
  s.K1_idx = (uint32_t)i1;
  if (s.C == 1) {
    s.C1 = 1;
  } else {
    s.K2_idx = i1 > 0 ? (uint32_t)i1 - 1 : 0;
    s.K2     = g_qpp_rows[s.K2_idx].K;
    s.C2     = (s.K1 != s.K2) ? (s.C * s.K1 - Bp) / (s.K1 - s.K2) : 0;
    s.C1     = s.C - s.C2;
  }
  s.F = s.C1 * s.K1 + s.C2 * s.K2 - Bp;
  return 0;
}

inline void qpp_tables(int cb_idx, std::vector<uint16_t>& fwd, std::vector<uint16_t>& rev)
{
  const uint64_t K = g_qpp_rows[cb_idx].K, f1 = g_qpp_rows[cb_idx].f1, f2 = g_qpp_rows[cb_idx].f2;
  fwd.assign(K, 0);
  rev.assign(K, 0);
  // PI(i+1) - PI(i) = f1 + f2 (2i + 1): keep everything reduced mod K, no 64-bit products needed
  uint64_t p = 0, g = (f1 + f2) % K, step = (2 * f2) % K;
  for (uint64_t i = 0; i < K; i++) {
    fwd[i] = (uint16_t)p;
    rev[p] = (uint16_t)i;
    p      = (p + g) % K;
    g      = (g + step) % K;
  }
}

constexpr uint32_t CRC24A_POLY = 0x1864CFB;
constexpr uint32_t CRC24B_POLY = 0x1800063;

// pow[m] = x^m mod g(x) (24-bit), duplicated into both int16 halves the way the SISO kernel consumes it
inline void crc_pow_table(uint32_t poly, int n, std::vector<CrcPow>& out)
{
  out.resize(n);
  uint32_t r = 1;
  for (int m = 0; m < n; m++) {
    out[m].lo16x2 = (r & 0xFFFFu) * 0x00010001u;
    out[m].hi8x2  = ((r >> 16) & 0xFFu) * 0x00010001u;
    r <<= 1;
    if (r & 0x1000000u) {
      r ^= poly;
    }
    r &= 0xFFFFFFu;
  }
}

// Syndrome weights in the order each constituent decoder visits the bits (see TdecView::crc_nat / crc_perm)
inline void crc_visit_tables(uint32_t poly, int cb_idx, std::vector<CrcPow>& nat, std::vector<CrcPow>& perm)
{
  const int             K = g_qpp_rows[cb_idx].K;
  std::vector<CrcPow>   pw;
  std::vector<uint16_t> fwd, rev;
  crc_pow_table(poly, K, pw);
  qpp_tables(cb_idx, fwd, rev);
  nat.resize(K);
  perm.resize(K);
  for (int j = 0; j < K; j++) {
    nat[j]  = pw[K - 1 - j];
    perm[j] = pw[K - 1 - fwd[j]];
  }
}

// ---- rate matching geometry (36.212 5.1.4.1, rm_turbo.c:175-248) -------------------------------------------------
static const uint8_t g_rm_colperm[32] = {0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30,
                                         1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31};

struct RmGeom {
  int K, R, Kpi, Nd, Ncb;
  int k0[4];
};

inline RmGeom rm_geom(int K)
{
  RmGeom g;
  g.K   = K;
  g.R   = (K + 4 - 1) / 32 + 1;
  g.Kpi = 32 * g.R;
  g.Nd  = g.Kpi - (K + 4);
  g.Ncb = 3 * g.Kpi;
  for (int rv = 0; rv < 4; rv++) {
    int c    = (g.Ncb + 8 * g.R - 1) / (8 * g.R);
    g.k0[rv] = g.R * (2 * c * rv + 2);
  }
  return g;
}

// circular-buffer position -> natural index 3*bit+stream (bit in 0..K+3), -1 for a dummy position
inline int rm_pos_to_natural(const RmGeom& g, int p)
{
  int stream, q;
  if (p < g.Kpi) {
    stream = 0;
    q      = p;
  } else {
    stream = 1 + ((p - g.Kpi) & 1);
    q      = (p - g.Kpi) >> 1;
  }
  const int col = q / g.R, row = q % g.R;
  int       e   = (stream < 2) ? row * 32 + g_rm_colperm[col] : (g_rm_colperm[col] + 32 * row + 1) % g.Kpi;
  return e < g.Nd ? -1 : 3 * (e - g.Nd) + stream;
}

// scatter form, as the reference stores it: table[i] = natural destination of the i-th received value
inline void rm_scatter_table(int cb_idx, int rv, std::vector<uint16_t>& table)
{
  const RmGeom g = rm_geom(g_qpp_rows[cb_idx].K);
  const int    n = 3 * g.K + 12;
  table.resize(n);
  int i = 0;
  for (int j = 0; i < n; j++) {
    int d = rm_pos_to_natural(g, (g.k0[rv] + j) % g.Ncb);
    if (d >= 0) {
      table[i++] = (uint16_t)d;
    }
  }
}

// gather form used by the kernels: inv[d] = index i (0..n-1) of the received value that lands on natural index d
inline void rm_gather_table(int cb_idx, int rv, std::vector<uint16_t>& inv)
{
  std::vector<uint16_t> t;
  rm_scatter_table(cb_idx, rv, t);
  inv.assign(t.size(), 0);
  for (size_t i = 0; i < t.size(); i++) inv[t[i]] = (uint16_t)i;
}

} // namespace b200
