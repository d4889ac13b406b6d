/*
 * TEST INFRASTRUCTURE ONLY — plain-C CPU restatement of the reference's receive-side hot path.
 *
 * Purpose: the checker for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg on machines where the
 * reference sources are not available.  The product library (srslte_b200/csrc) never includes, links or loads
 * this file.  Every function cites the reference lines whose behaviour it restates (paths relative to the
 * reference root).  Parity status: PINNED — tests/test_oracle_vs_reference.py checks each function below against
 * the reference's own code (oracle/_ref/libsrsref.so, built from the unmodified sources) and tests/golden/ holds
 * vectors generated from that build (tools/gen_golden.py).
 *
 * Arithmetic conventions restated from the reference:
 *   - LLRs are int16; every add/sub in the decoder and the de-matcher wraps modulo 2^16
 *     (lib/src/phy/fec/turbo/turbodecoder_gen.c:58-198 uses plain int16_t, vector_simd.c:132 _mm_sub_epi16).
 *   - LLR > 0 decides bit 1 (turbodecoder_gen.c:266).
 */
#include <complex.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define ORC_NOF_K 188
#define ORC_MAX_K 6144
#define ORC_INF 10000 /* turbodecoder_gen.c:37 */

typedef struct {
  uint16_t K, f1, f2;
} qpp_row_t;

/* TS 36.212 Table 5.1.3-3 (same data as cbsegm.c:32-43 and tc_interl_lte.c:39-59) */
static const qpp_row_t qpp_rows[ORC_NOF_K] = {
#include "qpp_table.inc"
};

static inline int16_t w16(int v)
{
  return (int16_t)(uint16_t)(unsigned)v; /* two's-complement wrap, what the reference's int16_t stores do */
}

/* ---------------------------------------------------------------- code-block sizes (cbsegm.c:119-151) */

int orc_nof_cb_sizes(void)
{
  return ORC_NOF_K;
}

int orc_cbsize(uint32_t idx)
{
  return idx < ORC_NOF_K ? (int)qpp_rows[idx].K : -1;
}

/* first table entry >= K (cbsegm.c:119-130) */
int orc_cbindex(uint32_t K)
{
  for (int j = 0; j < ORC_NOF_K; j++) {
    if (qpp_rows[j].K >= K) {
      return j;
    }
  }
  return -1;
}

/* 36.212 5.1.2 segmentation as computed by cbsegm.c:62-117; out = F C K1 K2 K1_idx K2_idx C1 C2 tbs */
int orc_cbsegm(uint32_t tbs, uint32_t* out)
{
  memset(out, 0, 9 * sizeof(uint32_t));
  if (tbs == 0) {
    return 0;
  }
  uint32_t B = tbs + 24, C, Bp;
  if (B <= ORC_MAX_K) {
    C  = 1;
    Bp = B;
  } else {
    C  = (B + (ORC_MAX_K - 24) - 1) / (ORC_MAX_K - 24);
    Bp = B + 24 * C;
  }
  int idx1 = orc_cbindex((Bp - 1) / C + 1);
  if (idx1 < 0) {
    return -1;
  }
  uint32_t K1 = qpp_rows[idx1].K, K2 = 0, K2i = 0, C1 = 1, C2 = 0;
  if (C > 1) {
    K2i = idx1 > 0 ? (uint32_t)idx1 - 1 : 0;
    K2  = idx1 > 0 ? qpp_rows[idx1 - 1].K : K1;
    C2  = (K1 != K2) ? (C * K1 - Bp) / (K1 - K2) : 0;
    C1  = C - C2;
  }
  out[0] = C1 * K1 + C2 * K2 - Bp;
  out[1] = C;
  out[2] = K1;
  out[3] = K2;
  out[4] = (uint32_t)idx1;
  out[5] = K2i;
  out[6] = C1;
  out[7] = C2;
  out[8] = tbs;
  return 0;
}

/* ---------------------------------------------------------------- QPP interleaver (tc_interl_lte.c:69-94) */

int orc_interleaver(uint32_t K, uint16_t* fwd, uint16_t* rev)
{
  int idx = orc_cbindex(K);
  if (idx < 0 || qpp_rows[idx].K != K) {
    return -1;
  }
  uint64_t f1 = qpp_rows[idx].f1, f2 = qpp_rows[idx].f2;
  for (uint64_t i = 0; i < K; i++) {
    uint64_t p = (f1 * i + f2 * i * i) % K;
    fwd[i]     = (uint16_t)p;
    rev[p]     = (uint16_t)i;
  }
  return 0;
}

/* ---------------------------------------------------------------- CRC24A/B, MSB first, zero init (crc.c:30-160) */

#define ORC_CRC24A 0x1864CFB /* phy_common.h:72 */
#define ORC_CRC24B 0x1800063 /* phy_common.h:73 */

static uint32_t crc24_bits_msb(uint32_t poly, const uint8_t* bytes, int nbits)
{
  uint32_t reg = 0;
  for (int i = 0; i < nbits; i++) {
    uint32_t bit = (bytes[i >> 3] >> (7 - (i & 7))) & 1u;
    uint32_t top = ((reg >> 23) & 1u) ^ bit;
    reg          = (reg << 1) & 0xFFFFFFu;
    if (top) {
      reg ^= (poly & 0xFFFFFFu);
    }
  }
  return reg;
}

/* kind 0 = CRC24A, 1 = CRC24B; nbits multiple of 8 like srsran_crc_checksum_byte (crc.c:147) */
uint32_t orc_crc24(int kind, const uint8_t* bytes, int nbits)
{
  return crc24_bits_msb(kind ? ORC_CRC24B : ORC_CRC24A, bytes, nbits & ~7);
}

/* ---------------------------------------------------------------- turbo encoder (turbocoder.c:77-185) */

typedef struct {
  uint8_t r0, r1, r2;
} rsc_t;

/* one step of the 1 + D^2 + D^3 / 1 + D + D^3 recursive systematic constituent code */
static inline uint8_t rsc_step(rsc_t* s, uint8_t bit)
{
  uint8_t fb  = bit ^ s->r2 ^ s->r1;
  uint8_t par = s->r2 ^ s->r0 ^ fb;
  s->r2       = s->r1;
  s->r1       = s->r0;
  s->r0       = fb;
  return par;
}

/* bits: K values 0/1; out: 3K+12 values, d0 d1 d2 interleaved then 12 tail bits in the reference's order */
int orc_tcod_encode(const uint8_t* bits, uint8_t* out, uint32_t K)
{
  uint16_t* fwd = malloc(sizeof(uint16_t) * K);
  uint16_t* rev = malloc(sizeof(uint16_t) * K);
  if (orc_interleaver(K, fwd, rev)) {
    free(fwd);
    free(rev);
    return -1;
  }
  rsc_t a = {0, 0, 0}, b = {0, 0, 0};
  for (uint32_t i = 0; i < K; i++) {
    out[3 * i]     = bits[i];
    out[3 * i + 1] = rsc_step(&a, bits[i]);
    out[3 * i + 2] = rsc_step(&b, bits[fwd[i]]);
  }
  uint32_t k = 3 * K;
  for (int t = 0; t < 3; t++) { /* termination of encoder 1: input = feedback so the register drains */
    uint8_t x = a.r2 ^ a.r1;
    out[k++]  = x;
    out[k++]  = rsc_step(&a, x);
  }
  for (int t = 0; t < 3; t++) {
    uint8_t x = b.r2 ^ b.r1;
    out[k++]  = x;
    out[k++]  = rsc_step(&b, x);
  }
  free(fwd);
  free(rev);
  return 0;
}

/* ---------------------------------------------------------------- rate matching (rm_turbo.c:70-71,175-248,981-1068) */

static const uint8_t rm_colperm[32] = {0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30,
                                       1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31};

typedef struct {
  int D, R, Kpi, Nd, Ncb;
} rm_geom_t;

static rm_geom_t rm_geom(uint32_t K)
{
  rm_geom_t g;
  g.D   = (int)K + 4;
  g.R   = (g.D - 1) / 32 + 1;
  g.Kpi = 32 * g.R;
  g.Nd  = g.Kpi - g.D;
  g.Ncb = 3 * g.Kpi;
  return g;
}

static int rm_k0(const rm_geom_t* g, uint32_t rv)
{
  /* rm_turbo.c:185: R * (2 * ceil(Ncb / (8R)) * rv + 2) */
  int c = (g->Ncb + 8 * g->R - 1) / (8 * g->R);
  return g->R * (2 * c * (int)rv + 2);
}

/*
 * Circular-buffer position p -> index into the natural stream 3*bit + stream (bit in 0..K+3), or -1 for a dummy.
 * Stream 0 and 1: position inside the sub-block = col*R + row holds element row*32 + perm[col];
 * stream 2: element (perm[col] + 32*row + 1) mod Kpi  (rm_turbo.c:195-219).
 */
static int rm_pos_to_natural(const rm_geom_t* g, int p)
{
  int stream, q, e;
  if (p < g->Kpi) {
    stream = 0;
    q      = p;
  } else if (((p - g->Kpi) & 1) == 0) {
    stream = 1;
    q      = (p - g->Kpi) / 2;
  } else {
    stream = 2;
    q      = (p - g->Kpi - 1) / 2;
  }
  int col = q / g->R, row = q % g->R;
  if (stream < 2) {
    e = row * 32 + rm_colperm[col];
  } else {
    e = (rm_colperm[col] + 32 * row + 1) % g->Kpi;
  }
  if (e < g->Nd) {
    return -1;
  }
  return 3 * (e - g->Nd) + stream;
}

/* table[i] = natural index that receives the i-th transmitted value, i < 3K+12 (rm_turbo.c:175-248) */
int orc_rm_table(uint32_t cb_idx, uint32_t rv, uint16_t* table)
{
  if (cb_idx >= ORC_NOF_K || rv > 3) {
    return -2;
  }
  uint32_t  K = qpp_rows[cb_idx].K;
  rm_geom_t g = rm_geom(K);
  int       n = 3 * (int)K + 12, k0 = rm_k0(&g, rv), i = 0;
  for (int j = 0; i < n; j++) {
    int d = rm_pos_to_natural(&g, (k0 + j) % g.Ncb);
    if (d >= 0) {
      table[i++] = (uint16_t)d;
    }
  }
  return 0;
}

/* Transmit side, for test-vector synthesis only: E bits of redundancy version rv from 3K+12 coded bits */
int orc_rm_tx(const uint8_t* coded, uint32_t K, uint8_t* out, uint32_t E, uint32_t rv)
{
  int cb_idx = orc_cbindex(K);
  if (cb_idx < 0 || qpp_rows[cb_idx].K != K || rv > 3) {
    return -2;
  }
  uint32_t  n     = 3 * K + 12;
  uint16_t* table = malloc(sizeof(uint16_t) * n);
  orc_rm_table((uint32_t)cb_idx, rv, table);
  for (uint32_t i = 0; i < E; i++) {
    out[i] = coded[table[i % n]];
  }
  free(table);
  return 0;
}

/* soft[table[i mod n]] += in[i], int16 wrap (rm_turbo.c:435-437), natural layout */
int orc_rm_rx(const int16_t* in, int16_t* soft, uint32_t E, uint32_t cb_idx, uint32_t rv)
{
  if (cb_idx >= ORC_NOF_K || rv > 3) {
    return -2; /* SRSRAN_ERROR_INVALID_INPUTS, rm_turbo.c:442-444 */
  }
  uint32_t  n     = 3 * (uint32_t)qpp_rows[cb_idx].K + 12;
  uint16_t* table = malloc(sizeof(uint16_t) * n);
  orc_rm_table(cb_idx, rv, table);
  for (uint32_t i = 0; i < E; i++) {
    uint16_t d = table[i % n];
    soft[d]    = w16(soft[d] + in[i]);
  }
  free(table);
  return 0;
}

/* ---------------------------------------------------------------- max-log-MAP SISO (turbodecoder_gen.c:58-236) */

typedef struct {
  int16_t* beta; /* 8*(K+4) */
} siso_ws_t;

/*
 * L = SISO(x_in, apriori|NULL, par).  x_in and par have K+3 entries (3 tail steps), apriori K.
 * Backward sweep stores the un-normalised metrics, normalises every 4th step below K; forward sweep forms
 * L[k-1] = max over 1-branches - max over 0-branches, all in wrapping int16.
 */
static void siso_run(siso_ws_t* ws, const int16_t* x_in, const int16_t* ap, const int16_t* par, int16_t* L, int K)
{
  int16_t* beta = ws->beta;
  int16_t  B[8], A[8], nb[8];

  B[0] = 0;
  for (int i = 1; i < 8; i++) B[i] = -ORC_INF;
  /* (the reference also parks this vector at beta[8*(K+3)], never read by the forward sweep) */

  for (int k = K + 2; k >= 0; k--) {
    int16_t x  = x_in[k];
    if (ap && k < K) x = w16(x + ap[k]);
    int16_t y  = par[k];
    int16_t xy = w16(x + y);
    /* candidate through the 0-input ("new") and 1-input ("m_b") branches, turbodecoder_gen.c:79-95 */
    int16_t c0[8] = {B[0], w16(B[0] + xy), w16(B[1] + x), w16(B[1] + y), w16(B[2] + y), w16(B[2] + x), w16(B[3] + xy), B[3]};
    int16_t c1[8] = {w16(B[4] + xy), B[4], w16(B[5] + y), w16(B[5] + x), w16(B[6] + x), w16(B[6] + y), B[7], w16(B[7] + xy)};
    for (int i = 0; i < 8; i++) {
      nb[i]           = c1[i] > c0[i] ? c1[i] : c0[i];
      beta[8 * k + i] = nb[i];
    }
    if ((k % 4) == 0 && k < K) {
      for (int i = 1; i < 8; i++) nb[i] = w16(nb[i] - nb[0]);
      nb[0] = 0;
    }
    memcpy(B, nb, sizeof(B));
  }

  A[0] = 0;
  for (int i = 1; i < 8; i++) A[i] = -ORC_INF;
  for (int k = 1; k <= K; k++) {
    int16_t x = x_in[k - 1];
    if (ap) x = w16(x + ap[k - 1]);
    int16_t y  = par[k - 1];
    int16_t xy = w16(x + y);
    /* branches carrying input bit 0 (m) and 1 (n), indexed by destination state, turbodecoder_gen.c:139-155 */
    int16_t m[8] = {A[0], w16(A[3] + y), w16(A[4] + y), A[7], A[1], w16(A[2] + y), w16(A[5] + y), A[6]};
    int16_t n[8] = {w16(A[1] + xy), w16(A[2] + x), w16(A[5] + x), w16(A[6] + xy),
                    w16(A[0] + xy), w16(A[3] + x), w16(A[4] + x), w16(A[7] + xy)};
    const int16_t* bk = &beta[8 * k];
    int16_t        m0 = w16(m[0] + bk[0]), m1 = w16(n[0] + bk[0]);
    for (int i = 1; i < 8; i++) {
      int16_t t0 = w16(m[i] + bk[i]), t1 = w16(n[i] + bk[i]);
      if (t0 > m0) m0 = t0;
      if (t1 > m1) m1 = t1;
    }
    for (int i = 0; i < 8; i++) A[i] = m[i] > n[i] ? m[i] : n[i];
    if ((k % 4) == 0) {
      for (int i = 1; i < 8; i++) A[i] = w16(A[i] - A[0]);
      A[0] = 0;
    }
    L[k - 1] = w16(m1 - m0);
  }
}

typedef struct {
  int       K;
  int       n_pass;
  uint16_t *fwd, *rev;
  int16_t * syst, *par0, *par1, *app1, *app2, *ext1, *ext2;
  siso_ws_t ws;
} tdec_t;

static int tdec_alloc(tdec_t* d)
{
  memset(d, 0, sizeof(*d));
  size_t n = ORC_MAX_K + 16;
  d->fwd   = malloc(sizeof(uint16_t) * n);
  d->rev   = malloc(sizeof(uint16_t) * n);
  d->syst  = calloc(n, sizeof(int16_t));
  d->par0  = calloc(n, sizeof(int16_t));
  d->par1  = calloc(n, sizeof(int16_t));
  d->app1  = calloc(n, sizeof(int16_t));
  d->app2  = calloc(n, sizeof(int16_t));
  d->ext1  = calloc(n, sizeof(int16_t));
  d->ext2  = calloc(n, sizeof(int16_t));
  d->ws.beta = calloc(8 * n, sizeof(int16_t));
  return 0;
}

static void tdec_release(tdec_t* d)
{
  free(d->fwd);
  free(d->rev);
  free(d->syst);
  free(d->par0);
  free(d->par1);
  free(d->app1);
  free(d->app2);
  free(d->ext1);
  free(d->ext2);
  free(d->ws.beta);
}

static int tdec_new_cb(tdec_t* d, int K)
{
  if (orc_interleaver((uint32_t)K, d->fwd, d->rev)) {
    return -1;
  }
  d->K      = K;
  d->n_pass = 0;
  return 0;
}

/* One SISO pass of the schedule in turbodecoder_iter.h:72-144, natural-layout input (turbodecoder_gen.c:238-258) */
static void tdec_pass(tdec_t* d, const int16_t* in)
{
  int K = d->K, n = d->n_pass;
  if (n == 0) {
    for (int i = 0; i < K; i++) {
      d->syst[i] = in[3 * i];
      d->par0[i] = in[3 * i + 1];
      d->par1[i] = in[3 * i + 2];
    }
    for (int t = 0; t < 3; t++) {
      d->syst[K + t] = in[3 * K + 2 * t];
      d->par0[K + t] = in[3 * K + 2 * t + 1];
      d->app2[K + t] = in[3 * K + 6 + 2 * t];
      d->par1[K + t] = in[3 * K + 6 + 2 * t + 1];
    }
  }
  if ((n & 1) == 0) {
    if (n > 0) {
      for (int i = 0; i < K; i++) d->app1[i] = w16(d->app1[i] - d->ext1[i]);
    }
    siso_run(&d->ws, d->syst, n > 0 ? d->app1 : NULL, d->par0, d->ext1, K);
  } else {
    if (n > 1) {
      for (int i = 0; i < K; i++) d->ext1[i] = w16(d->ext1[i] - d->app1[i]);
    }
    for (int i = 0; i < K; i++) d->app2[d->rev[i]] = d->ext1[i];
    siso_run(&d->ws, d->app2, NULL, d->par1, d->ext2, K);
    for (int i = 0; i < K; i++) d->app1[d->fwd[i]] = d->ext2[i];
  }
  d->n_pass++;
}

/* hard decision after the pass just run: MSB-first, LLR>0 -> 1 (turbodecoder.c:370-378, turbodecoder_gen.c:260-277) */
static void tdec_decide(const tdec_t* d, uint8_t* out)
{
  const int16_t* v = (d->n_pass & 1) ? d->ext1 : d->app1;
  for (int i = 0; i < d->K / 8; i++) {
    uint8_t b = 0;
    for (int j = 0; j < 8; j++) b = (uint8_t)((b << 1) | (v[8 * i + j] > 0));
    out[i] = b;
  }
}

/* npass passes on one block, decisions after every pass: out[p*K/8 ..] */
int orc_tdec_passes(const int16_t* llr, uint32_t K, uint32_t npass, uint8_t* out)
{
  tdec_t d;
  tdec_alloc(&d);
  int r = tdec_new_cb(&d, (int)K);
  for (uint32_t p = 0; !r && p < npass; p++) {
    tdec_pass(&d, llr);
    tdec_decide(&d, &out[(size_t)p * (K / 8)]);
  }
  tdec_release(&d);
  return r;
}

/* Debug/white-box view: the a-posteriori vector the decision is taken on after each pass (K int16 per pass) */
int orc_tdec_passes_llr(const int16_t* llr, uint32_t K, uint32_t npass, int16_t* out)
{
  tdec_t d;
  tdec_alloc(&d);
  int r = tdec_new_cb(&d, (int)K);
  for (uint32_t p = 0; !r && p < npass; p++) {
    tdec_pass(&d, llr);
    memcpy(&out[(size_t)p * K], (d.n_pass & 1) ? d.ext1 : d.app1, sizeof(int16_t) * K);
  }
  tdec_release(&d);
  return r;
}

typedef struct {
  const int16_t* llr;
  uint32_t       first, last, K, max_pass;
  int            crc_kind;
  uint32_t       crc_len;
  int            early_stop;
  uint8_t *      out, *crc_ok, *npass;
} orc_job_t;

/* the per-code-block loop of decode_tb_cb (sch.c:420-454) */
static void* orc_job(void* arg)
{
  orc_job_t* j = arg;
  tdec_t     d;
  tdec_alloc(&d);
  uint32_t K    = j->K;
  size_t   nllr = 3 * (size_t)K + 12;
  for (uint32_t cb = j->first; cb < j->last; cb++) {
    uint8_t* data = &j->out[(size_t)cb * (K / 8)];
    tdec_new_cb(&d, (int)K);
    uint32_t noi = 0;
    int      ok = 0, stop = 0;
    do {
      tdec_pass(&d, &j->llr[cb * nllr]);
      tdec_decide(&d, data);
      noi++;
      if (j->crc_kind != 2 && !ok) {
        uint32_t len = j->crc_kind == 1 ? j->crc_len : K;
        if (crc24_bits_msb(j->crc_kind == 1 ? ORC_CRC24A : ORC_CRC24B, data, (int)len) == 0) {
          ok           = 1;
          j->npass[cb] = (uint8_t)noi;
          stop         = j->early_stop;
        }
      }
    } while (noi < j->max_pass && !stop);
    j->crc_ok[cb] = (uint8_t)ok;
    if (!ok) j->npass[cb] = (uint8_t)noi;
  }
  tdec_release(&d);
  return NULL;
}

/* same contract as ref_decode_batch in ref_harness.c; crc_kind 0: CRC24B over K bits, 1: CRC24A over crc_len, 2: none */
int orc_decode_batch(const int16_t* llr,
                     uint32_t       ncb,
                     uint32_t       K,
                     uint32_t       max_pass,
                     int            crc_kind,
                     uint32_t       crc_len,
                     int            early_stop,
                     uint8_t*       out,
                     uint8_t*       crc_ok,
                     uint8_t*       npass,
                     int            nthreads,
                     double*        seconds)
{
  int cbi = orc_cbindex(K);
  if (cbi < 0 || qpp_rows[cbi].K != K) {
    return -1;
  }
  if (nthreads < 1) nthreads = 1;
  if ((uint32_t)nthreads > ncb) nthreads = ncb ? (int)ncb : 1;
  orc_job_t*      jobs = calloc((size_t)nthreads, sizeof(*jobs));
  pthread_t*      th   = calloc((size_t)nthreads, sizeof(*th));
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = (orc_job_t){llr,
                          (uint32_t)(((uint64_t)ncb * t) / nthreads),
                          (uint32_t)(((uint64_t)ncb * (t + 1)) / nthreads),
                          K,
                          max_pass,
                          crc_kind,
                          crc_len,
                          early_stop,
                          out,
                          crc_ok,
                          npass};
    pthread_create(&th[t], NULL, orc_job, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  free(jobs);
  free(th);
  return 0;
}

/* ---------------------------------------------------------------- OFDM receive (ofdm.c:38-212,334-362,387-466) */

typedef float _Complex cf_t;
typedef double _Complex cd_t;

static int ofdm_cp_len(int c, int N)
{
  return (int)ceilf(((float)c * (float)N) / 2048.0f); /* phy_common.h:125 */
}

/* default (non "standard rate") symbol sizes, phy_common.c:361-385; std!=0 selects phy_common.c:342-359 */
int orc_symbol_sz(uint32_t nof_prb, int std)
{
  if (nof_prb == 0) return -1;
  if (std) {
    if (nof_prb <= 6) return 128;
    if (nof_prb <= 15) return 256;
    if (nof_prb <= 25) return 512;
    if (nof_prb <= 50) return 1024;
    if (nof_prb <= 75) return 1536;
    if (nof_prb <= 110) return 2048;
    return -1;
  }
  if (nof_prb <= 6) return 128;
  if (nof_prb <= 15) return 256;
  if (nof_prb <= 25) return 384;
  if (nof_prb <= 50) return 768;
  if (nof_prb <= 75) return 1024;
  if (nof_prb <= 110) return 1536;
  return -1;
}

static int pick_radix(int n)
{
  if (n % 4 == 0) return 4;
  if (n % 2 == 0) return 2;
  for (int f = 3; f * f <= n; f += 2)
    if (n % f == 0) return f;
  return n;
}

static void dft_f64(const cd_t* x, int xs, cd_t* y, cd_t* tmp, int n, const cd_t* tw, int tstep, int ntop)
{
  if (n == 1) {
    y[0] = x[0];
    return;
  }
  int r = pick_radix(n), m = n / r;
  for (int q = 0; q < r; q++) dft_f64(x + (size_t)q * xs, xs * r, tmp + (size_t)q * m, y, m, tw, tstep * r, ntop);
  for (int k = 0; k < m; k++) {
    for (int p = 0; p < r; p++) {
      int  kk  = k + m * p;
      cd_t acc = 0;
      for (int q = 0; q < r; q++) acc += tmp[(size_t)q * m + k] * tw[((long)q * kk * tstep) % ntop];
      y[kk] = acc;
    }
  }
}

/*
 * nsf subframes of srsran_ofdm_rx_sf semantics (normal subframe type):
 *   in  : nsf * 15N samples (NOT modified; the reference multiplies the half-subcarrier shift into its bound
 *         input buffer in place, ofdm.c:455-457 -- the oracle leaves the caller's array alone)
 *   out : nsf * nsym*2 * 12*nof_prb, symbol major
 * symbol_sz==0 derives N from nof_prb with the non-standard table (the reference default).
 */
int orc_ofdm_rx(uint32_t    nof_prb,
                int         cp_ext,
                uint32_t    symbol_sz,
                float       freq_shift,
                float       rx_window_offset,
                int         normalize,
                int         keep_dc,
                const cf_t* in,
                cf_t*       out,
                uint32_t    nsf)
{
  int N = symbol_sz ? (int)symbol_sz : orc_symbol_sz(nof_prb, 0);
  if (N <= 0) return -1;
  int nsym  = cp_ext ? 6 : 7;
  int cp1   = cp_ext ? ofdm_cp_len(512, N) : ofdm_cp_len(160, N);
  int cp2   = cp_ext ? ofdm_cp_len(512, N) : ofdm_cp_len(144, N);
  int nre   = 12 * (int)nof_prb;
  int sf_sz = 15 * N, slot_sz = sf_sz / 2;
  int shift = isnormal(freq_shift);
  int noff  = 0;
  if (isnormal(rx_window_offset)) {
    float w = rx_window_offset < 0 ? 0 : (rx_window_offset > 100 ? 100 : rx_window_offset);
    noff    = (int)roundf((float)cp2 * w); /* ofdm.c:130-133 */
  }
  int dc = (!keep_dc && !shift) ? 1 : 0; /* ofdm.c:209 */

  cd_t* tw  = malloc(sizeof(cd_t) * (size_t)N);
  cd_t* x   = malloc(sizeof(cd_t) * (size_t)N);
  cd_t* X   = malloc(sizeof(cd_t) * (size_t)N);
  cd_t* tmp = malloc(sizeof(cd_t) * (size_t)N);
  for (int k = 0; k < N; k++) tw[k] = cexp(-I * 2.0 * M_PI * (double)k / (double)N);
  double norm = normalize ? 1.0 / sqrt((double)N) : 1.0;

  for (uint32_t s = 0; s < nsf; s++) {
    const cf_t* sf = in + (size_t)s * sf_sz;
    for (int slot = 0; slot < 2; slot++) {
      for (int l = 0; l < nsym; l++) {
        /* first sample of the FFT window: CP stripped, then slid back noff samples into the CP (ofdm.c:160) */
        int start = slot * slot_sz + cp1 + l * (N + cp2) - noff;
        /* position of this symbol's CP start inside the subframe, for the shift phase reference (ofdm.c:347-355) */
        int sym_begin = slot * slot_sz + (l == 0 ? 0 : cp1 + N + (l - 1) * (N + cp2));
        int cplen     = (l == 0) ? cp1 : cp2;
        for (int t = 0; t < N; t++) {
          int  idx = start + t;
          cd_t v   = 0;
          if (idx >= 0 && idx < sf_sz) v = (cd_t)sf[idx];
          if (shift) {
            /* shift_buffer is laid out symbol by symbol; a window that slid into the previous symbol's tail
             * picks up that symbol's phase ramp, exactly like the in-place multiply does */
            int    rel;
            double ph;
            if (idx >= sym_begin) {
              rel = idx - sym_begin;
              ph  = ((double)rel - (double)cplen);
            } else {
              /* sample belongs to the previous symbol of the subframe */
              int pl      = l - 1, pslot = slot;
              if (pl < 0) {
                pl    = nsym - 1;
                pslot = slot - 1;
              }
              if (pslot < 0) {
                ph = 0;
                v  = 0;
              } else {
                int pb  = pslot * slot_sz + (pl == 0 ? 0 : cp1 + N + (pl - 1) * (N + cp2));
                int pcp = (pl == 0) ? cp1 : cp2;
                ph      = ((double)(idx - pb) - (double)pcp);
              }
            }
            v *= cexp(I * 2.0 * M_PI * ph * (double)freq_shift / (double)N);
          }
          x[t] = v;
        }
        dft_f64(x, 1, X, tmp, N, tw, 1, N);
        cf_t* o = out + ((size_t)s * 2 * nsym + (size_t)slot * nsym + l) * nre;
        for (int r = 0; r < nre; r++) {
          int  bin = r < nre / 2 ? N - nre / 2 + r : dc + (r - nre / 2); /* ofdm.c:410-411 */
          cd_t v   = X[bin];
          if (noff) v *= cexp(I * 2.0 * M_PI * (double)noff * (double)bin / (double)N); /* ofdm.c:134-136,405-407 */
          o[r] = (cf_t)(v * norm);
        }
      }
    }
  }
  free(tw);
  free(x);
  free(X);
  free(tmp);
  return 0;
}

/* ---------------------------------------------------------------- int16 soft demapper (demod_soft.c) */

static inline int16_t sat16(int v)
{
  return (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v));
}

static inline int16_t abs16_wrap(int16_t v)
{
  return v < 0 ? w16(-(int)v) : v; /* _mm_abs_epi16: |-32768| stays -32768 */
}

/*
 * mod: 1 QPSK, 2 16QAM, 3 64QAM (srsran_mod_t numbering, modem_table.h).  Restates the x86 build of the
 * reference (LV_HAVE_SSE): groups of 4 symbols are scaled by -S, rounded to nearest-even and saturated to int16
 * (cvtps_epi32 + packs_epi32, demod_soft.c:250-282,569-627); the <4 leftover symbols use the scalar tail that
 * truncates S*x toward zero and negates afterwards (demod_soft.c:283-298,629-642).
 */
int orc_demod_s(int mod, const cf_t* sym, int16_t* llr, int n)
{
  const float* f = (const float*)sym;
  if (mod == 3) {
    const int16_t t1 = (int16_t)(4 * 700 / sqrtf(42)), t2 = (int16_t)(2 * 700 / sqrtf(42));
    int           n4 = 4 * (n / 4);
    for (int i = 0; i < n; i++) {
      int16_t y[2];
      for (int c = 0; c < 2; c++) {
        if (i < n4) {
          y[c] = sat16((int)lrintf(f[2 * i + c] * -700.0f));
        } else {
          y[c] = w16(-(int)(int16_t)(700 * f[2 * i + c]));
        }
      }
      int16_t a0 = w16(abs16_wrap(i < n4 ? y[0] : w16(-y[0])) - t1);
      int16_t a1 = w16(abs16_wrap(i < n4 ? y[1] : w16(-y[1])) - t1);
      llr[6 * i + 0] = y[0];
      llr[6 * i + 1] = y[1];
      llr[6 * i + 2] = a0;
      llr[6 * i + 3] = a1;
      llr[6 * i + 4] = w16(abs16_wrap(a0) - t2);
      llr[6 * i + 5] = w16(abs16_wrap(a1) - t2);
    }
    return 0;
  }
  if (mod == 1) {
    /* demod_qpsk_lte_s (demod_soft.c:115-118) -> srsran_vec_convert_fi (vector_simd.c:436-472, AVX2 build): blocks of 16
     * floats are scaled, truncated toward zero (cvttps, simd.h:1893) and saturated by the pack; the remaining floats take
     * the plain C cast */
    const float scale = (float)(-100 * M_SQRT2);
    int         nf = 2 * n, body = 16 * (nf / 16);
    for (int i = 0; i < nf; i++) {
      float v = f[i] * scale;
      llr[i]  = i < body ? sat16((int)v) : w16((int)v);
    }
    return 0;
  }
  if (mod == 2) {
    const int16_t t  = (int16_t)(2 * 400 / sqrtf(10));
    int           n4 = 4 * (n / 4);
    for (int i = 0; i < n; i++) {
      int16_t y[2];
      for (int c = 0; c < 2; c++) {
        if (i < n4) {
          y[c] = sat16((int)lrintf(f[2 * i + c] * -400.0f));
        } else {
          y[c] = w16(-(int)(int16_t)(400 * f[2 * i + c]));
        }
      }
      llr[4 * i + 0] = y[0];
      llr[4 * i + 1] = y[1];
      if (i < n4) {
        llr[4 * i + 2] = w16(abs16_wrap(y[0]) - t);
        llr[4 * i + 3] = w16(abs16_wrap(y[1]) - t);
      } else {
        /* scalar tail: abs(yre) - 2*S/sqrtf(10) evaluated in float then truncated (demod_soft.c:295-296) */
        llr[4 * i + 2] = (int16_t)((float)abs((int)(int16_t)(400 * f[2 * i])) - 2 * 400 / sqrtf(10));
        llr[4 * i + 3] = (int16_t)((float)abs((int)(int16_t)(400 * f[2 * i + 1])) - 2 * 400 / sqrtf(10));
      }
    }
    return 0;
  }
  return -1;
}

/* ---------------------------------------------------------------- transport block decode (sch.c:370-572) */

#define ORC_SOFTBUFFER_SIZE 18600 /* softbuffer.h:56 */

/*
 * decode_tb + decode_tb_cb for one transport block.
 *   e_bits : G = nof_e_bits soft bits;  soft : C buffers of ORC_SOFTBUFFER_SIZE int16 (combined in place);
 *   cb_crc : C flags, in/out (blocks already decoded are skipped and their bytes are expected in `data`);
 *   data   : tbs/8 + 3 bytes (+ K/8 slack);  iter_sum : sum of passes run (avg_iterations * C).
 * Returns 0 when every code block CRC and the transport block CRC24A match, -1 on CRC failure, -2 on invalid input.
 */
int orc_decode_tb(const int16_t* e_bits,
                  uint32_t       nof_e_bits,
                  uint32_t       tbs,
                  uint32_t       Qm,
                  uint32_t       rv,
                  uint32_t       max_iter,
                  int16_t*       soft,
                  uint8_t*       cb_crc,
                  uint8_t*       data,
                  uint32_t*      iter_sum)
{
  uint32_t s[9];
  if (!e_bits || !soft || !data || Qm == 0 || orc_cbsegm(tbs, s)) return -2;
  uint32_t F = s[0], C = s[1], K1 = s[2], K2 = s[3], K1i = s[4], K2i = s[5], C1 = s[6];
  if (iter_sum) *iter_sum = 0;
  if (tbs == 0 || C == 0) return 0; /* sch.c:517-519 */
  if (F) return -2;                 /* sch.c:521-524 */
  if (C > 32) return -2;
  data[tbs / 8] = data[tbs / 8 + 1] = data[tbs / 8 + 2] = 0; /* sch.c:537-539 */

  tdec_t d;
  tdec_alloc(&d);
  uint32_t Gp = nof_e_bits / Qm, gamma = Gp % C, n_e = Qm * (Gp / C);
  for (uint32_t cb = 0; cb < C; cb++) {
    if (cb_crc[cb]) continue; /* sch.c:390 (the saved bytes are the caller's business here) */
    uint32_t K    = cb < C1 ? K1 : K2;
    uint32_t Ki   = cb < C1 ? K1i : K2i;
    uint32_t rlen = C == 1 ? K : K - 24;
    uint32_t rp = cb * n_e, n_e2 = n_e;
    if (cb > C - gamma) { /* sch.c:403, the reference's own comparison */
      n_e2 = n_e + Qm;
      rp   = (C - gamma) * n_e + (cb - (C - gamma)) * n_e2;
    }
    int16_t* sb = soft + (size_t)cb * ORC_SOFTBUFFER_SIZE;
    orc_rm_rx(&e_bits[rp], sb, n_e2, Ki, rv);
    tdec_new_cb(&d, (int)K);
    uint8_t* out = &data[cb * rlen / 8];
    uint32_t noi = 0;
    int      ok  = 0;
    do {
      tdec_pass(&d, sb);
      tdec_decide(&d, out);
      noi++;
      if (iter_sum) (*iter_sum)++;
      uint32_t len = C > 1 ? K : tbs + 24;
      if (crc24_bits_msb(C > 1 ? ORC_CRC24B : ORC_CRC24A, out, (int)len) == 0) ok = 1;
    } while (noi < max_iter && !ok);
    cb_crc[cb] = (uint8_t)ok;
  }
  tdec_release(&d);
  for (uint32_t cb = 0; cb < C; cb++)
    if (!cb_crc[cb]) return -1;
  uint32_t par_rx = crc24_bits_msb(ORC_CRC24A, data, (int)tbs);
  uint32_t par_tx = ((uint32_t)data[tbs / 8] << 16) | ((uint32_t)data[tbs / 8 + 1] << 8) | data[tbs / 8 + 2];
  return (par_rx == par_tx && par_rx) ? 0 : -1; /* sch.c:553 */
}

/* ================================================================================================================
 * PUSCH receive chain between the OFDM demodulator and the rate de-matcher
 * ================================================================================================================ */

/* ---- pseudo-random (Gold) sequence, TS 36.211 7.2 as used by sequence.c:33-120,494-548 --------------------------------
 * c(n) = x1(n+1600) ^ x2(n+1600), x1(n+31) = x1(n+3)^x1(n), x2(n+31) = x2(n+3)^x2(n+2)^x2(n+1)^x2(n),
 * x1(0)=1, x2 = c_init.  Plain bit-serial form. */
static void orc_gold(uint32_t c_init, uint8_t* c, uint32_t len)
{
  uint32_t x1 = 1, x2 = c_init;
  for (uint32_t n = 0; n < 1600 + len; n++) {
    if (n >= 1600) c[n - 1600] = (uint8_t)((x1 ^ x2) & 1u);
    uint32_t f1 = ((x1 >> 3) ^ x1) & 1u;
    uint32_t f2 = ((x2 >> 3) ^ (x2 >> 2) ^ (x2 >> 1) ^ x2) & 1u;
    x1          = (x1 >> 1) | (f1 << 30);
    x2          = (x2 >> 1) | (f2 << 30);
  }
}

int orc_gold_bits(uint32_t c_init, uint8_t* c, uint32_t len)
{
  orc_gold(c_init, c, len);
  return 0;
}

/* sequences.c:120-147 + sequence.c:494-548: out[i] = c[i] ? -in[i] : in[i], int16 wrap (-(-32768) stays -32768) */
void orc_pusch_seq_apply_s(const int16_t* in, int16_t* out, uint32_t rnti, uint32_t nslot, uint32_t cell_id, uint32_t len)
{
  uint8_t* c = malloc(len ? len : 1);
  orc_gold(((rnti & 0xFFFFu) << 14) + ((nslot / 2) << 9) + cell_id, c, len);
  for (uint32_t i = 0; i < len; i++) out[i] = c[i] ? w16(-(int)in[i]) : in[i];
  free(c);
}

/* sch.c:660-681,993-1020 without RI bits: q is symbol(column)-major as the demodulator leaves it, g row-major:
 * g[(j*cols + i)*Qm + k] = q[(i*rows + j)*Qm + k], rows = H'/N_pusch_symbs, cols = N_pusch_symbs */
int orc_ulsch_deinterleave(const int16_t* q_bits, int16_t* g_bits, uint32_t Qm, uint32_t H_prime_total, uint32_t N_pusch_symbs)
{
  const uint32_t rows = H_prime_total / N_pusch_symbs, cols = N_pusch_symbs;
  uint32_t       idx = 0;
  for (uint32_t j = 0; j < rows; j++)
    for (uint32_t i = 0; i < cols; i++)
      for (uint32_t k = 0; k < Qm; k++) g_bits[idx++] = q_bits[j * Qm + i * rows * Qm + k];
  return 0;
}

/* dft_precoding.c:39-64,114-126 over dft_fftw.c (norm => 1/sqrt(N), dft_fftw.c:343-350): nof_symbols transforms of
 * 12*nof_prb points, forward (tx) e^{-j} or backward (rx) e^{+j}, scaled by 1/sqrt(N).  float64 arithmetic. */
int orc_dft_precoding(const float complex* in, float complex* out, uint32_t nof_prb, uint32_t nof_symbols, int is_tx)
{
  const uint32_t  N = 12 * nof_prb;
  double complex* w = malloc(sizeof(double complex) * N);
  const double    sgn = is_tx ? -1.0 : 1.0, sc = 1.0 / sqrt((double)N);
  for (uint32_t m = 0; m < N; m++) w[m] = cexp(I * sgn * 2.0 * M_PI * (double)m / (double)N);
  for (uint32_t s = 0; s < nof_symbols; s++) {
    for (uint32_t k = 0; k < N; k++) {
      double complex acc = 0;
      for (uint32_t n = 0; n < N; n++) acc += (double complex)in[s * N + n] * w[(uint32_t)(((uint64_t)k * n) % N)];
      out[s * N + k] = (float complex)(acc * sc);
    }
  }
  free(w);
  return 0;
}

/* precoding.c:182-305,357: x = y conj(h) / ((|h|^2 + noise) * scaling), one receive antenna.  The reference's SIMD body
 * only adds the noise term when noise_estimate > 0 (precoding.c:235) */
int orc_predecoding_single(const float complex* y, const float complex* h, float complex* x, int n, float scaling, float noise)
{
  for (int i = 0; i < n; i++) {
    float hr = crealf(h[i]), hi = cimagf(h[i]), yr = crealf(y[i]), yi = cimagf(y[i]);
    float hh = hr * hr + hi * hi;
    if (noise > 0) hh += noise;
    float re = yr * hr + yi * hi, im = yi * hr - yr * hi;
    x[i]     = ((re / hh) * (1.0f / scaling)) + I * ((im / hh) * (1.0f / scaling));
  }
  return n;
}

/* Link parameters shared with oracle/ref_harness.c (all uint32):
 *  0 cell_id  1 cell nof_prb  2 cp_ext  3 dmrs cyclic_shift  4 delta_ss  5 group_hopping  6 sequence_hopping
 *  7 rnti  8 tti  9 L_prb  10 n_prb  11 mod  12 tbs  13 rv  14 n_dmrs  15 max_nof_iterations */
enum { OP_CELL_ID, OP_NOF_PRB, OP_CP_EXT, OP_CSHIFT, OP_DELTA_SS, OP_GH, OP_SH, OP_RNTI, OP_TTI, OP_L_PRB, OP_N_PRB, OP_MOD,
       OP_TBS, OP_RV, OP_N_DMRS, OP_MAX_ITER };

static uint32_t orc_prime_lower_than(uint32_t n) /* primes.c: largest prime < n */
{
  for (uint32_t p = n - 1; p >= 2; p--) {
    int ok = 1;
    for (uint32_t d = 2; d * d <= p; d++)
      if (p % d == 0) {
        ok = 0;
        break;
      }
    if (ok) return p;
  }
  return 0;
}

/* refsignal_ul.c:95-181,227-249,337-357 + zc_sequence.c:205-235,273-300 + phy_common.c:471-489.
 * PUSCH DMRS of one subframe, r[2][12*L_prb] (slot-major).  M_sc >= 36 is the Zadoff-Chu form; the 1- and 2-PRB base
 * sequences are the phi(n) pi/4 table look-ups of TS 36.211 5.5.1.2 (zc_sequence.c:175-183,232-247). */
#include "phi_table.inc"
int orc_dmrs_pusch_gen(const uint32_t* p, float complex* r)
{
  static const uint32_t n_dmrs_1[8] = {0, 2, 3, 4, 6, 8, 9, 10}, n_dmrs_2[8] = {0, 6, 3, 4, 2, 8, 10, 9}; /* 36.211 5.5.2.1.1 */
  const uint32_t cell_id = p[OP_CELL_ID], L = p[OP_L_PRB], sf_idx = p[OP_TTI] % 10, dss = p[OP_DELTA_SS];
  const uint32_t nsymb = p[OP_CP_EXT] ? 6 : 7, M = 12 * L;
  if (L < 1 || p[OP_CSHIFT] > 7 || p[OP_N_DMRS] > 7 || dss > 29) return -1;
  uint8_t        c[8 * 7 * 20];
  const uint32_t c_init = ((cell_id / 30) << 5) + (((cell_id % 30) + dss) % 30);
  orc_gold(c_init, c, 8 * nsymb * 20);
  uint8_t cg[160];
  orc_gold(cell_id / 30, cg, 160);
  const uint32_t Nzc = orc_prime_lower_than(M);
  for (uint32_t ns = 2 * sf_idx; ns < 2 * sf_idx + 2; ns++) {
    uint32_t n_prs = 0, f_gh = 0;
    for (int i = 0; i < 8; i++) n_prs += (uint32_t)c[8 * nsymb * ns + i] << i;
    if (p[OP_GH])
      for (int i = 0; i < 8; i++) f_gh += (uint32_t)cg[8 * ns + i] << i;
    const uint32_t n_cs  = (n_dmrs_1[p[OP_CSHIFT]] + n_dmrs_2[p[OP_N_DMRS]] + n_prs) % 12;
    const float    alpha = (float)(2 * M_PI * (n_cs) / 12);
    const uint32_t u     = (f_gh + (cell_id % 30) + dss) % 30;
    uint32_t       v     = 0;
    if (L >= 6 && p[OP_SH]) v = c[ns]; /* same generator, first 20 bits (refsignal_ul.c:121-126) */
    /* zc_sequence.c:205-218 */
    const float n_sz  = (float)Nzc;
    float       q_hat = n_sz * (u + 1) / 31, qf;
    if ((((uint32_t)(2 * q_hat)) % 2) == 0) qf = q_hat + 0.5 + v;
    else qf = q_hat + 0.5 - v;
    const float q = (float)(uint32_t)qf;
    for (uint32_t i = 0; i < M; i++) {
      const float m = (float)(i % (Nzc ? Nzc : 1));
      float       arg;
      if (M == 12) arg = (float)(2 * (orc_phi12[u][i] - '0') - 3) * (float)M_PI_4; /* srsran_vec_sc_prod_fcc: float times float */
      else if (M == 24) arg = (float)(2 * (orc_phi24[u][i] - '0') - 3) * (float)M_PI_4;
      else arg = (float)(-M_PI * q * m * (m + 1) / n_sz); /* double expression rounded into a float (cf_t) */
      /* the reference is built with -mfma and GCC contracts arg + alpha*i into one fused multiply-add (oracle/Makefile
       * uses the reference's own ISA flags); restated explicitly so that this file does not depend on its own flags */
      r[(ns % 2) * M + i] = cexpf(I * fmaf(alpha, (float)i, arg));
    }
  }
  return 0;
}

/* srsran_conv_same_cf, extrapolating variant (convolution.c:181-218), real filter of odd length M <= 7 */
static void orc_conv_same(const float complex* in, const float* f, float complex* out, uint32_t N, uint32_t M)
{
  float complex first[16], last[16];
  for (uint32_t i = 0; i < M + M / 2; i++) {
    if (i < M / 2) first[i] = (2 + M / 2 - i) * in[1] - (1 + M / 2 - i) * in[0];
    else first[i] = in[i - M / 2];
  }
  for (uint32_t i = 0; i < M + M / 2; i++) {
    if (i >= M - 1) last[i] = (2 + i - M / 2) * in[N - 1] - (1 + i - M / 2) * in[N - 2];
    else last[i] = in[N - M + i + 1];
  }
  uint32_t i = 0, j = 0;
  for (; i < M / 2; i++) {
    float complex a = 0;
    for (uint32_t t = 0; t < M; t++) a += first[i + t] * f[t];
    out[i] = a;
  }
  for (; i < N - M / 2; i++) {
    float complex a = 0;
    for (uint32_t t = 0; t < M; t++) a += in[i - M / 2 + t] * f[t];
    out[i] = a;
  }
  for (; i < N; i++, j++) {
    float complex a = 0;
    for (uint32_t t = 0; t < M; t++) a += last[j + t] * f[t];
    out[i] = a;
  }
}

/* chest_ul.c:225-357,370-400 with the object defaults of chest_ul.c:82-83 (3-tap filter, w = 0.3333), no linear
 * interpolation (the estimate of a slot's DMRS symbol is copied to its other symbols), no TA measurement.
 * grid/ce: 2*nsymb symbols of 12*nof_prb; dmrs: the known sequence [2][12*L_prb]; meas = {noise_estimate, snr, cfo_hz, 0} */
int orc_chest_ul_pusch(const uint32_t* p, const float complex* grid, const float complex* dmrs, float complex* ce, float* meas)
{
  const uint32_t nsymb = p[OP_CP_EXT] ? 6 : 7, R = 12 * p[OP_NOF_PRB], M = 12 * p[OP_L_PRB], off = 12 * p[OP_N_PRB];
  const float    w = 0.3333f;
  const float    f[3] = {w, 1 - 2 * w, w};
  float complex* ls   = malloc(sizeof(float complex) * 2 * M);
  float          noise = 0, rxpow = 0;
  for (uint32_t s = 0; s < 2; s++) {
    const uint32_t l = (s + 1) * nsymb - 4;
    for (uint32_t i = 0; i < M; i++) {
      const float complex y = grid[l * R + off + i];
      ls[s * M + i]         = y * conjf(dmrs[s * M + i]);
      rxpow += crealf(y) * crealf(y) + cimagf(y) * cimagf(y);
    }
  }
  double complex dot = 0;
  for (uint32_t i = 0; i < M; i++) dot += (double complex)ls[i] * conj((double complex)ls[M + i]);
  meas[2] = (float)(carg(dot) / (2.0 * M_PI * 0.0005));
  for (uint32_t s = 0; s < 2; s++) {
    const uint32_t l = (s + 1) * nsymb - 4;
    orc_conv_same(&ls[s * M], f, &ce[l * R + off], M, 3);
    float pw = 0;
    for (uint32_t i = 0; i < M; i++) {
      const float complex d = ce[l * R + off + i] - ls[s * M + i];
      pw += crealf(d) * crealf(d) + cimagf(d) * cimagf(d);
    }
    noise += pw / (float)M;
    for (uint32_t k = 0; k < nsymb; k++) {
      const uint32_t dst = s * nsymb + k;
      if (dst != l) memcpy(&ce[dst * R + off], &ce[l * R + off], sizeof(float complex) * M);
    }
  }
  noise /= 2;
  const float a = 7.419 * w * w + 0.1117 * w - 0.005387; /* chest_ul.c:216-219 */
  noise         = noise / (a * 0.8);
  meas[0]       = noise;
  meas[1]       = (rxpow / (float)(2 * M)) / noise;
  meas[3]       = 0;
  free(ls);
  return 0;
}
