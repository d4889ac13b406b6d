/*
 * TEST INFRASTRUCTURE ONLY — flat C entry points over the UNMODIFIED reference sources.
 *
 * This file is compiled together with the reference's own .c files (oracle/Makefile, target `ref`)
 * into oracle/_ref/libsrsref.so.  It contains no algorithm: every function below only allocates the
 * reference's objects and calls the reference's public API, so that tests/ and bench.py's
 * cpu_baseline / --impl reference legs can drive the real thing through ctypes without mirroring the
 * reference's struct layouts in Python.  The product library never links or loads this.
 */
#include <pthread.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "srsran/phy/common/phy_common.h"
#include "srsran/phy/dft/ofdm.h"
#include "srsran/phy/fec/cbsegm.h"
#include "srsran/phy/fec/crc.h"
#include "srsran/phy/fec/turbo/rm_turbo.h"
#include "srsran/phy/fec/turbo/tc_interl.h"
#include "srsran/phy/fec/turbo/turbocoder.h"
#include "srsran/phy/fec/turbo/turbodecoder.h"
#include "srsran/phy/modem/demod_soft.h"
#include "srsran/phy/utils/vector.h"

static pthread_mutex_t g_init_mutex = PTHREAD_MUTEX_INITIALIZER;
static int             g_tables     = 0;

static void ensure_tables(void)
{
  pthread_mutex_lock(&g_init_mutex);
  if (!g_tables) {
    srsran_rm_turbo_gentables(); /* rm_turbo.c:276 */
    g_tables = 1;
  }
  pthread_mutex_unlock(&g_init_mutex);
}

int ref_nof_cb_sizes(void)
{
  return SRSRAN_NOF_TC_CB_SIZES;
}

int ref_cbsize(uint32_t idx)
{
  return srsran_cbsegm_cbsize(idx); /* cbsegm.c */
}

int ref_cbindex(uint32_t K)
{
  return srsran_cbsegm_cbindex(K); /* cbsegm.c:119 */
}

/* out[0..8] = F C K1 K2 K1_idx K2_idx C1 C2 tbs */
int ref_cbsegm(uint32_t tbs, uint32_t* out)
{
  srsran_cbsegm_t s;
  memset(&s, 0, sizeof(s));
  int r  = srsran_cbsegm(&s, tbs);
  out[0] = s.F;
  out[1] = s.C;
  out[2] = s.K1;
  out[3] = s.K2;
  out[4] = s.K1_idx;
  out[5] = s.K2_idx;
  out[6] = s.C1;
  out[7] = s.C2;
  out[8] = s.tbs;
  return r;
}

int ref_interleaver(uint32_t K, uint16_t* fwd, uint16_t* rev)
{
  srsran_tc_interl_t t;
  if (srsran_tc_interl_init(&t, K)) {
    return -1;
  }
  int r = srsran_tc_interl_LTE_gen(&t, K); /* tc_interl_lte.c:64 */
  if (!r) {
    memcpy(fwd, t.forward, sizeof(uint16_t) * K);
    memcpy(rev, t.reverse, sizeof(uint16_t) * K);
  }
  srsran_tc_interl_free(&t);
  return r;
}

uint32_t ref_crc_byte(uint32_t poly, int order, const uint8_t* bytes, int nbits)
{
  srsran_crc_t c;
  srsran_crc_init(&c, poly, order);
  return srsran_crc_checksum_byte(&c, bytes, nbits); /* crc.c:147 */
}

uint32_t ref_crc_bits(uint32_t poly, int order, uint8_t* bits, int nbits)
{
  srsran_crc_t c;
  srsran_crc_init(&c, poly, order);
  return srsran_crc_checksum(&c, bits, nbits); /* crc.c:84 */
}

/* bits: one bit per byte, K of them; out: 3K+12 bits, one per byte (turbocoder.c:77) */
int ref_tcod_encode(uint8_t* bits, uint8_t* out, uint32_t K)
{
  srsran_tcod_t t;
  if (srsran_tcod_init(&t, SRSRAN_TCOD_MAX_LEN_CB)) {
    return -1;
  }
  int r = srsran_tcod_encode(&t, bits, out, K);
  srsran_tcod_free(&t);
  return r;
}

/* coded: 3K+12 bits; produces E bits for redundancy version rv (rm_turbo.c:981).
 * The circular buffer is (re)built with rv 0 first, as the reference requires. */
int ref_rm_tx(uint8_t* coded, uint32_t K, uint8_t* out, uint32_t E, uint32_t rv)
{
  uint32_t wlen = 3 * (K + 4 + 32);
  uint8_t* w    = calloc(wlen, 1);
  uint8_t* tmp  = calloc(E + 16, 1);
  int      r    = srsran_rm_turbo_tx(w, wlen, coded, 3 * K + 12, rv == 0 ? out : tmp, E, 0);
  if (!r && rv != 0) {
    r = srsran_rm_turbo_tx(w, wlen, coded, 3 * K + 12, out, E, rv);
  }
  free(w);
  free(tmp);
  return r;
}

/* soft += dematch(in) ; natural!=0 selects the natural 3*i+j layout (enable_input_tdec=false), rm_turbo.c:403 */
int ref_rm_rx(int16_t* in, int16_t* soft, uint32_t E, uint32_t cb_idx, uint32_t rv, int natural)
{
  ensure_tables();
  return srsran_rm_turbo_rx_lut_(in, soft, E, cb_idx, rv, natural ? false : true);
}

/* table[i] = natural destination of the i-th received value (SURVEY 8c recovery trick) */
int ref_rm_table(uint32_t cb_idx, uint32_t rv, uint16_t* table)
{
  ensure_tables();
  int      K = srsran_cbsegm_cbsize(cb_idx);
  uint32_t n = 3 * K + 12;
  int16_t* a = srsran_vec_i16_malloc(n + 64);
  int16_t* o = srsran_vec_i16_malloc(n + 64);
  /* values are < 2^15 only for n < 32768: true for all K (max 18444) */
  for (uint32_t i = 0; i < n; i++) a[i] = (int16_t)(i + 1);
  memset(o, 0, sizeof(int16_t) * (n + 64));
  int r = srsran_rm_turbo_rx_lut_(a, o, n, cb_idx, rv, false);
  for (uint32_t d = 0; d < n; d++) {
    if (o[d] > 0) table[o[d] - 1] = (uint16_t)d;
  }
  free(a);
  free(o);
  return r;
}

/* Float spec implementation used by rm_turbo_test as the known-answer (rm_turbo.c:1070) */
int ref_rm_rx_float(float* in, uint32_t E, float* out, uint32_t out_len, uint32_t rv)
{
  uint32_t wlen = 3 * (out_len / 3 + 32 + 4);
  float*   w    = malloc(sizeof(float) * wlen);
  for (uint32_t i = 0; i < wlen; i++) w[i] = SRSRAN_RX_NULL;
  int r = srsran_rm_turbo_rx(w, wlen, in, E, out, out_len, rv, 0);
  free(w);
  return r;
}

/*
 * Run exactly npass SISO passes of implementation `impl` (SRSRAN_TDEC_GENERIC=1 is the parity oracle) on one
 * code block in natural layout; after every pass store the decided bytes (K/8 each) into out[pass*K/8 ...].
 */
int ref_tdec_passes(int impl, int16_t* llr, uint32_t K, uint32_t npass, uint8_t* out)
{
  srsran_tdec_t h;
  if (srsran_tdec_init_manual(&h, SRSRAN_TCOD_MAX_LEN_CB, (srsran_tdec_impl_type_t)impl)) {
    return -1;
  }
  srsran_tdec_force_not_sb(&h);
  if (srsran_tdec_new_cb(&h, K)) {
    srsran_tdec_free(&h);
    return -1;
  }
  int16_t* in = srsran_vec_i16_malloc(3 * K + 12 + 64);
  memcpy(in, llr, sizeof(int16_t) * (3 * K + 12));
  for (uint32_t p = 0; p < npass; p++) {
    srsran_tdec_iteration(&h, in, &out[(size_t)p * (K / 8)]); /* turbodecoder.c:527 */
  }
  free(in);
  srsran_tdec_free(&h);
  return 0;
}

typedef struct {
  int       impl;
  int16_t*  llr;
  uint32_t  first, last, K, max_pass;
  int       crc_kind; /* 0: CRC24B over K bits, 1: CRC24A over crc_len bits, 2: none */
  uint32_t  crc_len;
  int       early_stop;
  uint8_t*  out;
  uint8_t*  crc_ok;
  uint8_t*  npass;
  int       err;
} job_t;

/* The per-code-block loop of decode_tb_cb (sch.c:420-454): new_cb, then up to max_pass x (one pass + CRC). */
static void* job_run(void* arg)
{
  job_t*        j = (job_t*)arg;
  srsran_tdec_t h;
  srsran_crc_t  crc;
  uint32_t      K    = j->K;
  size_t        nllr = 3 * (size_t)K + 12;
  if (j->impl == SRSRAN_TDEC_AUTO) {
    j->err = srsran_tdec_init(&h, SRSRAN_TCOD_MAX_LEN_CB);
  } else {
    j->err = srsran_tdec_init_manual(&h, SRSRAN_TCOD_MAX_LEN_CB, (srsran_tdec_impl_type_t)j->impl);
  }
  if (j->err) {
    return NULL;
  }
  srsran_tdec_force_not_sb(&h);
  srsran_crc_init(&crc, j->crc_kind == 1 ? SRSRAN_LTE_CRC24A : SRSRAN_LTE_CRC24B, 24);
  int16_t* in = srsran_vec_i16_malloc(nllr + 64);
  for (uint32_t cb = j->first; cb < j->last; cb++) {
    memcpy(in, &j->llr[cb * nllr], sizeof(int16_t) * nllr);
    uint8_t* data = &j->out[(size_t)cb * (K / 8)];
    srsran_tdec_new_cb(&h, K);
    uint32_t noi  = 0;
    bool     ok   = false;
    bool     stop = false;
    do {
      srsran_tdec_iteration(&h, in, data);
      noi++;
      if (j->crc_kind != 2 && !ok) {
        uint32_t len = j->crc_kind == 1 ? j->crc_len : K;
        if (!srsran_crc_checksum_byte(&crc, data, (int)len)) {
          ok             = true;
          j->npass[cb]   = (uint8_t)noi;
          stop           = j->early_stop != 0;
        }
      }
    } while (noi < j->max_pass && !stop);
    j->crc_ok[cb] = ok ? 1 : 0;
    if (!ok) {
      j->npass[cb] = (uint8_t)noi;
    }
  }
  free(in);
  srsran_tdec_free(&h);
  return NULL;
}

/*
 * Batched decode loop over ncb code blocks of equal K (natural layout, ncb*(3K+12) int16).
 * npass[cb] = pass count at which the CRC first matched (or passes run when it never did).
 * With early_stop==0 every block runs exactly max_pass passes and `out` holds the decision of the last one.
 * Returns elapsed seconds of the threaded region in *seconds (CLOCK_MONOTONIC).
 */
int ref_decode_batch(int      impl,
                     int16_t* llr,
                     uint32_t ncb,
                     uint32_t K,
                     uint32_t max_pass,
                     int      crc_kind,
                     uint32_t crc_len,
                     int      early_stop,
                     uint8_t* out,
                     uint8_t* crc_ok,
                     uint8_t* npass,
                     int      nthreads,
                     double*  seconds)
{
  if (nthreads < 1) nthreads = 1;
  if ((uint32_t)nthreads > ncb) nthreads = (int)(ncb ? ncb : 1);
  job_t*     jobs = calloc((size_t)nthreads, sizeof(job_t));
  pthread_t* th   = calloc((size_t)nthreads, sizeof(pthread_t));
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < nthreads; t++) {
    jobs[t].impl       = impl;
    jobs[t].llr        = llr;
    jobs[t].first      = (uint32_t)(((uint64_t)ncb * t) / nthreads);
    jobs[t].last       = (uint32_t)(((uint64_t)ncb * (t + 1)) / nthreads);
    jobs[t].K          = K;
    jobs[t].max_pass   = max_pass;
    jobs[t].crc_kind   = crc_kind;
    jobs[t].crc_len    = crc_len;
    jobs[t].early_stop = early_stop;
    jobs[t].out        = out;
    jobs[t].crc_ok     = crc_ok;
    jobs[t].npass      = npass;
    pthread_create(&th[t], NULL, job_run, &jobs[t]);
  }
  int err = 0;
  for (int t = 0; t < nthreads; t++) {
    pthread_join(th[t], NULL);
    err |= jobs[t].err;
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (seconds) {
    *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  }
  free(jobs);
  free(th);
  return err;
}

/* ---- OFDM ------------------------------------------------------------------------------------------- */

void ref_use_standard_symbol_size(int enabled)
{
  srsran_use_standard_symbol_size(enabled != 0);
}

int ref_symbol_sz(uint32_t nof_prb)
{
  return srsran_symbol_sz(nof_prb);
}

/*
 * One srsran_ofdm_rx_init_cfg + nsf x srsran_ofdm_rx_sf (ofdm.c:290,453).  in: nsf*sf_sz samples,
 * out: nsf*nof_symbols*2*12*nof_prb.  The reference multiplies the shift into its input buffer in place, so the
 * caller's `in` is copied per subframe into the bound buffer.
 */
int ref_ofdm_rx(uint32_t nof_prb,
                int      cp_ext,
                uint32_t symbol_sz,
                float    freq_shift,
                float    rx_window_offset,
                int      normalize,
                int      keep_dc,
                cf_t*    in,
                cf_t*    out,
                uint32_t nsf,
                double*  seconds)
{
  srsran_ofdm_t     q;
  srsran_ofdm_cfg_t cfg;
  memset(&q, 0, sizeof(q));
  memset(&cfg, 0, sizeof(cfg));
  srsran_cp_t cp   = cp_ext ? SRSRAN_CP_EXT : SRSRAN_CP_NORM;
  uint32_t    N    = symbol_sz ? symbol_sz : (uint32_t)srsran_symbol_sz(nof_prb);
  uint32_t    sfsz = SRSRAN_SF_LEN(N);
  uint32_t    nre  = 2 * SRSRAN_CP_NSYMB(cp) * 12 * nof_prb;
  cf_t*       ib   = srsran_vec_cf_malloc(sfsz + 4096);
  cf_t*       ob   = srsran_vec_cf_malloc(nre);
  /* leave head-room in front of the buffer: the windowed plan reads in_buffer - window_offset_n .. (ofdm.c:160) */
  cfg.nof_prb          = nof_prb;
  cfg.in_buffer        = ib;
  cfg.out_buffer       = ob;
  cfg.cp               = cp;
  cfg.sf_type          = SRSRAN_SF_NORM;
  cfg.normalize        = normalize != 0;
  cfg.freq_shift_f     = freq_shift;
  cfg.rx_window_offset = rx_window_offset;
  cfg.symbol_sz        = symbol_sz;
  cfg.keep_dc          = keep_dc != 0;
  if (srsran_ofdm_rx_init_cfg(&q, &cfg)) {
    return -1;
  }
  struct timespec t0, t1;
  double          acc = 0;
  for (uint32_t s = 0; s < nsf; s++) {
    memcpy(ib, &in[(size_t)s * sfsz], sizeof(cf_t) * sfsz);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    srsran_ofdm_rx_sf(&q);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    acc += (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    memcpy(&out[(size_t)s * nre], ob, sizeof(cf_t) * nre);
  }
  if (seconds) *seconds = acc;
  srsran_ofdm_rx_free(&q);
  free(ib);
  free(ob);
  return 0;
}

/* Tx side, only to synthesise time-domain test input (ofdm.c tx path). in: nsf*nre, out: nsf*sf_sz */
int ref_ofdm_tx(uint32_t nof_prb,
                int      cp_ext,
                uint32_t symbol_sz,
                float    freq_shift,
                int      normalize,
                int      keep_dc,
                cf_t*    in,
                cf_t*    out,
                uint32_t nsf)
{
  srsran_ofdm_t     q;
  srsran_ofdm_cfg_t cfg;
  memset(&q, 0, sizeof(q));
  memset(&cfg, 0, sizeof(cfg));
  srsran_cp_t cp   = cp_ext ? SRSRAN_CP_EXT : SRSRAN_CP_NORM;
  uint32_t    N    = symbol_sz ? symbol_sz : (uint32_t)srsran_symbol_sz(nof_prb);
  uint32_t    sfsz = SRSRAN_SF_LEN(N);
  uint32_t    nre  = 2 * SRSRAN_CP_NSYMB(cp) * 12 * nof_prb;
  cf_t*       ib   = srsran_vec_cf_malloc(nre);
  cf_t*       ob   = srsran_vec_cf_malloc(sfsz);
  cfg.nof_prb      = nof_prb;
  cfg.in_buffer    = ib;
  cfg.out_buffer   = ob;
  cfg.cp           = cp;
  cfg.sf_type      = SRSRAN_SF_NORM;
  cfg.normalize    = normalize != 0;
  cfg.freq_shift_f = freq_shift;
  cfg.symbol_sz    = symbol_sz;
  cfg.keep_dc      = keep_dc != 0;
  if (srsran_ofdm_tx_init_cfg(&q, &cfg)) {
    return -1;
  }
  for (uint32_t s = 0; s < nsf; s++) {
    memcpy(ib, &in[(size_t)s * nre], sizeof(cf_t) * nre);
    srsran_ofdm_tx_sf(&q);
    memcpy(&out[(size_t)s * sfsz], ob, sizeof(cf_t) * sfsz);
  }
  srsran_ofdm_tx_free(&q);
  free(ib);
  free(ob);
  return 0;
}

/* ---- soft demodulation (demod_soft.c:871); mod: 0 BPSK, 1 QPSK, 2 16QAM, 3 64QAM, 4 256QAM --------------- */
int ref_demod_s(int mod, const cf_t* symbols, short* llr, int nsymbols)
{
  return srsran_demod_soft_demodulate_s((srsran_mod_t)mod, symbols, llr, nsymbols);
}

/* ---- PUSCH receive chain beyond OFDM: DMRS, channel estimation, equaliser, transform de-precoding,
 *      descrambling, UL-SCH de-interleave ------------------------------------------------ */
#include "srsran/phy/ch_estimation/chest_ul.h"
#include "srsran/phy/ch_estimation/refsignal_ul.h"
#include "srsran/phy/dft/dft_precoding.h"
#include "srsran/phy/mimo/precoding.h"
#include "srsran/phy/phch/pusch.h"
#include "srsran/phy/phch/ra.h"
#include "srsran/phy/common/sequence.h"

/* sequences.c:139 */
void ref_pusch_seq_apply_s(const int16_t* in, int16_t* out, uint32_t rnti, uint32_t nslot, uint32_t cell_id, uint32_t len)
{
  srsran_sequence_pusch_apply_s(in, out, (uint16_t)rnti, nslot, cell_id, len);
}

/* sch.c:993 (not declared in a header), no RI bits */
void ulsch_deinterleave(int16_t*          q_bits,
                        uint32_t          Qm,
                        uint32_t          H_prime_total,
                        uint32_t          N_pusch_symbs,
                        int16_t*          g_bits,
                        srsran_uci_bit_t* ri_bits,
                        uint32_t          nof_ri_bits,
                        uint8_t*          ri_present,
                        uint32_t*         inteleaver_lut);

int ref_ulsch_deinterleave(int16_t* q_bits, int16_t* g_bits, uint32_t Qm, uint32_t H_prime_total, uint32_t N_pusch_symbs)
{
  uint32_t  n   = H_prime_total * Qm;
  uint8_t*  tmp = calloc(n, 1);
  uint32_t* lut = calloc(n, sizeof(uint32_t));
  if (!tmp || !lut) return -1;
  ulsch_deinterleave(q_bits, Qm, H_prime_total, N_pusch_symbs, g_bits, NULL, 0, tmp, lut);
  free(tmp);
  free(lut);
  return 0;
}

/* dft_precoding.c:114; is_tx selects the forward plan (precoding) or the backward one (receiver) */
int ref_dft_precoding(cf_t* in, cf_t* out, uint32_t nof_prb, uint32_t nof_symbols, int is_tx)
{
  srsran_dft_precoding_t q;
  if (srsran_dft_precoding_init(&q, nof_prb, is_tx != 0)) return -1;
  int r = srsran_dft_precoding(&q, in, out, nof_prb, nof_symbols);
  srsran_dft_precoding_free(&q);
  return r;
}

/* precoding.c:357 */
int ref_predecoding_single(cf_t* y, cf_t* h, cf_t* x, int nof_symbols, float scaling, float noise_estimate)
{
  return srsran_predecoding_single(y, h, x, NULL, nof_symbols, scaling, noise_estimate);
}

/* Link parameters, all uint32:
 *  0 cell_id  1 cell nof_prb  2 cp_ext  3 dmrs cyclic_shift  4 delta_ss  5 group_hopping  6 sequence_hopping
 *  7 rnti  8 tti  9 L_prb  10 n_prb  11 mod (srsran_mod_t)  12 tbs  13 rv  14 n_dmrs (cyclic shift for DMRS, 0..7)
 *  15 max_nof_iterations  16 shortened (the subframe's last symbol carries the SRS: srsran_ul_sf_cfg_t.shortened) */
enum { P_CELL_ID, P_NOF_PRB, P_CP_EXT, P_CSHIFT, P_DELTA_SS, P_GH, P_SH, P_RNTI, P_TTI, P_L_PRB, P_N_PRB, P_MOD, P_TBS, P_RV,
       P_N_DMRS, P_MAX_ITER, P_SHORTENED, P_COUNT };

static srsran_cell_t link_cell(const uint32_t* p)
{
  srsran_cell_t cell;
  memset(&cell, 0, sizeof(cell));
  cell.nof_prb         = p[P_NOF_PRB];
  cell.nof_ports       = 1;
  cell.id              = p[P_CELL_ID];
  cell.cp              = p[P_CP_EXT] ? SRSRAN_CP_EXT : SRSRAN_CP_NORM;
  cell.phich_length    = SRSRAN_PHICH_NORM;
  cell.phich_resources = SRSRAN_PHICH_R_1;
  cell.frame_type      = SRSRAN_FDD;
  return cell;
}

static void link_cfg(const uint32_t* p, srsran_cell_t* cell, srsran_pusch_cfg_t* cfg, srsran_ul_sf_cfg_t* sf,
                     srsran_refsignal_dmrs_pusch_cfg_t* dmrs)
{
  memset(cfg, 0, sizeof(*cfg));
  memset(sf, 0, sizeof(*sf));
  memset(dmrs, 0, sizeof(*dmrs));
  sf->tti                   = p[P_TTI];
  sf->shortened             = p[P_SHORTENED] != 0;
  dmrs->cyclic_shift        = p[P_CSHIFT];
  dmrs->delta_ss            = p[P_DELTA_SS];
  dmrs->group_hopping_en    = p[P_GH] != 0;
  dmrs->sequence_hopping_en = p[P_SH] != 0;
  cfg->rnti                 = (uint16_t)p[P_RNTI];
  cfg->grant.L_prb          = p[P_L_PRB];
  cfg->grant.n_prb[0] = cfg->grant.n_prb[1] = p[P_N_PRB];
  cfg->grant.n_prb_tilde[0] = cfg->grant.n_prb_tilde[1] = p[P_N_PRB];
  cfg->grant.tb.mod   = (srsran_mod_t)p[P_MOD];
  cfg->grant.tb.tbs   = (int)p[P_TBS];
  cfg->grant.tb.rv    = (int)p[P_RV];
  cfg->grant.n_dmrs   = p[P_N_DMRS];
  srsran_ra_ul_compute_nof_re(&cfg->grant, cell->cp, p[P_SHORTENED] ? 1 : 0); /* ra_ul.c:230 */
  cfg->max_nof_iterations = p[P_MAX_ITER];
  cfg->enable_64qam       = true;
}

/* refsignal_ul.c:337: r[2 * 12 * L_prb], slot-major */
int ref_dmrs_pusch_gen(const uint32_t* p, cf_t* r)
{
  srsran_cell_t                     cell = link_cell(p);
  srsran_pusch_cfg_t                cfg;
  srsran_ul_sf_cfg_t                sf;
  srsran_refsignal_dmrs_pusch_cfg_t dmrs;
  srsran_refsignal_ul_t             q;
  link_cfg(p, &cell, &cfg, &sf, &dmrs);
  memset(&q, 0, sizeof(q));
  if (srsran_refsignal_ul_set_cell(&q, cell)) return -1;
  return srsran_refsignal_dmrs_pusch_gen(&q, &dmrs, p[P_L_PRB], p[P_TTI] % 10, p[P_N_DMRS], r);
}

/* chest_ul.c:370: grid = 2*nsymb*12*nof_prb RE of one subframe; ce_out same size; meas = {noise_estimate, snr, cfo_hz, ta_us} */
int ref_chest_ul_pusch(const uint32_t* p, cf_t* grid, cf_t* ce_out, float* meas)
{
  srsran_cell_t                     cell = link_cell(p);
  srsran_pusch_cfg_t                cfg;
  srsran_ul_sf_cfg_t                sf;
  srsran_refsignal_dmrs_pusch_cfg_t dmrs;
  srsran_chest_ul_t                 q;
  srsran_chest_ul_res_t             res;
  link_cfg(p, &cell, &cfg, &sf, &dmrs);
  if (srsran_chest_ul_init(&q, cell.nof_prb)) return -1;
  if (srsran_chest_ul_set_cell(&q, cell)) return -2;
  srsran_chest_ul_pregen(&q, &dmrs, NULL);
  if (srsran_chest_ul_res_init(&res, cell.nof_prb)) return -3;
  memset(res.ce, 0, sizeof(cf_t) * res.nof_re);
  int r = srsran_chest_ul_estimate_pusch(&q, &sf, &cfg, grid, &res);
  memcpy(ce_out, res.ce, sizeof(cf_t) * 2 * SRSRAN_CP_NSYMB(cell.cp) * 12 * cell.nof_prb);
  meas[0] = res.noise_estimate;
  meas[1] = res.snr;
  meas[2] = res.cfo_hz;
  meas[3] = res.ta_us;
  srsran_chest_ul_res_free(&res);
  srsran_chest_ul_free(&q);
  return r;
}

/* Transmit side, to synthesise test input: srsran_pusch_encode (pusch.c) + DMRS (refsignal_ul.c:192) into a zeroed grid */
int ref_pusch_encode(const uint32_t* p, uint8_t* data, cf_t* grid)
{
  srsran_cell_t                     cell = link_cell(p);
  srsran_pusch_cfg_t                cfg;
  srsran_ul_sf_cfg_t                sf;
  srsran_refsignal_dmrs_pusch_cfg_t dmrs;
  srsran_pusch_t                    tx;
  srsran_softbuffer_tx_t            sb;
  srsran_refsignal_ul_t             rs;
  link_cfg(p, &cell, &cfg, &sf, &dmrs);
  ensure_tables();
  if (srsran_pusch_init_ue(&tx, cell.nof_prb)) return -1;
  if (srsran_pusch_set_cell(&tx, cell)) return -2;
  if (srsran_softbuffer_tx_init(&sb, cell.nof_prb)) return -3;
  srsran_softbuffer_tx_reset(&sb);
  cfg.softbuffers.tx = &sb;
  srsran_pusch_data_t pdata;
  memset(&pdata, 0, sizeof(pdata));
  pdata.ptr = data;
  uint32_t nre = 2 * SRSRAN_CP_NSYMB(cell.cp) * 12 * cell.nof_prb;
  memset(grid, 0, sizeof(cf_t) * nre);
  int r = srsran_pusch_encode(&tx, &sf, &cfg, &pdata, grid);
  if (r == 0) {
    cf_t* rp = srsran_vec_cf_malloc(2 * 12 * p[P_L_PRB]);
    memset(&rs, 0, sizeof(rs));
    if (srsran_refsignal_ul_set_cell(&rs, cell)) r = -4;
    else if (srsran_refsignal_dmrs_pusch_gen(&rs, &dmrs, p[P_L_PRB], p[P_TTI] % 10, p[P_N_DMRS], rp)) r = -5;
    else srsran_refsignal_dmrs_pusch_put(&rs, &cfg, rp, grid);
    free(rp);
  }
  srsran_softbuffer_tx_free(&sb);
  srsran_pusch_free(&tx);
  return r;
}

/* Receive side: chest (unless use_identity_ce) + srsran_pusch_decode (pusch.c:358).  Besides the payload it hands out the
 * object's intermediate buffers: d = de-precoded symbols (nof_re), q = descrambled soft bits, g = de-interleaved soft bits
 * (nof_bits each), so a pipeline can be compared stage by stage. */
int ref_pusch_decode(const uint32_t* p, cf_t* grid, int use_identity_ce, uint8_t* data, int* crc_ok, float* meas, cf_t* d_out,
                     int16_t* q_out, int16_t* g_out, cf_t* ce_out)
{
  srsran_cell_t                     cell = link_cell(p);
  srsran_pusch_cfg_t                cfg;
  srsran_ul_sf_cfg_t                sf;
  srsran_refsignal_dmrs_pusch_cfg_t dmrs;
  srsran_pusch_t                    rx;
  srsran_softbuffer_rx_t            sb;
  srsran_chest_ul_t                 chest;
  srsran_chest_ul_res_t             res;
  link_cfg(p, &cell, &cfg, &sf, &dmrs);
  ensure_tables();
  if (srsran_pusch_init_enb(&rx, cell.nof_prb)) return -1;
  if (srsran_pusch_set_cell(&rx, cell)) return -2;
  if (srsran_softbuffer_rx_init(&sb, cell.nof_prb)) return -3;
  srsran_softbuffer_rx_reset(&sb);
  cfg.softbuffers.rx = &sb;
  if (srsran_chest_ul_res_init(&res, cell.nof_prb)) return -4;
  if (use_identity_ce) {
    srsran_chest_ul_res_set_identity(&res);
    res.noise_estimate = 0;
  } else {
    if (srsran_chest_ul_init(&chest, cell.nof_prb)) return -5;
    if (srsran_chest_ul_set_cell(&chest, cell)) return -6;
    srsran_chest_ul_pregen(&chest, &dmrs, NULL);
    memset(res.ce, 0, sizeof(cf_t) * res.nof_re);
    if (srsran_chest_ul_estimate_pusch(&chest, &sf, &cfg, grid, &res)) return -7;
    srsran_chest_ul_free(&chest);
  }
  if (meas) {
    meas[0] = res.noise_estimate;
    meas[1] = res.snr;
    meas[2] = res.cfo_hz;
    meas[3] = res.ta_us;
  }
  if (ce_out) memcpy(ce_out, res.ce, sizeof(cf_t) * 2 * SRSRAN_CP_NSYMB(cell.cp) * 12 * cell.nof_prb);
  srsran_pusch_res_t out;
  memset(&out, 0, sizeof(out));
  out.data = data;
  int r    = srsran_pusch_decode(&rx, &sf, &cfg, &res, grid, &out);
  if (crc_ok) *crc_ok = out.crc ? 1 : 0;
  if (meas) meas[4] = out.avg_iterations_block;
  if (d_out) memcpy(d_out, rx.d, sizeof(cf_t) * cfg.grant.nof_re);
  if (q_out) memcpy(q_out, rx.q, sizeof(int16_t) * cfg.grant.tb.nof_bits);
  if (g_out) memcpy(g_out, rx.g, sizeof(int16_t) * cfg.grant.tb.nof_bits);
  srsran_chest_ul_res_free(&res);
  srsran_softbuffer_rx_free(&sb);
  srsran_pusch_free(&rx);
  return r;
}

/* CPU baseline of the PUSCH receive chain after the OFDM demodulator (FFTW is not available here): nthreads workers, each
 * with its own srsran_chest_ul_t / srsran_pusch_t / soft buffer like one cc_worker, run chest + srsran_pusch_decode over
 * their share of nsf subframes (all the same link parameters, grids[nsf][2*nsymb*12*nof_prb]).  Object set-up is outside the
 * timed region.  ok_out[nsf] = CRC verdicts. */
typedef struct {
  const uint32_t* p;
  cf_t*           grids;
  uint32_t        first, count;
  uint8_t*        ok;
  int             err;
  pthread_barrier_t* bar;
} pusch_job_t;

static void* pusch_job_run(void* arg)
{
  pusch_job_t*                      j    = (pusch_job_t*)arg;
  srsran_cell_t                     cell = link_cell(j->p);
  srsran_pusch_cfg_t                cfg;
  srsran_ul_sf_cfg_t                sf;
  srsran_refsignal_dmrs_pusch_cfg_t dmrs;
  srsran_pusch_t                    rx;
  srsran_softbuffer_rx_t            sb;
  srsran_chest_ul_t                 chest;
  srsran_chest_ul_res_t             res;
  link_cfg(j->p, &cell, &cfg, &sf, &dmrs);
  uint32_t nre  = 2 * SRSRAN_CP_NSYMB(cell.cp) * 12 * cell.nof_prb;
  uint8_t* data = calloc(cfg.grant.tb.tbs / 8 + 64, 1);
  j->err        = 0;
  if (srsran_pusch_init_enb(&rx, cell.nof_prb) || srsran_pusch_set_cell(&rx, cell) || srsran_softbuffer_rx_init(&sb, cell.nof_prb) ||
      srsran_chest_ul_init(&chest, cell.nof_prb) || srsran_chest_ul_set_cell(&chest, cell) || srsran_chest_ul_res_init(&res, cell.nof_prb)) {
    j->err = -1;
  } else {
    srsran_chest_ul_pregen(&chest, &dmrs, NULL);
  }
  pthread_barrier_wait(j->bar); /* start of the timed region */
  if (!j->err) {
    for (uint32_t s = j->first; s < j->first + j->count; s++) {
      srsran_softbuffer_rx_reset(&sb);
      cfg.softbuffers.rx = &sb;
      srsran_pusch_res_t out;
      memset(&out, 0, sizeof(out));
      out.data = data;
      if (srsran_chest_ul_estimate_pusch(&chest, &sf, &cfg, &j->grids[(size_t)s * nre], &res) ||
          srsran_pusch_decode(&rx, &sf, &cfg, &res, &j->grids[(size_t)s * nre], &out)) {
        j->err = -2;
        break;
      }
      j->ok[s] = out.crc ? 1 : 0;
    }
  }
  pthread_barrier_wait(j->bar); /* end of the timed region */
  free(data);
  return NULL;
}

int ref_pusch_rx_bench(const uint32_t* p, cf_t* grids, uint32_t nsf, int nthreads, uint8_t* ok_out, double* seconds)
{
  if (nthreads < 1) nthreads = 1;
  ensure_tables();
  pthread_t*        th   = calloc(nthreads, sizeof(pthread_t));
  pusch_job_t*      jobs = calloc(nthreads, sizeof(pusch_job_t));
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, NULL, nthreads + 1);
  uint32_t per = (nsf + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; t++) {
    uint32_t first = (uint32_t)t * per;
    jobs[t].p      = p;
    jobs[t].grids  = grids;
    jobs[t].first  = first < nsf ? first : nsf;
    jobs[t].count  = first < nsf ? (first + per <= nsf ? per : nsf - first) : 0;
    jobs[t].ok     = ok_out;
    jobs[t].bar    = &bar;
    pthread_create(&th[t], NULL, pusch_job_run, &jobs[t]);
  }
  struct timespec t0, t1;
  pthread_barrier_wait(&bar);
  clock_gettime(CLOCK_MONOTONIC, &t0);
  pthread_barrier_wait(&bar);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  int err = 0;
  for (int t = 0; t < nthreads; t++) {
    pthread_join(th[t], NULL);
    if (jobs[t].err) err = jobs[t].err;
  }
  if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  pthread_barrier_destroy(&bar);
  free(th);
  free(jobs);
  return err;
}

/* ---------------------------------------------------------------------------------------------------------------
 * PUSCH with control information multiplexed into it (TS 36.212 5.2.2.6-5.2.2.8; sch.c:1022-1195, uci.c).
 * UCI parameters, all uint32:
 *  0 nof_ack (uci_cfg.ack[0].nof_acks)  1 ack bits (bit a = ack_value[a])  2 ri_len  3 ri
 *  4 cqi kind: 0 none, 1 wideband (4 bits), 2 wideband + PMI, 2 ports, rank 1 (6 bits), 3 higher-layer subband (4 + 2N bits)
 *  5 N (kind 3)  6 wideband cqi  7 subband differential cqi (kind 3) / pmi (kind 2)
 *  8 I_offset_ack  9 I_offset_ri  10 I_offset_cqi */
enum { U_NOF_ACK, U_ACK_BITS, U_RI_LEN, U_RI, U_CQI_KIND, U_CQI_N, U_CQI_WB, U_CQI_SB, U_IOFF_ACK, U_IOFF_RI, U_IOFF_CQI, U_COUNT };

static void link_uci(const uint32_t* u, srsran_pusch_cfg_t* cfg, srsran_uci_value_t* val)
{
  memset(val, 0, sizeof(*val));
  cfg->uci_cfg.ack[0].nof_acks = u[U_NOF_ACK];
  for (uint32_t a = 0; a < u[U_NOF_ACK]; a++) val->ack.ack_value[a] = (u[U_ACK_BITS] >> a) & 1u;
  cfg->uci_cfg.cqi.ri_len = u[U_RI_LEN];
  val->ri                 = (uint8_t)u[U_RI];
  if (u[U_CQI_KIND]) {
    cfg->uci_cfg.cqi.data_enable = true;
    if (u[U_CQI_KIND] == 3) {
      cfg->uci_cfg.cqi.type                   = SRSRAN_CQI_TYPE_SUBBAND_HL;
      cfg->uci_cfg.cqi.N                      = u[U_CQI_N];
      val->cqi.subband_hl.wideband_cqi_cw0     = (uint8_t)u[U_CQI_WB];
      val->cqi.subband_hl.subband_diff_cqi_cw0 = u[U_CQI_SB];
    } else {
      cfg->uci_cfg.cqi.type          = SRSRAN_CQI_TYPE_WIDEBAND;
      cfg->uci_cfg.cqi.pmi_present   = u[U_CQI_KIND] == 2;
      val->cqi.wideband.wideband_cqi = (uint8_t)u[U_CQI_WB];
      val->cqi.wideband.pmi          = (uint8_t)u[U_CQI_SB];
    }
  }
  cfg->uci_offset.I_offset_ack = u[U_IOFF_ACK];
  cfg->uci_offset.I_offset_ri  = u[U_IOFF_RI];
  cfg->uci_offset.I_offset_cqi = u[U_IOFF_CQI];
}

/* srsran_pusch_encode with pdata.uci set + DMRS, like ref_pusch_encode */
int ref_pusch_encode_uci(const uint32_t* p, const uint32_t* u, uint8_t* data, cf_t* grid)
{
  srsran_cell_t                     cell = link_cell(p);
  srsran_pusch_cfg_t                cfg;
  srsran_ul_sf_cfg_t                sf;
  srsran_refsignal_dmrs_pusch_cfg_t dmrs;
  srsran_pusch_t                    tx;
  srsran_softbuffer_tx_t            sb;
  srsran_refsignal_ul_t             rs;
  srsran_pusch_data_t               pdata;
  link_cfg(p, &cell, &cfg, &sf, &dmrs);
  memset(&pdata, 0, sizeof(pdata));
  link_uci(u, &cfg, &pdata.uci);
  ensure_tables();
  if (srsran_pusch_init_ue(&tx, cell.nof_prb)) return -1;
  if (srsran_pusch_set_cell(&tx, cell)) return -2;
  if (srsran_softbuffer_tx_init(&sb, cell.nof_prb)) return -3;
  srsran_softbuffer_tx_reset(&sb);
  cfg.softbuffers.tx = &sb;
  pdata.ptr          = data;
  uint32_t nre       = 2 * SRSRAN_CP_NSYMB(cell.cp) * 12 * cell.nof_prb;
  memset(grid, 0, sizeof(cf_t) * nre);
  int r = srsran_pusch_encode(&tx, &sf, &cfg, &pdata, grid);
  if (r == 0) {
    cf_t* rp = srsran_vec_cf_malloc(2 * 12 * p[P_L_PRB]);
    memset(&rs, 0, sizeof(rs));
    if (srsran_refsignal_ul_set_cell(&rs, cell)) r = -4;
    else if (srsran_refsignal_dmrs_pusch_gen(&rs, &dmrs, p[P_L_PRB], p[P_TTI] % 10, p[P_N_DMRS], rp)) r = -5;
    else srsran_refsignal_dmrs_pusch_put(&rs, &cfg, rp, grid);
    free(rp);
  }
  srsran_softbuffer_tx_free(&sb);
  srsran_pusch_free(&tx);
  return r;
}

/* chest + srsran_pusch_decode with the same UCI configuration.  uci_out (int32):
 *  0..9 ack_value  10 ack.valid  11 ri  12 cqi.data_crc  13 number of packed bits  14.. the cqi payload bits (srsran_cqi_value_pack of what was
 *  decoded; room for 14 + 200 values)
 * q_out: the object's q->q after the call (descrambled, HARQ-ACK positions zeroed), g_out: q->g */
int ref_pusch_decode_uci(const uint32_t* p, const uint32_t* u, cf_t* grid, int use_identity_ce, uint8_t* data, int* crc_ok, float* meas,
                         int16_t* q_out, int16_t* g_out, int32_t* uci_out)
{
  srsran_cell_t                     cell = link_cell(p);
  srsran_pusch_cfg_t                cfg;
  srsran_ul_sf_cfg_t                sf;
  srsran_refsignal_dmrs_pusch_cfg_t dmrs;
  srsran_pusch_t                    rx;
  srsran_softbuffer_rx_t            sb;
  srsran_chest_ul_t                 chest;
  srsran_chest_ul_res_t             res;
  srsran_uci_value_t                sent;
  link_cfg(p, &cell, &cfg, &sf, &dmrs);
  link_uci(u, &cfg, &sent);
  ensure_tables();
  if (srsran_pusch_init_enb(&rx, cell.nof_prb)) return -1;
  if (srsran_pusch_set_cell(&rx, cell)) return -2;
  if (srsran_softbuffer_rx_init(&sb, cell.nof_prb)) return -3;
  srsran_softbuffer_rx_reset(&sb);
  cfg.softbuffers.rx = &sb;
  if (srsran_chest_ul_res_init(&res, cell.nof_prb)) return -4;
  if (use_identity_ce) {
    srsran_chest_ul_res_set_identity(&res);
    res.noise_estimate = 0;
  } else {
    if (srsran_chest_ul_init(&chest, cell.nof_prb)) return -5;
    if (srsran_chest_ul_set_cell(&chest, cell)) return -6;
    srsran_chest_ul_pregen(&chest, &dmrs, NULL);
    memset(res.ce, 0, sizeof(cf_t) * res.nof_re);
    if (srsran_chest_ul_estimate_pusch(&chest, &sf, &cfg, grid, &res)) return -7;
    srsran_chest_ul_free(&chest);
  }
  if (meas) {
    meas[0] = res.noise_estimate;
    meas[1] = res.snr;
    meas[2] = res.cfo_hz;
    meas[3] = res.ta_us;
  }
  srsran_pusch_res_t out;
  memset(&out, 0, sizeof(out));
  out.data = data;
  memset(out.uci.ack.ack_value, 2, SRSRAN_UCI_MAX_ACK_BITS);
  int r = srsran_pusch_decode(&rx, &sf, &cfg, &res, grid, &out);
  if (crc_ok) *crc_ok = out.crc ? 1 : 0;
  if (meas) meas[4] = out.avg_iterations_block;
  if (q_out) memcpy(q_out, rx.q, sizeof(int16_t) * cfg.grant.tb.nof_bits);
  if (g_out) memcpy(g_out, rx.g, sizeof(int16_t) * cfg.grant.tb.nof_bits);
  for (int a = 0; a < 10; a++) uci_out[a] = out.uci.ack.ack_value[a];
  uci_out[10] = out.uci.ack.valid ? 1 : 0;
  uci_out[11] = out.uci.ri;
  uci_out[12] = out.uci.cqi.data_crc ? 1 : 0;
  uci_out[13] = 0;
  if (cfg.uci_cfg.cqi.data_enable) {
    /* the packer writes a second codeword once the decoded RI says rank > 1 (cqi.c: higher-layer subband reports), i.e. up to
     * 2 * (4 + 2N) bits: more than SRSRAN_CQI_MAX_BITS for N > 14 -- give it room; the caller compares the first cqi_len bits */
    uint8_t buff[512];
    memset(buff, 0, sizeof(buff));
    int n       = srsran_cqi_value_pack(&cfg.uci_cfg.cqi, &out.uci.cqi, buff);
    if (n > 200) n = 200;
    uci_out[13] = n;
    for (int i = 0; i < n; i++) uci_out[14 + i] = buff[i];
  }
  srsran_chest_ul_res_free(&res);
  srsran_softbuffer_rx_free(&sb);
  srsran_pusch_free(&rx);
  return r;
}

/* uci.c:395-427 through its exported wrappers (K_segm = the argument named tbs there) */
uint32_t ref_qprime_ack(uint32_t L_prb, uint32_t nof_symbols, uint32_t K_segm, uint32_t nof_ack, float beta)
{
  return srsran_qprime_ack_ext(L_prb, nof_symbols, K_segm, nof_ack, beta);
}

/* The two field decoders on their own (uci.c:289-330,637-713).  q_bits for the HARQ-ACK / RI form is the whole descrambled
 * subframe (the function finds the field's positions itself), c_seq its scrambling sequence, one bit per byte; for the CQI form
 * it is the front of the de-interleaved stream.  K_segm is set the way srsran_ulsch_decode does (sch.c:1136). */
static void uci_only_cfg(const uint32_t* p, srsran_pusch_cfg_t* cfg)
{
  srsran_cell_t                     cell = link_cell(p);
  srsran_ul_sf_cfg_t                sf;
  srsran_refsignal_dmrs_pusch_cfg_t dmrs;
  srsran_cbsegm_t                   seg;
  link_cfg(p, &cell, cfg, &sf, &dmrs);
  srsran_cbsegm(&seg, p[P_TBS]);
  cfg->K_segm = seg.C1 * seg.K1 + seg.C2 * seg.K2;
}

int ref_uci_decode_ack_ri(const uint32_t* p, int16_t* q_bits, uint8_t* c_seq, float beta, uint32_t nof_bits, int is_ri, uint8_t* data,
                          int* valid)
{
  srsran_pusch_cfg_t cfg;
  uci_only_cfg(p, &cfg);
  uint32_t          Qm   = srsran_mod_bits_x_symbol(cfg.grant.tb.mod);
  srsran_uci_bit_t* bits = calloc(4 * 12 * 110 * 8, sizeof(srsran_uci_bit_t));
  bool              v    = false;
  int r = srsran_uci_decode_ack_ri(&cfg, q_bits, c_seq, beta, cfg.grant.tb.nof_bits / Qm, 0, bits, data, is_ri ? NULL : &v, nof_bits, is_ri != 0);
  if (valid) *valid = v ? 1 : 0;
  free(bits);
  return r;
}

int ref_uci_decode_cqi(const uint32_t* p, int16_t* q_bits, float beta, uint32_t Q_prime_ri, uint32_t cqi_len, uint8_t* data, int* crc)
{
  srsran_pusch_cfg_t     cfg;
  srsran_uci_cqi_pusch_t q;
  uci_only_cfg(p, &cfg);
  if (srsran_uci_cqi_init(&q)) return -100;
  bool ok = false;
  int  r  = srsran_uci_decode_cqi_pusch(&q, &cfg, q_bits, beta, Q_prime_ri, cqi_len, data, &ok);
  if (crc) *crc = ok ? 1 : 0;
  srsran_uci_cqi_free(&q);
  return r;
}
