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
