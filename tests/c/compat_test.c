/*
 * Exercises the reference-named C API exported by libsrslte_b200.so the way the reference's own unit tests do
 * (turbodecoder_test.c:200-300, rm_turbo_test.c:100-190, ofdm_test.c:100-180) and checks every result against the CPU
 * oracle (liboracle_port.so).  Plain C, includes only include/srslte_b200_srsran_api.h.  Exit code 0 = all equal.
 */
#include <complex.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "srslte_b200_srsran_api.h"

/* oracle (test infrastructure) */
int      orc_tcod_encode(const uint8_t* bits, uint8_t* out, uint32_t K);
int      orc_tdec_passes(const int16_t* llr, uint32_t K, uint32_t npass, uint8_t* out);
int      orc_rm_rx(const int16_t* in, int16_t* soft, uint32_t E, uint32_t cb_idx, uint32_t rv);
int      orc_ofdm_rx(uint32_t nof_prb, int cp_ext, uint32_t symbol_sz, float freq_shift, float rx_window_offset, int normalize,
                     int keep_dc, const cf_t* in, cf_t* out, uint32_t nsf);
uint32_t orc_crc24(int kind, const uint8_t* bytes, int nbits);
int      orc_cbsegm(uint32_t tbs, uint32_t* out);
int      orc_dft_precoding(const float _Complex* in, float _Complex* out, uint32_t nof_prb, uint32_t nof_symbols, int is_tx);

static int fails = 0;
#define CHECK(c, ...)                                                                                                  \
  do {                                                                                                                 \
    if (!(c)) {                                                                                                        \
      fails++;                                                                                                         \
      printf("FAIL %s:%d: ", __FILE__, __LINE__);                                                                      \
      printf(__VA_ARGS__);                                                                                             \
      printf("\n");                                                                                                    \
    }                                                                                                                  \
  } while (0)

static float gauss(void)
{
  float u1 = (rand() + 1.0f) / (RAND_MAX + 2.0f), u2 = rand() / (float)RAND_MAX;
  return sqrtf(-2 * logf(u1)) * cosf(6.2831853f * u2);
}

static void test_tdec(uint32_t K, float sigma)
{
  uint8_t* bits  = malloc(K);
  uint8_t* coded = malloc(3 * K + 12);
  int16_t* llr   = malloc(sizeof(int16_t) * (3 * K + 12));
  uint8_t  out[768], want[8 * 768];
  for (uint32_t i = 0; i < K; i++) bits[i] = rand() & 1;
  orc_tcod_encode(bits, coded, K);
  for (uint32_t i = 0; i < 3 * K + 12; i++) {
    float v = 16.0f * ((coded[i] ? 1.0f : -1.0f) + sigma * gauss());
    llr[i]  = (int16_t)fmaxf(-31, fminf(31, rintf(v)));
  }
  orc_tdec_passes(llr, K, 8, want);

  srsran_tdec_t h;
  CHECK(srsran_tdec_init(&h, SRSRAN_TCOD_MAX_LEN_CB) == SRSRAN_SUCCESS, "tdec init");
  srsran_tdec_force_not_sb(&h);
  /* iteration by iteration, as decode_tb_cb drives it (sch.c:420-454) */
  CHECK(srsran_tdec_new_cb(&h, K) == 0, "new_cb");
  for (int p = 0; p < 8; p++) {
    srsran_tdec_iteration(&h, llr, out);
    CHECK(memcmp(out, &want[p * (K / 8)], K / 8) == 0, "K=%u pass %d decision differs", K, p);
    CHECK(srsran_tdec_get_nof_iterations(&h) == p + 1, "n_iter");
  }
  /* run_all, as turbodecoder_test.c:267-273 */
  for (uint32_t n = 1; n <= 8; n += 3) {
    CHECK(srsran_tdec_run_all(&h, llr, out, n, K) == SRSRAN_SUCCESS, "run_all");
    CHECK(memcmp(out, &want[(n - 1) * (K / 8)], K / 8) == 0, "K=%u run_all %u differs", K, n);
  }
  CHECK(srsran_tdec_new_cb(&h, K + 1) == -1, "invalid K must be rejected");
  CHECK(srsran_tdec_autoimp_get_subblocks(K) == 0, "natural layout");
  srsran_tdec_free(&h);
  free(bits);
  free(coded);
  free(llr);
}

static void test_rm(uint32_t cb_idx, uint32_t rv, float frac)
{
  int      K = srsran_cbsegm_cbsize(cb_idx);
  uint32_t n = 3 * K + 12, E = (uint32_t)(frac * n);
  int16_t* e = malloc(sizeof(int16_t) * E);
  int16_t* a = calloc(n + 64, sizeof(int16_t));
  int16_t* b = calloc(n + 64, sizeof(int16_t));
  for (uint32_t i = 0; i < E; i++) e[i] = (int16_t)(rand() % 200 - 100);
  for (uint32_t i = 0; i < n; i++) a[i] = b[i] = (int16_t)(rand() % 2000 - 1000);
  CHECK(srsran_rm_turbo_rx_lut(e, a, E, cb_idx, rv) == 0, "rm rx");
  orc_rm_rx(e, b, E, cb_idx, rv);
  CHECK(memcmp(a, b, sizeof(int16_t) * n) == 0, "rm cb_idx=%u rv=%u E=%u differs", cb_idx, rv, E);
  CHECK(srsran_rm_turbo_rx_lut(e, a, E, 188, rv) == SRSRAN_ERROR_INVALID_INPUTS, "invalid cb_idx");
  free(e);
  free(a);
  free(b);
}

static void test_ofdm(uint32_t prb, uint32_t N, float shift, float off)
{
  srsran_ofdm_t     q;
  srsran_ofdm_cfg_t cfg;
  memset(&q, 0, sizeof(q));
  memset(&cfg, 0, sizeof(cfg));
  uint32_t n    = N ? N : (uint32_t)srsran_symbol_sz(prb);
  cf_t*    in   = malloc(sizeof(cf_t) * 15 * n);
  cf_t*    out  = malloc(sizeof(cf_t) * 14 * 12 * prb);
  cf_t*    want = malloc(sizeof(cf_t) * 14 * 12 * prb);
  for (uint32_t i = 0; i < 15 * n; i++) in[i] = gauss() + I * gauss();
  cfg.nof_prb          = prb;
  cfg.in_buffer        = in;
  cfg.out_buffer       = out;
  cfg.cp               = SRSRAN_CP_NORM;
  cfg.freq_shift_f     = shift;
  cfg.rx_window_offset = off;
  cfg.symbol_sz        = N;
  CHECK(srsran_ofdm_rx_init_cfg(&q, &cfg) == SRSRAN_SUCCESS, "ofdm init");
  CHECK(q.sf_sz == 15 * n && q.nof_re == 12 * prb && q.nof_symbols == 7, "ofdm geometry");
  orc_ofdm_rx(prb, 0, N, shift, off, 0, 0, in, want, 1);
  srsran_ofdm_rx_sf(&q);
  double num = 0, den = 0;
  for (uint32_t i = 0; i < 14 * 12 * prb; i++) {
    num += pow(cabsf(out[i] - want[i]), 2);
    den += pow(cabsf(want[i]), 2);
  }
  CHECK(sqrt(num / den) < 1e-4, "ofdm prb=%u N=%u rel err %g", prb, n, sqrt(num / den));
  srsran_ofdm_rx_free(&q);
  free(in);
  free(out);
  free(want);
}

static void test_dft(int N, int backward)
{
  srsran_dft_plan_t p;
  cf_t*             in  = malloc(sizeof(cf_t) * N);
  cf_t*             out = malloc(sizeof(cf_t) * N);
  for (int i = 0; i < N; i++) in[i] = gauss() + I * gauss();
  CHECK(srsran_dft_plan_c(&p, N, backward ? SRSRAN_DFT_BACKWARD : SRSRAN_DFT_FORWARD) == 0, "dft plan");
  srsran_dft_plan_set_norm(&p, true);
  srsran_dft_run_c(&p, in, out);
  double num = 0, den = 0, sgn = backward ? 1.0 : -1.0;
  for (int k = 0; k < N; k += (N > 64 ? 37 : 1)) {
    double _Complex acc = 0;
    for (int n = 0; n < N; n++) acc += (double _Complex)in[n] * cexp(sgn * I * 2 * M_PI * (double)k * n / N);
    acc /= sqrt((double)N);
    num += pow(cabs(acc - out[k]), 2);
    den += pow(cabs(acc), 2);
  }
  CHECK(sqrt(num / den) < 1e-4, "dft N=%d backward=%d rel err %g", N, backward, sqrt(num / den));
  srsran_dft_plan_free(&p);
  free(in);
  free(out);
}

static void test_dft_precoding(uint32_t nof_prb)
{
  srsran_dft_precoding_t tx, rx;
  const int              N = 12 * (int)nof_prb, nsym = 12;
  cf_t*                  in  = malloc(sizeof(cf_t) * N * nsym);
  cf_t*                  mid = malloc(sizeof(cf_t) * N * nsym);
  cf_t*                  out = malloc(sizeof(cf_t) * N * nsym);
  for (int i = 0; i < N * nsym; i++) in[i] = gauss() + I * gauss();
  CHECK(srsran_dft_precoding_init_tx(&tx, nof_prb) == 0 && srsran_dft_precoding_init_rx(&rx, nof_prb) == 0, "dft_precoding init");
  CHECK(srsran_dft_precoding(&tx, in, mid, nof_prb, nsym) == 0, "dft_precoding tx");
  CHECK(srsran_dft_precoding(&rx, mid, out, nof_prb, nsym) == 0, "dft_precoding rx");
  /* forward transform against the oracle's restatement, and the round trip */
  float _Complex* want = malloc(sizeof(float _Complex) * N * nsym);
  orc_dft_precoding((const float _Complex*)in, want, nof_prb, nsym, 1);
  double num = 0, den = 0, num2 = 0;
  for (int i = 0; i < N * nsym; i++) {
    num += pow(cabsf(mid[i] - want[i]), 2);
    den += pow(cabsf(want[i]), 2);
    num2 += pow(cabsf(out[i] - in[i]), 2);
  }
  CHECK(sqrt(num / den) < 1e-4 && sqrt(num2 / den) < 1e-4, "dft_precoding prb=%u rel err %g round trip %g", nof_prb, sqrt(num / den),
        sqrt(num2 / den));
  CHECK(srsran_dft_precoding(&rx, mid, out, 7, nsym) != 0, "dft_precoding must refuse 7 PRB");
  srsran_dft_precoding_free(&tx);
  srsran_dft_precoding_free(&rx);
  free(in);
  free(mid);
  free(out);
  free(want);
}

int main(void)
{
  srand(1234);
  uint32_t s[9];
  srsran_cbsegm_t seg;
  for (uint32_t tbs = 16; tbs < 80000; tbs += 1237) {
    orc_cbsegm(tbs, s);
    srsran_cbsegm(&seg, tbs);
    CHECK(seg.F == s[0] && seg.C == s[1] && seg.K1 == s[2] && seg.K2 == s[3] && seg.C1 == s[6] && seg.C2 == s[7], "cbsegm %u", tbs);
  }
  srsran_crc_t crc;
  uint8_t      msg[1024];
  for (int i = 0; i < 1024; i++) msg[i] = rand() & 0xFF;
  srsran_crc_init(&crc, 0x1864CFB, 24);
  CHECK(srsran_crc_checksum_byte(&crc, msg, 8192) == orc_crc24(0, msg, 8192), "crc24a");
  srsran_crc_init(&crc, 0x1800063, 24);
  CHECK(srsran_crc_checksum_byte(&crc, msg, 6144) == orc_crc24(1, msg, 6144), "crc24b");

  test_tdec(40, 0.8f);
  test_tdec(504, 0.9f);
  test_tdec(6144, 0.93f);
  srsran_rm_turbo_gentables();
  for (uint32_t rv = 0; rv < 4; rv++) {
    test_rm(0, rv, 0.6f);
    test_rm(60, rv, 1.0f);
    test_rm(187, rv, 1.9f);
  }
  srsran_rm_turbo_free_tables();
  test_ofdm(6, 0, 0, 0);
  test_ofdm(25, 0, -0.5f, 0.5f);
  test_ofdm(100, 2048, -0.5f, 0.5f);
  srsran_use_standard_symbol_size(true);
  CHECK(srsran_symbol_sz(100) == 2048 && srsran_symbol_sz(25) == 512, "standard sizes");
  srsran_use_standard_symbol_size(false);
  CHECK(srsran_symbol_sz(100) == 1536 && srsran_symbol_sz(25) == 384, "default sizes");
  test_dft(12, 0);
  test_dft(128, 0);
  test_dft(1536, 1);
  test_dft(2048, 0);
  test_dft(1200, 1);
  test_dft(60, 0);
  CHECK(srsran_dft_precoding_valid_prb(100) && srsran_dft_precoding_valid_prb(81) && !srsran_dft_precoding_valid_prb(7) &&
            !srsran_dft_precoding_valid_prb(101) && srsran_dft_precoding_get_valid_prb(99) == 96,
        "valid prb");
  test_dft_precoding(6);
  test_dft_precoding(25);
  test_dft_precoding(100);
  printf(fails ? "compat_test: %d FAILURES\n" : "compat_test: all checks passed\n", fails);
  return fails ? 1 : 0;
}
