/* Per-call latency of the reference-named API of libsrslte_b200.so: what a caller that keeps the reference's per-block /
 * per-subframe calling pattern pays (one call = host->device copy, kernels, device->host copy, synchronise).
 *   gcc -O2 -Iinclude tools/compat_latency.c -o tools/compat_latency -Lsrslte_b200 -lsrslte_b200 -lm -Wl,-rpath,$PWD/srslte_b200 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "srslte_b200_srsran_api.h"

static double now_us(void)
{
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec * 1e6 + t.tv_nsec * 1e-3;
}

int main(void)
{
  const uint32_t K = 6144;
  int16_t*       llr = malloc(sizeof(int16_t) * (3 * K + 12));
  uint8_t        out[768];
  srand(1);
  for (uint32_t i = 0; i < 3 * K + 12; i++) llr[i] = (int16_t)(rand() % 63 - 31);
  srsran_tdec_t h;
  if (srsran_tdec_init(&h, K) != SRSRAN_SUCCESS) return 1;
  srsran_tdec_run_all(&h, llr, out, 8, K);
  const int reps = 50;
  double    t0   = now_us();
  for (int r = 0; r < reps; r++) srsran_tdec_run_all(&h, llr, out, 8, K);
  printf("srsran_tdec_run_all      K=6144, 8 iterations : %8.1f us per call\n", (now_us() - t0) / reps);
  srsran_tdec_new_cb(&h, K);
  srsran_tdec_iteration(&h, llr, out);
  t0 = now_us();
  for (int r = 0; r < reps; r++) srsran_tdec_iteration(&h, llr, out);
  printf("srsran_tdec_iteration    K=6144, one iteration + decision bytes : %8.1f us per call\n", (now_us() - t0) / reps);
  srsran_tdec_free(&h);

  srsran_rm_turbo_gentables();
  int16_t* e    = malloc(sizeof(int16_t) * 20000);
  int16_t* soft = calloc(3 * K + 12, sizeof(int16_t));
  for (int i = 0; i < 20000; i++) e[i] = (int16_t)(rand() % 63 - 31);
  srsran_rm_turbo_rx_lut_(e, soft, 20000, 187, 0, false);
  t0 = now_us();
  for (int r = 0; r < reps; r++) srsran_rm_turbo_rx_lut_(e, soft, 20000, 187, 0, false);
  printf("srsran_rm_turbo_rx_lut_  K=6144, E=20000 : %8.1f us per call\n", (now_us() - t0) / reps);

  const uint32_t N   = 2048;
  cf_t*          in  = calloc(15 * N, sizeof(cf_t));
  cf_t*          grd = calloc(14 * 1200, sizeof(cf_t));
  srsran_ofdm_t  q;
  if (srsran_ofdm_rx_init(&q, SRSRAN_CP_NORM, in, grd, 100) != SRSRAN_SUCCESS) return 2;
  srsran_ofdm_rx_sf(&q);
  t0 = now_us();
  for (int r = 0; r < reps; r++) srsran_ofdm_rx_sf(&q);
  printf("srsran_ofdm_rx_sf        100 PRB, N=2048 : %8.1f us per call\n", (now_us() - t0) / reps);
  srsran_ofdm_rx_free(&q);
  return 0;
}
