// The reference's per-object C API (include/srslte_b200_srsran_api.h) on top of the GPU engines.
// These entries exist so that existing callers and the reference's own unit tests link and behave the same; the
// throughput path is the batched API of include/srslte_b200.h.  One code block / one subframe per call means one or two
// kernel launches plus a synchronous copy per call -- correct, not fast.
#include <math.h>
#include <string.h>

#include <mutex>
#include <new>
#include <vector>

#include "../../include/srslte_b200_srsran_api.h"
#include "b200_runtime.h"
#include "lte_tables.h"
#include "ofdm_kernels.h"
#include "tdec_engine.h"
#include "tdec_kernels.h"

using namespace b200;

static int compat_device()
{
  const char* e = getenv("SRSLTE_B200_DEVICE"); // which GPU the per-object API binds to (the batched API takes it explicitly)
  return e ? atoi(e) : 0;
}

// ===================================================================================================================
// code block tables, CRC, interleaver tables (host-side helpers of the decode loop)
extern "C" {

int srsran_cbsegm(srsran_cbsegm_t* s, uint32_t tbs)
{
  if (!s) return SRSRAN_ERROR_INVALID_INPUTS;
  CbSegm c;
  memset(s, 0, sizeof(*s));
  if (tbs == 0) return SRSRAN_SUCCESS; // cbsegm.c:67-69
  if (cb_segmentation(tbs, c) != 0) return SRSRAN_ERROR;
  s->F = c.F; s->C = c.C; s->K1 = c.K1; s->K2 = c.K2; s->K1_idx = c.K1_idx; s->K2_idx = c.K2_idx;
  s->C1 = c.C1; s->C2 = c.C2; s->tbs = c.tbs;
  s->L_tb = 24; // cbsegm.c:98-99
  s->L_cb = 24;
  return SRSRAN_SUCCESS;
}

int srsran_cbsegm_cbsize(uint32_t index)
{
  return cb_size(index) > 0 ? cb_size(index) : SRSRAN_ERROR;
}

bool srsran_cbsegm_cbsize_isvalid(uint32_t size)
{
  return cb_index_exact(size) >= 0;
}

int srsran_cbsegm_cbindex(uint32_t long_cb)
{
  int i = cb_index(long_cb);
  return i < 0 ? SRSRAN_ERROR : i;
}

int srsran_crc_init(srsran_crc_t* h, uint32_t crc_poly, int crc_order)
{
  if (!h || crc_order < 8 || crc_order > 32) return SRSRAN_ERROR;
  h->polynom    = (int)crc_poly;
  h->order      = crc_order;
  h->crcinit    = 0;
  h->crcmask    = ((((uint64_t)1 << (crc_order - 1)) - 1) << 1) | 1;
  h->crchighbit = (uint64_t)1 << (crc_order - 1);
  for (uint32_t i = 0; i < 256; i++) { // byte-at-a-time table, MSB first (crc.c:30-46)
    uint64_t r = (uint64_t)i << (crc_order - 8);
    for (int b = 0; b < 8; b++) r = (r & h->crchighbit) ? ((r << 1) ^ crc_poly) : (r << 1);
    h->table[i] = r & h->crcmask;
  }
  return SRSRAN_SUCCESS;
}

uint32_t srsran_crc_checksum_byte(srsran_crc_t* h, const uint8_t* data, int len)
{
  uint64_t       crc = 0;
  const uint32_t sh  = (uint32_t)h->order - 8u;
  for (int i = 0; i < len / 8; i++) crc = ((crc << 8) ^ h->table[((crc >> sh) & 0xFF) ^ data[i]]) & h->crcmask;
  h->crcinit = crc;
  return (uint32_t)crc;
}

int srsran_tc_interl_init(srsran_tc_interl_t* h, uint32_t max_long_cb)
{
  if (!h) return SRSRAN_ERROR;
  h->max_long_cb = max_long_cb;
  h->forward     = (uint16_t*)malloc(sizeof(uint16_t) * max_long_cb);
  h->reverse     = (uint16_t*)malloc(sizeof(uint16_t) * max_long_cb);
  if (!h->forward || !h->reverse) {
    perror("malloc");
    srsran_tc_interl_free(h);
    return SRSRAN_ERROR;
  }
  return SRSRAN_SUCCESS;
}

void srsran_tc_interl_free(srsran_tc_interl_t* h)
{
  if (!h) return;
  free(h->forward);
  free(h->reverse);
  memset(h, 0, sizeof(*h));
}

int srsran_tc_interl_LTE_gen_interl(srsran_tc_interl_t* h, uint32_t long_cb, uint32_t interl_win)
{
  if (long_cb > h->max_long_cb) {
    B200_LOG_ERROR("Interleaver initiated for max_long_cb=%u", h->max_long_cb);
    return SRSRAN_ERROR;
  }
  const int idx = cb_index_exact(long_cb);
  if (idx < 0) {
    B200_LOG_ERROR("Can't find long_cb=%u in valid TC CB table", long_cb);
    return SRSRAN_ERROR;
  }
  std::vector<uint16_t> f, r;
  qpp_tables(idx, f, r);
  if (interl_win <= 1) {
    memcpy(h->forward, f.data(), sizeof(uint16_t) * long_cb);
    memcpy(h->reverse, r.data(), sizeof(uint16_t) * long_cb);
    return SRSRAN_SUCCESS;
  }
  // sub-block re-indexing of tc_interl_lte.c:95-105 (only the CPU SIMD decoders consume it)
  const uint32_t W = interl_win, L = long_cb / W;
  for (uint32_t i = 0; i < long_cb; i++) {
    const uint32_t in = (i % W) * L + i / W;
    const uint32_t ff = f[in], rr = r[in];
    h->forward[i]     = (uint16_t)((ff % L) * W + ff / L);
    h->reverse[i]     = (uint16_t)((rr % L) * W + rr / L);
  }
  return SRSRAN_SUCCESS;
}

int srsran_tc_interl_LTE_gen(srsran_tc_interl_t* h, uint32_t long_cb)
{
  return srsran_tc_interl_LTE_gen_interl(h, long_cb, 1);
}

} // extern "C"

// ===================================================================================================================
// rate de-matching through a process-wide engine (the reference's tables are process-wide statics too, rm_turbo.c:79-81)
static std::mutex          g_rm_mutex;
static srsran_b200_sch_t*  g_rm_sch = nullptr;

extern "C" {

void srsran_rm_turbo_gentables(void)
{
  std::lock_guard<std::mutex> lk(g_rm_mutex);
  if (!g_rm_sch) {
    if (srsran_b200_sch_init(&g_rm_sch, compat_device()) != SRSRAN_SUCCESS) g_rm_sch = nullptr;
  }
}

void srsran_rm_turbo_free_tables(void)
{
  std::lock_guard<std::mutex> lk(g_rm_mutex);
  if (g_rm_sch) srsran_b200_sch_free(g_rm_sch);
  g_rm_sch = nullptr;
}

int srsran_rm_turbo_rx_lut_(int16_t* input, int16_t* output, uint32_t in_len, uint32_t cb_idx, uint32_t rv_idx, bool enable_input_tdec)
{
  (void)enable_input_tdec; // the GPU decoder takes the natural layout, so both variants coincide
  if (rv_idx >= 4 || cb_idx >= (uint32_t)NOF_CB_SIZES) {
    printf("Invalid inputs rv_idx=%d, cb_idx=%d\n", rv_idx, cb_idx); // rm_turbo.c:442-444
    return SRSRAN_ERROR_INVALID_INPUTS;
  }
  srsran_rm_turbo_gentables();
  std::lock_guard<std::mutex> lk(g_rm_mutex);
  if (!g_rm_sch) return SRSRAN_ERROR;
  srsran_b200_rm_cb_t j;
  j.cb_idx      = cb_idx;
  j.rv          = rv_idx;
  j.E           = in_len;
  j.new_data    = 0;
  j.in_offset   = 0;
  j.soft_offset = 0;
  return srsran_b200_rm_turbo_rx_batch(g_rm_sch, input, in_len, output, 3 * (uint32_t)cb_size(cb_idx) + 12, &j, 1, 0, nullptr);
}

int srsran_rm_turbo_rx_lut(int16_t* input, int16_t* output, uint32_t in_len, uint32_t cb_idx, uint32_t rv_idx)
{
  return srsran_rm_turbo_rx_lut_(input, output, in_len, cb_idx, rv_idx, true);
}

// The reference's 8-bit soft-bit container (rm_turbo.c:447-483): output[deinter[i % out_len]] += input[i] in int8, which wraps
// modulo 256.  Computed here with the int16 kernel on widened values and narrowed again: the low byte of the wrapped 16-bit
// sum IS the wrapped 8-bit sum, so the result equals the reference's scalar form (natural layout, like the 16-bit entry).
int srsran_rm_turbo_rx_lut_8bit(int8_t* input, int8_t* output, uint32_t in_len, uint32_t cb_idx, uint32_t rv_idx)
{
  if (rv_idx >= 4 || cb_idx >= (uint32_t)NOF_CB_SIZES) {
    printf("Invalid inputs rv_idx=%d, cb_idx=%d\n", rv_idx, cb_idx);
    return SRSRAN_ERROR_INVALID_INPUTS;
  }
  if (!input || !output) return SRSRAN_ERROR_INVALID_INPUTS;
  const uint32_t       n = 3 * (uint32_t)cb_size(cb_idx) + 12;
  std::vector<int16_t> in16(in_len), out16(n);
  for (uint32_t i = 0; i < in_len; i++) in16[i] = (int16_t)input[i];
  for (uint32_t i = 0; i < n; i++) out16[i] = (int16_t)output[i];
  const int rc = srsran_rm_turbo_rx_lut_(in16.data(), out16.data(), in_len, cb_idx, rv_idx, false);
  if (rc != SRSRAN_SUCCESS) return rc;
  for (uint32_t i = 0; i < n; i++) output[i] = (int8_t)(uint8_t)(uint16_t)out16[i];
  return SRSRAN_SUCCESS;
}

} // extern "C"

// ===================================================================================================================
// srsran_tdec_t: one code block at a time on the batched engine (a batch of one)
struct CompatTdec {
  TdecEngine eng;
  TdecView   view{};
  int        K      = 0;
  int        cb_idx = -1;
  uint8_t *  d_out = nullptr;
  int16_t*   d_llr = nullptr;
  std::vector<int16_t> wide; // 8-bit entries: the input widened to int16
};

static CompatTdec* tdec_of(srsran_tdec_t* h)
{
  return h ? (CompatTdec*)h->dec16_hdlr[0] : nullptr;
}

extern "C" {

int srsran_tdec_init_manual(srsran_tdec_t* h, uint32_t max_long_cb, srsran_tdec_impl_type_t dec_type)
{
  if (!h) return SRSRAN_ERROR_INVALID_INPUTS;
  memset(h, 0, sizeof(*h)); // turbodecoder.c:154
  CompatTdec* c = new (std::nothrow) CompatTdec();
  if (!c) return SRSRAN_ERROR;
  if (c->eng.init(compat_device(), 0) != B200_SUCCESS || cudaMalloc(&c->d_out, MAX_CB_LEN / 8) != cudaSuccess ||
      cudaMalloc(&c->d_llr, (3 * MAX_CB_LEN + 12) * sizeof(int16_t)) != cudaSuccess ||
      c->eng.ws.arena.reserve(TdecEngine::workspace_bytes(MAX_CB_LEN, 1)) != B200_SUCCESS) {
    c->eng.destroy();
    delete c;
    return SRSRAN_ERROR;
  }
  h->dec16_hdlr[0]    = c;
  h->max_long_cb      = max_long_cb;
  h->dec_type         = dec_type;
  h->current_llr_type = SRSRAN_TDEC_16;
  h->current_cbidx    = -1;
  return SRSRAN_SUCCESS;
}

int srsran_tdec_init(srsran_tdec_t* h, uint32_t max_long_cb)
{
  return srsran_tdec_init_manual(h, max_long_cb, SRSRAN_TDEC_AUTO);
}

void srsran_tdec_free(srsran_tdec_t* h)
{
  if (!h) return;
  CompatTdec* c = tdec_of(h);
  if (c) {
    if (c->d_out) cudaFree(c->d_out);
    if (c->d_llr) cudaFree(c->d_llr);
    c->eng.destroy();
    delete c;
  }
  memset(h, 0, sizeof(*h)); // turbodecoder.c:362
}

void srsran_tdec_force_not_sb(srsran_tdec_t* h)
{
  if (h) h->force_not_sb = true;
}

uint32_t srsran_tdec_autoimp_get_subblocks(uint32_t)
{
  return 0; // natural layout for every size (see the header)
}

uint32_t srsran_tdec_autoimp_get_subblocks_8bit(uint32_t)
{
  return 0;
}

int srsran_tdec_new_cb(srsran_tdec_t* h, uint32_t long_cb)
{
  if (!h || !tdec_of(h)) return -1;
  if (long_cb > h->max_long_cb) {
    B200_LOG_ERROR("TDEC was initialized for max_long_cb=%u", h->max_long_cb); // turbodecoder.c:512-515
    return -1;
  }
  h->n_iter          = 0;
  h->current_long_cb = long_cb;
  h->current_cbidx   = cb_index_exact(long_cb);
  if (h->current_cbidx < 0) {
    B200_LOG_ERROR("Invalid CB length %u", long_cb);
    return -1;
  }
  return 0;
}

int srsran_tdec_get_nof_iterations(srsran_tdec_t* h)
{
  return h ? h->n_iter : 0;
}

static int tdec_one_pass(srsran_tdec_t* h, int16_t* input)
{
  CompatTdec*  c  = tdec_of(h);
  const int    K  = (int)h->current_long_cb;
  cudaStream_t st = c->eng.pipe_stream[0];
  if (cudaSetDevice(c->eng.ctx->device) != cudaSuccess) return SRSRAN_ERROR;
  if (h->n_iter == 0) { // first pass reads the input (turbodecoder_iter.h:99-101)
    std::vector<TdecGroupSpec> one(1);
    one[0] = TdecGroupSpec{K, h->current_cbidx, SRSRAN_B200_CRC_NONE, 1, 0, 0, 0};
    if (c->eng.prepare(c->eng.ws, one, st) != B200_SUCCESS) return SRSRAN_ERROR;
    c->view               = c->eng.ws.plan.v;
    c->view.early_stop    = 0;
    c->view.max_pass      = 1 << 30;
    c->view.split_percent = 51;
    c->K                  = K;
    c->cb_idx             = h->current_cbidx;
    B200_CUDA_TRY(cudaMemcpyAsync(c->d_llr, input, (3 * (size_t)K + 12) * sizeof(int16_t), cudaMemcpyHostToDevice, st));
    launch_load_natural(c->view, K, c->d_llr, nullptr, true, st);
    g_kernel_launches++;
  }
  launch_siso_pass(c->view, h->n_iter, SISO_LOW_LATENCY, c->eng.sm_count, st); // one code block: latency is all that matters
  g_kernel_launches++;
  h->n_iter++;
  return SRSRAN_SUCCESS;
}

static int tdec_decide(srsran_tdec_t* h, uint8_t* output)
{
  CompatTdec*  c  = tdec_of(h);
  cudaStream_t st = c->eng.pipe_stream[0];
  launch_decide(c->view, c->K, c->d_out, nullptr, nullptr, nullptr, st);
  g_kernel_launches++;
  B200_CUDA_TRY(cudaMemcpyAsync(output, c->d_out, h->current_long_cb / 8, cudaMemcpyDeviceToHost, st));
  B200_CUDA_TRY(cudaStreamSynchronize(st));
  B200_CUDA_TRY(cudaGetLastError());
  return SRSRAN_SUCCESS;
}

void srsran_tdec_iteration(srsran_tdec_t* h, int16_t* input, uint8_t* output)
{
  if (!h || !tdec_of(h) || h->current_cbidx < 0) return; // turbodecoder.c:529
  if (tdec_one_pass(h, input) == SRSRAN_SUCCESS) tdec_decide(h, output);
}

int srsran_tdec_run_all(srsran_tdec_t* h, int16_t* input, uint8_t* output, uint32_t nof_iterations, uint32_t long_cb)
{
  if (srsran_tdec_new_cb(h, long_cb)) return SRSRAN_ERROR;
  do {
    if (tdec_one_pass(h, input) != SRSRAN_SUCCESS) return SRSRAN_ERROR;
  } while (h->n_iter < (int)nof_iterations); // at least one pass, like turbodecoder.c:542-544
  return tdec_decide(h, output);
}

// 8-bit LLR input (turbodecoder.c:410-484,551-577).  The reference sends int8 inputs either to its saturating 8-bit window
// decoders or -- for every length those cannot take -- through convert_8_to_16 into a 16-bit decoder (turbodecoder.c:441-470).
// This library always takes the second route: the values are widened and decoded with the generic int16 arithmetic, i.e. the
// result is bit-exact with srsran_tdec_run_all on the widened values (and NOT with the reference's approximate 8-bit window
// decoders, which differ from its own 16-bit decoders as well; tests/test_compat_8bit_gpu.py compares block error rates).
static void widen8(CompatTdec* c, const int8_t* in, uint32_t n)
{
  c->wide.resize(n);
  for (uint32_t i = 0; i < n; i++) c->wide[i] = (int16_t)in[i];
}

void srsran_tdec_iteration_8bit(srsran_tdec_t* h, int8_t* input, uint8_t* output)
{
  if (!h || !tdec_of(h) || h->current_cbidx < 0 || !input) return; // turbodecoder.c:553
  CompatTdec* c = tdec_of(h);
  if (h->n_iter == 0) widen8(c, input, 3 * h->current_long_cb + 12); // only the first iteration reads the input (:466)
  if (tdec_one_pass(h, c->wide.data()) == SRSRAN_SUCCESS) tdec_decide(h, output);
}

int srsran_tdec_run_all_8bit(srsran_tdec_t* h, int8_t* input, uint8_t* output, uint32_t nof_iterations, uint32_t long_cb)
{
  if (!input || srsran_tdec_new_cb(h, long_cb)) return SRSRAN_ERROR;
  CompatTdec* c = tdec_of(h);
  widen8(c, input, 3 * long_cb + 12);
  do {
    if (tdec_one_pass(h, c->wide.data()) != SRSRAN_SUCCESS) return SRSRAN_ERROR;
  } while (h->n_iter < (int)nof_iterations);
  return tdec_decide(h, output);
}

} // extern "C"

// ===================================================================================================================
// srsran_dft_plan_t: generic complex DFT of size 2^a 3^b 5^c on the OFDM kernel's FFT core
struct CompatDft {
  DeviceContext* ctx = nullptr;
  OfdmPlanDev    plan{};
  float2*        dW = nullptr;
  float2 *       d_in = nullptr, *d_out = nullptr;
  size_t         cap = 0;
  int            sm_count = 148;
  cudaStream_t   stream = nullptr;
  // guru geometry
  cf_t *gin = nullptr, *gout = nullptr;
  int   how_many = 1, idist = 0, odist = 0;
};

static void dft_destroy(CompatDft* d)
{
  if (!d) return;
  if (d->dW) cudaFree(d->dW);
  if (d->d_in) cudaFree(d->d_in);
  if (d->d_out) cudaFree(d->d_out);
  if (d->stream) cudaStreamDestroy(d->stream);
  delete d;
}

static int dft_setup(CompatDft* d, int N, bool forward)
{
  int       radix[OFDM_MAX_PASSES];
  const int npass = N >= 4 ? fft_factorise(N, radix) : 0;
  if (npass == 0) {
    B200_LOG_ERROR("DFT size %d not supported on the GPU path (2^a 3^b 5^c, >= 4)", N);
    return SRSRAN_ERROR;
  }
  OfdmPlanDev& p = d->plan;
  p              = OfdmPlanDev{};
  p.N            = N;
  p.R            = N;
  p.nsym         = 1;
  p.generic      = 1;
  p.inverse      = forward ? 0 : 1;
  p.npass        = npass;
  for (int i = 0; i < OFDM_MAX_PASSES; i++) p.radix[i] = radix[i];
  int tps = N / 16;
  if (tps < 8) tps = 8;
  if (tps > OFDM_THREADS) tps = OFDM_THREADS;
  p.tps = tps;
  if (d->dW) cudaFree(d->dW);
  std::vector<float2> W(N);
  for (int m = 0; m < N; m++) {
    double a = -2.0 * M_PI * (double)m / (double)N;
    W[m]     = make_float2((float)cos(a), (float)sin(a));
  }
  B200_CUDA_TRY(cudaMalloc(&d->dW, N * sizeof(float2)));
  B200_CUDA_TRY(cudaMemcpy(d->dW, W.data(), N * sizeof(float2), cudaMemcpyHostToDevice));
  p.W = d->dW;
  return SRSRAN_SUCCESS;
}

static int dft_exec(CompatDft* d, const cf_t* in, cf_t* out, int how_many, int idist, int odist)
{
  const int    N     = d->plan.N;
  const size_t in_n  = (size_t)(how_many - 1) * idist + N, out_n = (size_t)(how_many - 1) * odist + N;
  B200_CUDA_TRY(cudaSetDevice(d->ctx->device));
  if ((in_n + out_n) * sizeof(float2) > d->cap) {
    if (d->d_in) cudaFree(d->d_in);
    if (d->d_out) cudaFree(d->d_out);
    B200_CUDA_TRY(cudaMalloc(&d->d_in, in_n * sizeof(float2)));
    B200_CUDA_TRY(cudaMalloc(&d->d_out, out_n * sizeof(float2)));
    d->cap = (in_n + out_n) * sizeof(float2);
  }
  d->plan.idist = idist;
  d->plan.odist = odist;
  B200_CUDA_TRY(cudaMemcpyAsync(d->d_in, in, in_n * sizeof(float2), cudaMemcpyHostToDevice, d->stream));
  if (odist != N) B200_CUDA_TRY(cudaMemcpyAsync(d->d_out, out, out_n * sizeof(float2), cudaMemcpyHostToDevice, d->stream));
  if (launch_ofdm_rx(d->plan, d->d_in, d->d_out, (uint32_t)how_many, d->sm_count, d->stream) != B200_SUCCESS) return SRSRAN_ERROR;
  g_kernel_launches++;
  B200_CUDA_TRY(cudaMemcpyAsync(out, d->d_out, out_n * sizeof(float2), cudaMemcpyDeviceToHost, d->stream));
  B200_CUDA_TRY(cudaStreamSynchronize(d->stream));
  return SRSRAN_SUCCESS;
}

static int dft_new(srsran_dft_plan_t* plan, int N, srsran_dft_dir_t dir, bool guru)
{
  memset(plan, 0, sizeof(*plan));
  CompatDft* d = new (std::nothrow) CompatDft();
  if (!d) return SRSRAN_ERROR;
  d->ctx = device_context(compat_device());
  if (!d->ctx || cudaSetDevice(d->ctx->device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) != cudaSuccess ||
      dft_setup(d, N, dir == SRSRAN_DFT_FORWARD) != SRSRAN_SUCCESS) {
    dft_destroy(d);
    return SRSRAN_ERROR;
  }
  cudaDeviceGetAttribute(&d->sm_count, cudaDevAttrMultiProcessorCount, d->ctx->device);
  plan->p         = d;
  plan->size      = N;
  plan->init_size = N;
  plan->mode      = SRSRAN_DFT_COMPLEX;
  plan->dir       = dir;
  plan->forward   = dir == SRSRAN_DFT_FORWARD;
  plan->is_guru   = guru;
  return SRSRAN_SUCCESS;
}

extern "C" {

int srsran_dft_plan_c(srsran_dft_plan_t* plan, int dft_points, srsran_dft_dir_t dir)
{
  if (!plan) return SRSRAN_ERROR;
  if (dft_new(plan, dft_points, dir, false) != SRSRAN_SUCCESS) return SRSRAN_ERROR;
  plan->in  = malloc(sizeof(cf_t) * (size_t)dft_points); // plan-owned buffers (dft_fftw.c:106-111)
  plan->out = malloc(sizeof(cf_t) * (size_t)dft_points);
  return SRSRAN_SUCCESS;
}

int srsran_dft_plan_guru_c(srsran_dft_plan_t* plan, int dft_points, srsran_dft_dir_t dir, cf_t* in_buffer, cf_t* out_buffer,
                           int istride, int ostride, int how_many, int idist, int odist)
{
  if (!plan) return SRSRAN_ERROR;
  if (istride != 1 || ostride != 1) {
    B200_LOG_ERROR("guru DFT plans with element strides other than 1 are not supported");
    return SRSRAN_ERROR;
  }
  if (dft_new(plan, dft_points, dir, true) != SRSRAN_SUCCESS) return SRSRAN_ERROR;
  CompatDft* d = (CompatDft*)plan->p;
  d->gin       = in_buffer;
  d->gout      = out_buffer;
  d->how_many  = how_many;
  d->idist     = idist;
  d->odist     = odist;
  return SRSRAN_SUCCESS;
}

int srsran_dft_replan_c(srsran_dft_plan_t* plan, int new_dft_points)
{
  if (!plan || !plan->p) return SRSRAN_ERROR;
  if (dft_setup((CompatDft*)plan->p, new_dft_points, plan->forward) != SRSRAN_SUCCESS) return SRSRAN_ERROR;
  plan->size = new_dft_points;
  return SRSRAN_SUCCESS;
}

int srsran_dft_replan(srsran_dft_plan_t* plan, const int new_dft_points)
{
  if (!plan) return SRSRAN_ERROR;
  if (new_dft_points > plan->init_size) { // dft_fftw.c:92-104
    B200_LOG_ERROR("DFT: Error calling replan: new_dft_points (%d) must be lower or equal dft_size passed initially (%d)",
                   new_dft_points, plan->init_size);
    return -1;
  }
  return srsran_dft_replan_c(plan, new_dft_points);
}

void srsran_dft_plan_free(srsran_dft_plan_t* plan)
{
  if (!plan || !plan->size) return;
  if (!plan->is_guru) {
    free(plan->in);
    free(plan->out);
  }
  dft_destroy((CompatDft*)plan->p);
  memset(plan, 0, sizeof(*plan));
}

void srsran_dft_plan_set_mirror(srsran_dft_plan_t* plan, bool val) { plan->mirror = val; }
void srsran_dft_plan_set_db(srsran_dft_plan_t* plan, bool val) { plan->db = val; }
void srsran_dft_plan_set_norm(srsran_dft_plan_t* plan, bool val) { plan->norm = val; }
void srsran_dft_plan_set_dc(srsran_dft_plan_t* plan, bool val) { plan->dc = val; }

void srsran_dft_run_c_zerocopy(srsran_dft_plan_t* plan, const cf_t* in, cf_t* out)
{
  if (!plan || !plan->p) return;
  dft_exec((CompatDft*)plan->p, in, out, 1, plan->size, plan->size);
}

// copy_pre / copy_post of dft_fftw.c:297-320, then execute, norm, dB
void srsran_dft_run_c(srsran_dft_plan_t* plan, const cf_t* in, cf_t* out)
{
  if (!plan || !plan->p || plan->is_guru) return;
  const int N = plan->size, off = plan->dc ? 1 : 0;
  cf_t *    pi = (cf_t*)plan->in, *po = (cf_t*)plan->out;
  if (plan->mirror && !plan->forward) {
    const int hlen = N / 2;
    memset((void*)pi, 0, sizeof(cf_t) * off);
    memcpy(&pi[off], &in[hlen], sizeof(cf_t) * (N - hlen - off));
    memcpy(&pi[N - hlen], in, sizeof(cf_t) * hlen);
  } else {
    memcpy(pi, in, sizeof(cf_t) * N);
  }
  if (dft_exec((CompatDft*)plan->p, pi, po, 1, N, N) != SRSRAN_SUCCESS) return;
  float* f = (float*)po;
  if (plan->norm) {
    const float norm = 1.0f / sqrtf((float)N);
    for (int i = 0; i < 2 * N; i++) f[i] *= norm;
  }
  if (plan->db) {
    for (int i = 0; i < N; i++) {
      // dft_fftw.c:346-349 hands a COMPLEX value to srsran_convert_power_to_dB(float): the C conversion keeps the real part
      // only, so the reference's result is 10 log10(Re x) (NaN for a negative real part), imaginary part zero -- reproduced
      f[2 * i]     = 10.0f * log10f(f[2 * i]);
      f[2 * i + 1] = 0.0f;
    }
  }
  if (plan->mirror && plan->forward) {
    const int hlen = (N - 1) / 2 + 1;
    memcpy(out, &po[hlen], sizeof(cf_t) * (N - hlen));
    memcpy(&out[N - hlen], &po[off], sizeof(cf_t) * (hlen - off));
  } else {
    memcpy(out, po, sizeof(cf_t) * N);
  }
}

void srsran_dft_run_guru_c(srsran_dft_plan_t* plan)
{
  if (!plan || !plan->p) return;
  if (!plan->is_guru) {
    B200_LOG_ERROR("srsran_dft_run_guru_c: the selected plan is not guru!"); // dft_fftw.c:361
    return;
  }
  CompatDft* d = (CompatDft*)plan->p;
  dft_exec(d, d->gin, d->gout, d->how_many, d->idist, d->odist);
}


// ---- srsran_dft_precoding_t (dft_precoding.c:39-126) over the plans above -------------------------------------------------
bool srsran_dft_precoding_valid_prb(uint32_t nof_prb)
{
  if (nof_prb == 0 || nof_prb > 100) return nof_prb == 0; // the reference's table marks 0 valid and stops at 100 (:88-104)
  uint32_t n = nof_prb;
  for (uint32_t f : {2u, 3u, 5u}) {
    while (n % f == 0) n /= f;
  }
  return n == 1;
}

uint32_t srsran_dft_precoding_get_valid_prb(uint32_t nof_prb)
{
  while (!srsran_dft_precoding_valid_prb(nof_prb)) nof_prb--;
  return nof_prb;
}

void srsran_dft_precoding_free(srsran_dft_precoding_t* q)
{
  if (!q) return;
  for (uint32_t i = 1; i <= q->max_prb && i <= SRSRAN_MAX_PRB; i++) {
    if (q->dft_plan[i].p) srsran_dft_plan_free(&q->dft_plan[i]);
  }
  memset(q, 0, sizeof(*q));
}

int srsran_dft_precoding_init(srsran_dft_precoding_t* q, uint32_t max_prb, bool is_tx)
{
  if (!q || max_prb > SRSRAN_MAX_PRB) return SRSRAN_ERROR_INVALID_INPUTS;
  memset(q, 0, sizeof(*q));
  q->max_prb = max_prb;
  for (uint32_t i = 1; i <= max_prb; i++) {
    if (!srsran_dft_precoding_valid_prb(i)) continue;
    if (srsran_dft_plan_c(&q->dft_plan[i], (int)(12 * i), is_tx ? SRSRAN_DFT_FORWARD : SRSRAN_DFT_BACKWARD)) {
      B200_LOG_ERROR("Error: Creating DFT plan %u", i); // dft_precoding.c:51
      srsran_dft_precoding_free(q);
      return SRSRAN_ERROR;
    }
    srsran_dft_plan_set_norm(&q->dft_plan[i], true);
  }
  return SRSRAN_SUCCESS;
}

int srsran_dft_precoding_init_rx(srsran_dft_precoding_t* q, uint32_t max_prb) { return srsran_dft_precoding_init(q, max_prb, false); }
int srsran_dft_precoding_init_tx(srsran_dft_precoding_t* q, uint32_t max_prb) { return srsran_dft_precoding_init(q, max_prb, true); }

int srsran_dft_precoding(srsran_dft_precoding_t* q, cf_t* input, cf_t* output, uint32_t nof_prb, uint32_t nof_symbols)
{
  if (!q || !input || !output) return SRSRAN_ERROR_INVALID_INPUTS;
  if (!srsran_dft_precoding_valid_prb(nof_prb) || nof_prb == 0 || nof_prb > q->max_prb || !q->dft_plan[nof_prb].p) {
    B200_LOG_ERROR("Error invalid number of PRB (%u)", nof_prb); // dft_precoding.c:117
    return SRSRAN_ERROR;
  }
  if (nof_symbols == 0) return SRSRAN_SUCCESS;
  CompatDft* d = (CompatDft*)q->dft_plan[nof_prb].p;
  const int  N = 12 * (int)nof_prb;
  d->plan.gscale = 1.0f / sqrtf((float)N); // srsran_dft_plan_set_norm(true), applied by the kernel's last pass
  const int rc   = dft_exec(d, input, output, (int)nof_symbols, N, N);
  d->plan.gscale = 0.f;
  return rc;
}

} // extern "C"

// ===================================================================================================================
// srsran_ofdm_t (receive side)
//
// A receive object set up here keeps the batched engine's handle in fft_plan.p and marks itself with two self-pointers in
// the plan's (otherwise unused) buffer fields.  The mark matters because srsran_ofdm_set_freq_shift / set_normalize /
// set_non_mbsfn_region are shared with the TRANSMIT objects of the reference's own ofdm.c (srsran_ofdm_tx_init), whose
// fft_plan is an ordinary DFT plan of this library: those calls must then do what ofdm.c does on the struct, nothing else.
static bool ofdm_is_ours(const srsran_ofdm_t* q)
{
  return q && q->fft_plan.p && q->fft_plan.in == (const void*)q && q->fft_plan.out == (const void*)&q->fft_plan;
}

static srsran_b200_ofdm_t* ofdm_of(srsran_ofdm_t* q)
{
  return ofdm_is_ours(q) ? (srsran_b200_ofdm_t*)q->fft_plan.p : nullptr;
}

static int ofdm_apply(srsran_ofdm_t* q)
{
  srsran_b200_ofdm_cfg_t c;
  c.nof_prb          = q->cfg.nof_prb;
  c.cp_ext           = q->cfg.cp == SRSRAN_CP_EXT;
  c.symbol_sz        = q->cfg.symbol_sz;
  c.freq_shift_f     = q->cfg.freq_shift_f;
  c.rx_window_offset = q->cfg.rx_window_offset;
  c.normalize        = q->cfg.normalize;
  c.keep_dc          = q->cfg.keep_dc;
  int rc;
  if (ofdm_of(q)) {
    rc = srsran_b200_ofdm_rx_reconfigure(ofdm_of(q), &c);
  } else {
    srsran_b200_ofdm_t* h = nullptr;
    rc                    = srsran_b200_ofdm_rx_init(&h, compat_device(), &c);
    q->fft_plan.p         = h;
    q->fft_plan.in        = (void*)q;
    q->fft_plan.out       = (void*)&q->fft_plan;
  }
  if (rc != SRSRAN_SUCCESS) return SRSRAN_ERROR;
  uint32_t N, sf, ns, nre;
  srsran_b200_ofdm_rx_geometry(ofdm_of(q), &N, &sf, &ns, &nre);
  q->cfg.symbol_sz     = N;
  q->fft_plan.size     = (int)N;
  if (!q->fft_plan.init_size) q->fft_plan.init_size = (int)N;
  q->fft_plan.forward  = true;
  q->fft_plan.mirror   = true;
  q->fft_plan.norm     = q->cfg.normalize;
  q->fft_plan.dc       = (!q->cfg.keep_dc) && !isnormal(q->cfg.freq_shift_f);
  q->nof_symbols       = ns / 2;
  q->nof_symbols_mbsfn = 6;
  q->nof_re            = nre;
  q->nof_guards        = (N - nre) / 2;
  q->slot_sz           = sf / 2;
  q->sf_sz             = sf;
  if (q->cfg.nof_prb > q->max_prb) q->max_prb = q->cfg.nof_prb;
  if (isnormal(q->cfg.rx_window_offset)) {
    const int cp2      = (int)ceilf((((float)(q->cfg.cp == SRSRAN_CP_EXT ? 512 : 144)) * (float)N) / 2048.0f);
    q->window_offset_n = (uint32_t)roundf((float)cp2 * q->cfg.rx_window_offset);
  }
  return SRSRAN_SUCCESS;
}

extern "C" {

void srsran_use_standard_symbol_size(bool enabled)
{
  srsran_b200_use_standard_symbol_size(enabled ? 1 : 0);
}

int srsran_symbol_sz(uint32_t nof_prb)
{
  int n = srsran_b200_symbol_sz(nof_prb);
  return n > 0 ? n : SRSRAN_ERROR;
}

int srsran_ofdm_rx_init_cfg(srsran_ofdm_t* q, srsran_ofdm_cfg_t* cfg)
{
  if (!q || !cfg) return SRSRAN_ERROR_INVALID_INPUTS;
  if (cfg->sf_type == SRSRAN_SF_MBSFN) {
    B200_LOG_ERROR("MBSFN subframes are not part of the GPU path");
    return SRSRAN_ERROR;
  }
  if (q->max_prb > 0) { // already initialised: only the resizing parameters are taken (ofdm.c:50-58)
    q->cfg.cp        = cfg->cp;
    q->cfg.nof_prb   = cfg->nof_prb;
    q->cfg.symbol_sz = cfg->symbol_sz;
  } else {
    q->cfg = *cfg;
  }
  return ofdm_apply(q);
}

int srsran_ofdm_rx_init(srsran_ofdm_t* q, srsran_cp_t cp, cf_t* in_buffer, cf_t* out_buffer, uint32_t max_prb)
{
  if (!q) return SRSRAN_ERROR_INVALID_INPUTS;
  memset(q, 0, sizeof(*q)); // ofdm.c:245
  srsran_ofdm_cfg_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.cp         = cp;
  cfg.in_buffer  = in_buffer;
  cfg.out_buffer = out_buffer;
  cfg.nof_prb    = max_prb;
  cfg.sf_type    = SRSRAN_SF_NORM;
  return srsran_ofdm_rx_init_cfg(q, &cfg);
}

int srsran_ofdm_rx_init_mbsfn(srsran_ofdm_t*, srsran_cp_t, cf_t*, cf_t*, uint32_t)
{
  B200_LOG_ERROR("MBSFN subframes are not part of the GPU path");
  return SRSRAN_ERROR;
}

int srsran_ofdm_rx_set_prb(srsran_ofdm_t* q, srsran_cp_t cp, uint32_t nof_prb)
{
  if (!q) return SRSRAN_ERROR_INVALID_INPUTS;
  srsran_ofdm_cfg_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.cp      = cp;
  cfg.nof_prb = nof_prb;
  return srsran_ofdm_rx_init_cfg(q, &cfg); // ofdm.c:309-315 (symbol_sz 0: derived again from nof_prb)
}

void srsran_ofdm_rx_free(srsran_ofdm_t* q)
{
  if (!q) return;
  if (ofdm_of(q)) srsran_b200_ofdm_rx_free(ofdm_of(q));
  memset(q, 0, sizeof(*q)); // ofdm.c:240
}

void srsran_ofdm_rx_sf(srsran_ofdm_t* q)
{
  if (!q || !ofdm_of(q)) return;
  srsran_b200_ofdm_rx_sf_batch(ofdm_of(q), q->cfg.in_buffer, q->cfg.out_buffer, 1, 0, nullptr);
}

void srsran_ofdm_rx_sf_ng(srsran_ofdm_t* q, cf_t* input, cf_t* output)
{
  if (!q || !ofdm_of(q)) return;
  srsran_b200_ofdm_rx_sf_batch(ofdm_of(q), input, output, 1, 0, nullptr);
}

int srsran_ofdm_set_freq_shift(srsran_ofdm_t* q, float freq_shift)
{
  if (!q) return SRSRAN_ERROR_INVALID_INPUTS;
  q->cfg.freq_shift_f = freq_shift;
  if (ofdm_is_ours(q)) return ofdm_apply(q);
  // a transmit object of the reference's ofdm.c: ofdm.c:334-360 on its own struct (the per-sample rotation it multiplies its
  // output with, in the reference's precision: the angle in double, rounded to float, cexpf)
  if (!isnormal(freq_shift)) {
    q->fft_plan.dc = true;
    return SRSRAN_SUCCESS;
  }
  if (!q->shift_buffer) return SRSRAN_ERROR;
  const uint32_t N   = q->cfg.symbol_sz;
  float*         ptr = (float*)q->shift_buffer;
  for (uint32_t n = 0; n < 2; n++) {
    for (uint32_t i = 0; i < q->nof_symbols; i++) {
      // SRSRAN_CP_LEN_NORM / SRSRAN_CP_LEN_EXT (phy_common.h:125-128): ceil(c * symbol_sz / 2048) with c = 160 / 144 / 512
      const float    c     = q->cfg.cp == SRSRAN_CP_NORM ? (i == 0 ? 160.0f : 144.0f) : 512.0f;
      const uint32_t cplen = (uint32_t)(int)ceilf((c * (float)N) / 2048.0f);
      for (uint32_t t = 0; t < N + cplen; t++) {
        const float a = (float)(2.0 * M_PI * (double)((float)t - (float)cplen) * (double)freq_shift / (double)N);
        sincosf(a, &ptr[2 * t + 1], &ptr[2 * t]);
      }
      ptr += 2 * (N + cplen);
    }
  }
  q->fft_plan.dc = false; // ofdm.c:359
  return SRSRAN_SUCCESS;
}

void srsran_ofdm_set_normalize(srsran_ofdm_t* q, bool normalize_enable)
{
  if (!q) return;
  if (!ofdm_is_ours(q)) { // ofdm.c:557-560
    q->fft_plan.norm = normalize_enable;
    return;
  }
  q->cfg.normalize = normalize_enable;
  ofdm_apply(q);
}

void srsran_ofdm_set_non_mbsfn_region(srsran_ofdm_t* q, uint8_t non_mbsfn_region)
{
  if (q) q->non_mbsfn_region = non_mbsfn_region;
}

} // extern "C"
