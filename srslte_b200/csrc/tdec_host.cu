// Host side of the batched turbo decoder: workspace carving, pass scheduling, host<->device pipelining and the
// extern "C" entries declared in include/srslte_b200.h.
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <new>
#include <utility>

#include "../../include/srslte_b200.h"
#include "b200_runtime.h"
#include "lte_tables.h"
#include "tdec_engine.h"
#include "tdec_kernels.h"

namespace b200 {

std::atomic<uint64_t> g_kernel_launches{0};

size_t TdecEngine::workspace_bytes(int K, uint32_t ncb)
{
  const size_t ntiles = (ncb + TDEC_TILE_CB - 1) / TDEC_TILE_CB;
  const size_t vrows  = ntiles * (size_t)((K + 4) / 4) * 32 * sizeof(u4);
  size_t       total  = 0;
  total += 3 * (vrows + 256);                                      // S, P0, P1
  total += 3 * (ntiles * (size_t)(K / 8 + 1) * 32 * sizeof(u4) + 256); // S8, P08, P18
  total += ntiles * sizeof(uint32_t) + 256;                        // fmt
  total += ntiles * 32 * sizeof(u4) + 256;                         // S2T
  total += ntiles * (size_t)K * 32 * sizeof(uint32_t) + 256;       // E
  total += ntiles * (size_t)(K / 8) * 2 * 32 * sizeof(u4) + 256;   // CK
  total += ntiles * (size_t)(K / 8) * 32 * sizeof(uint16_t) + 256; // HB
  total += ntiles * TDEC_TILE_CB * sizeof(CbStatus) + 256;         // status
  return total;
}

// share of the trellis windows below the split (tdec_core.h: tdec_split), per kind of pass: 0 = first pass, 1 = DEC2 (odd),
// 2 = DEC1 (even, with a-priori input).  SRSLTE_B200_TDEC_SPLIT[_FIRST|_DEC2|_DEC1] are tuning knobs.
static int split_percent(int kind)
{
  static int pct[3] = {-1, -1, -1};
  if (pct[0] < 0) {
    static const char* names[3] = {"SRSLTE_B200_TDEC_SPLIT_FIRST", "SRSLTE_B200_TDEC_SPLIT_DEC2", "SRSLTE_B200_TDEC_SPLIT_DEC1"};
    static const int   dflt[3]  = {51, 51, 51};
    const char*        all      = getenv("SRSLTE_B200_TDEC_SPLIT");
    for (int k = 2; k >= 0; k--) {
      const char* e = getenv(names[k]);
      int         p = e ? atoi(e) : (all ? atoi(all) : dflt[k]);
      if (p < 1 || p > 99) p = dflt[k];
      pct[k] = p;
    }
  }
  return pct[kind];
}

int TdecEngine::carve(DeviceArena& arena, int K, uint32_t ncb, TdecView& v) const
{
  const size_t ntiles = (ncb + TDEC_TILE_CB - 1) / TDEC_TILE_CB;
  const size_t vrows  = ntiles * (size_t)((K + 4) / 4) * 32 * sizeof(u4);
  v.K                 = K;
  v.ntiles            = (int)ntiles;
  v.S                 = (u4*)arena.take(vrows);
  v.P0                = (u4*)arena.take(vrows);
  v.P1                = (u4*)arena.take(vrows);
  const size_t rows8  = ntiles * (size_t)(K / 8 + 1) * 32 * sizeof(u4);
  v.S8                = (u4*)arena.take(rows8);
  v.P08               = (u4*)arena.take(rows8);
  v.P18               = (u4*)arena.take(rows8);
  v.fmt               = (uint32_t*)arena.take(ntiles * sizeof(uint32_t));
  v.ws                = tdec_split(K, split_percent(2));
  v.S2T               = (u4*)arena.take(ntiles * 32 * sizeof(u4));
  v.E                 = (uint32_t*)arena.take(ntiles * (size_t)K * 32 * sizeof(uint32_t));
  v.CK                = (u4*)arena.take(ntiles * (size_t)(K / 8) * 2 * 32 * sizeof(u4));
  v.HB                = (uint16_t*)arena.take(ntiles * (size_t)(K / 8) * 32 * sizeof(uint16_t));
  v.status            = (CbStatus*)arena.take(ntiles * TDEC_TILE_CB * sizeof(CbStatus));
  if (!v.S || !v.P0 || !v.P1 || !v.S8 || !v.P08 || !v.P18 || !v.fmt || !v.S2T || !v.E || !v.CK || !v.HB || !v.status) {
    B200_LOG_ERROR("decoder workspace too small");
    return B200_ERROR;
  }
  return B200_SUCCESS;
}

int TdecEngine::init(int device, uint32_t max_cb_hint)
{
  ctx = device_context(device);
  if (!ctx) {
    B200_LOG_ERROR("no usable CUDA device %d (this library has no CPU fallback)", device);
    return B200_ERROR;
  }
  B200_CUDA_TRY(cudaSetDevice(device));
  for (int i = 0; i < 2; i++) {
    B200_CUDA_TRY(cudaStreamCreateWithFlags(&pipe_stream[i], cudaStreamNonBlocking));
  }

  if (max_cb_hint) {
    if (arena.reserve(workspace_bytes(MAX_CB_LEN, max_cb_hint)) != B200_SUCCESS) {
      return B200_ERROR;
    }
  }
  return B200_SUCCESS;
}

void TdecEngine::destroy()
{
  if (ctx) {
    cudaSetDevice(ctx->device);
  }
  for (int i = 0; i < 2; i++) {
    if (pipe_stream[i]) cudaStreamDestroy(pipe_stream[i]);
    pipe_arena[i].release();
    pipe_io[i].release();
  }
  arena.release();
  prof_reset(false);
}

void TdecEngine::prof_begin(int cls, cudaStream_t st)
{
  if (!profiling || spans.size() >= 8192) return;
  ProfSpan s;
  s.cls = cls;
  if (cudaEventCreate(&s.a) != cudaSuccess || cudaEventCreate(&s.b) != cudaSuccess) return;
  cudaEventRecord(s.a, st);
  spans.push_back(s);
}

void TdecEngine::prof_end(cudaStream_t st)
{
  if (!profiling || spans.empty() || spans.size() > 8192) return;
  cudaEventRecord(spans.back().b, st);
}

void TdecEngine::prof_reset(bool enable)
{
  for (auto& s : spans) {
    cudaEventDestroy(s.a);
    cudaEventDestroy(s.b);
  }
  spans.clear();
  profiling = enable;
}

// Per class: the time during which at least one span of that class was open (spans of concurrent streams overlap, so
// their plain sum would count that time twice) and the number of spans.
int TdecEngine::prof_get(double* ms_by_class, uint64_t* launches_by_class)
{
  for (int i = 0; i < 3; i++) {
    ms_by_class[i]       = 0;
    launches_by_class[i] = 0;
  }
  B200_CUDA_TRY(cudaDeviceSynchronize());
  if (spans.empty()) return B200_SUCCESS;
  std::vector<std::pair<double, double>> iv[3];
  for (auto& s : spans) {
    float t0 = 0, t1 = 0;
    // event times relative to the first span's start (negative when a concurrent stream started earlier)
    if (cudaEventElapsedTime(&t0, spans[0].a, s.a) != cudaSuccess) {
      float r = 0;
      if (cudaEventElapsedTime(&r, s.a, spans[0].a) != cudaSuccess) continue;
      t0 = -r;
    }
    if (cudaEventElapsedTime(&t1, s.a, s.b) != cudaSuccess) continue;
    iv[s.cls].push_back({(double)t0, (double)t0 + (double)t1});
    launches_by_class[s.cls]++;
  }
  for (int c = 0; c < 3; c++) {
    std::sort(iv[c].begin(), iv[c].end());
    double cur_a = 0, cur_b = -1e300;
    for (auto& x : iv[c]) {
      if (x.first > cur_b) {
        if (cur_b > -1e299) ms_by_class[c] += cur_b - cur_a;
        cur_a = x.first;
        cur_b = x.second;
      } else if (x.second > cur_b) {
        cur_b = x.second;
      }
    }
    if (cur_b > -1e299) ms_by_class[c] += cur_b - cur_a;
  }
  return B200_SUCCESS;
}

// Everything on `stream`, all pointers device memory.
int TdecEngine::run_device(DeviceArena&   ws,
                           const int16_t* llr_dev,
                           uint32_t       ncb,
                           int            K,
                           int            cb_idx,
                           uint32_t       max_passes,
                           int            crc_kind,
                           int            early_stop,
                           uint8_t*       out_dev,
                           uint8_t*       crc_ok_dev,
                           uint8_t*       npass_dev,
                           cudaStream_t   stream,
                           const uint64_t* llr_offsets_dev,
                           bool            offsets_aligned8,
                           bool            reset_ws)
{
  TdecView v;
  if (reset_ws) ws.reset();
  if (carve(ws, K, ncb, v) != B200_SUCCESS) {
    return B200_ERROR;
  }
  v.qpp_fwd    = ctx->qpp_fwd(cb_idx);
  v.crc_nat    = nullptr;
  v.crc_perm   = nullptr;
  if (crc_kind != SRSRAN_B200_CRC_NONE) {
    if (ctx->crc_visit(cb_idx, crc_kind == SRSRAN_B200_CRC24A ? 0 : 1, &v.crc_nat, &v.crc_perm) != B200_SUCCESS) {
      return B200_ERROR;
    }
  }
  v.early_stop = early_stop ? 1 : 0;
  v.max_pass   = (int)max_passes;

  prof_begin(0, stream);
  launch_load_natural(v, llr_dev, llr_offsets_dev, llr_offsets_dev ? offsets_aligned8 : ((reinterpret_cast<uintptr_t>(llr_dev) & 7u) == 0), ncb, stream);
  prof_end(stream);
  g_kernel_launches += 2;
  for (uint32_t p = 0; p < max_passes; p++) {
    prof_begin(1, stream);
    v.ws = tdec_split(K, split_percent(p == 0 ? 0 : ((p & 1) ? 1 : 2))); // the checkpoints are per-pass scratch: each kind of pass splits where it balances
    launch_siso_pass(v, (int)p, stream);
    prof_end(stream);
    g_kernel_launches++;
  }
  prof_begin(2, stream);
  launch_decide(v, ctx->qpp_rev(cb_idx), out_dev, crc_ok_dev, npass_dev, nullptr, ncb, stream);
  prof_end(stream);
  g_kernel_launches++;
  B200_CUDA_TRY(cudaGetLastError());
  return B200_SUCCESS;
}

int TdecEngine::run(const int16_t* llr,
                    uint32_t       ncb,
                    uint32_t       K,
                    uint32_t       max_passes,
                    int            crc_kind,
                    int            early_stop,
                    uint8_t*       out,
                    uint8_t*       crc_ok,
                    uint8_t*       npass,
                    uint32_t       flags,
                    cudaStream_t   stream)
{
  const int cb_idx = cb_index_exact(K);
  if (cb_idx < 0) {
    B200_LOG_ERROR("Invalid CB length %u", K); // turbodecoder.c:520
    return B200_ERROR;
  }
  if (!llr || !out || max_passes < 1 || max_passes > 255 || crc_kind < 0 || crc_kind > 2) {
    return B200_ERROR_INVALID_INPUTS;
  }
  if (ncb == 0) {
    return B200_SUCCESS;
  }
  B200_CUDA_TRY(cudaSetDevice(ctx->device));

  const bool   llr8 = (flags & SRSRAN_B200_FLAG_LLR_INT8) != 0;
  const size_t nllr = 3 * (size_t)K + 12;
  if (flags & SRSRAN_B200_FLAG_DEVICE_PTRS) {
    if (arena.reserve(workspace_bytes((int)K, ncb)) != B200_SUCCESS) {
      return B200_ERROR;
    }
    if (llr8) { // widen into a scratch vector first (one extra pass over 3 bytes per value, ~6 % of an 8-pass decode)
      if (pipe_io[0].reserve((size_t)ncb * nllr * sizeof(int16_t) + 1024) != B200_SUCCESS) return B200_ERROR;
      pipe_io[0].reset();
      int16_t* wide = (int16_t*)pipe_io[0].take((size_t)ncb * nllr * sizeof(int16_t));
      launch_widen_i8(reinterpret_cast<const int8_t*>(llr), wide, (size_t)ncb * nllr, stream);
      g_kernel_launches++;
      return run_device(arena, wide, ncb, (int)K, cb_idx, max_passes, crc_kind, early_stop, out, crc_ok, npass, stream);
    }
    return run_device(arena, llr, ncb, (int)K, cb_idx, max_passes, crc_kind, early_stop, out, crc_ok, npass, stream);
  }

  // Host pointers: cut the batch into chunks and ping-pong two streams so the copy of chunk c+1 overlaps the decode
  // of chunk c.  6 input bytes per info bit cross PCIe here, which is what bounds this path.
  const size_t   nb    = K / 8;
  const uint32_t chunk = ncb < 2 * kPipeChunkCb ? (ncb + 1) / 2 : kPipeChunkCb;
  for (int i = 0; i < 2; i++) {
    if (pipe_arena[i].reserve(workspace_bytes((int)K, chunk)) != B200_SUCCESS ||
        pipe_io[i].reserve(chunk * (nllr * (sizeof(int16_t) + (llr8 ? 1 : 0)) + nb + 2) + 2048) != B200_SUCCESS) {
      return B200_ERROR;
    }
  }
  int      rc = B200_SUCCESS;
  uint32_t c  = 0;
  for (uint32_t first = 0; first < ncb && rc == B200_SUCCESS; first += chunk, c++) {
    const int      s = (int)(c & 1);
    const uint32_t n = (ncb - first) < chunk ? (ncb - first) : chunk;
    cudaStream_t   st = pipe_stream[s];
    // the previous user of this slot must have drained before its buffers are overwritten
    B200_CUDA_TRY(cudaStreamSynchronize(st));
    pipe_io[s].reset();
    int16_t* d_llr = (int16_t*)pipe_io[s].take(n * nllr * sizeof(int16_t));
    uint8_t* d_out = (uint8_t*)pipe_io[s].take(n * nb);
    uint8_t* d_ok  = (uint8_t*)pipe_io[s].take(n);
    uint8_t* d_np  = (uint8_t*)pipe_io[s].take(n);
    if (llr8) {
      int8_t* d_llr8 = (int8_t*)pipe_io[s].take(n * nllr);
      B200_CUDA_TRY(cudaMemcpyAsync(d_llr8, reinterpret_cast<const int8_t*>(llr) + first * nllr, n * nllr, cudaMemcpyHostToDevice, st));
      launch_widen_i8(d_llr8, d_llr, (size_t)n * nllr, st);
      g_kernel_launches++;
    } else {
      B200_CUDA_TRY(cudaMemcpyAsync(d_llr, llr + first * nllr, n * nllr * sizeof(int16_t), cudaMemcpyHostToDevice, st));
    }
    rc = run_device(pipe_arena[s], d_llr, n, (int)K, cb_idx, max_passes, crc_kind, early_stop, d_out, d_ok, d_np, st);
    if (rc != B200_SUCCESS) break;
    B200_CUDA_TRY(cudaMemcpyAsync(out + first * nb, d_out, n * nb, cudaMemcpyDeviceToHost, st));
    if (crc_ok) B200_CUDA_TRY(cudaMemcpyAsync(crc_ok + first, d_ok, n, cudaMemcpyDeviceToHost, st));
    if (npass) B200_CUDA_TRY(cudaMemcpyAsync(npass + first, d_np, n, cudaMemcpyDeviceToHost, st));
  }
  for (int i = 0; i < 2; i++) {
    B200_CUDA_TRY(cudaStreamSynchronize(pipe_stream[i]));
  }
  return rc;
}

} // namespace b200

// ---------------------------------------------------------------------------------------------------------------
using namespace b200;

struct srsran_b200_tdec {
  TdecEngine eng;
};

extern "C" {

int srsran_b200_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    return 0;
  }
  return n;
}

uint64_t srsran_b200_kernel_launches(void)
{
  return g_kernel_launches.load();
}

int srsran_b200_tdec_init(srsran_b200_tdec_t** h, int device, uint32_t max_cb_hint)
{
  if (!h) {
    return B200_ERROR_INVALID_INPUTS;
  }
  *h                    = nullptr;
  srsran_b200_tdec_t* q = new (std::nothrow) srsran_b200_tdec_t();
  if (!q) {
    return B200_ERROR;
  }
  if (q->eng.init(device, max_cb_hint) != B200_SUCCESS) {
    q->eng.destroy();
    delete q;
    return B200_ERROR;
  }
  *h = q;
  return B200_SUCCESS;
}

void srsran_b200_tdec_free(srsran_b200_tdec_t* h)
{
  if (h) {
    h->eng.destroy();
    delete h;
  }
}

int srsran_b200_tdec_run(srsran_b200_tdec_t* h,
                         const int16_t*      llr,
                         uint32_t            ncb,
                         uint32_t            K,
                         uint32_t            max_passes,
                         int                 crc_kind,
                         int                 early_stop,
                         uint8_t*            out,
                         uint8_t*            crc_ok,
                         uint8_t*            npass,
                         uint32_t            flags,
                         void*               stream)
{
  if (!h) {
    return B200_ERROR_INVALID_INPUTS;
  }
  return h->eng.run(llr, ncb, K, max_passes, crc_kind, early_stop, out, crc_ok, npass, flags, (cudaStream_t)stream);
}

int srsran_b200_tdec_resident_tiles_per_sm(void)
{
  return siso_resident_tiles_per_sm();
}

void srsran_b200_tdec_profile_reset(srsran_b200_tdec_t* h, int enable)
{
  if (h) {
    h->eng.prof_reset(enable != 0);
  }
}

int srsran_b200_tdec_profile_get(srsran_b200_tdec_t* h, double* ms_by_class, uint64_t* launches_by_class)
{
  if (!h || !ms_by_class || !launches_by_class) {
    return B200_ERROR_INVALID_INPUTS;
  }
  return h->eng.prof_get(ms_by_class, launches_by_class);
}

} // extern "C"
