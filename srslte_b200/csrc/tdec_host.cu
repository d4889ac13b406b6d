// Host side of the batched turbo decoder: workspace carving, pass scheduling, host<->device pipelining and the
// extern "C" entries declared in include/srslte_b200.h.
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <new>
#include <vector>
#include <utility>

#include "../../include/srslte_b200.h"
#include "b200_runtime.h"
#include "lte_tables.h"
#include "tdec_engine.h"
#include "tdec_kernels.h"

namespace b200 {

std::atomic<uint64_t> g_kernel_launches{0};

// device bytes of one tile's own arrays (each carve-out is rounded up to 256 bytes)
static size_t tile_bytes(int K, bool with_int16)
{
  auto r256 = [](size_t b) { return (b + 255) / 256 * 256; };
  size_t total = 0;
  if (with_int16) total += 3 * r256((size_t)((K + 4) / 4) * 32 * sizeof(u4));  // S, P0, P1
  total += 3 * r256((size_t)(K / 8 + 1) * 32 * sizeof(u4));    // S8, P08, P18
  total += r256((size_t)K * 32 * sizeof(uint32_t));            // E
  total += r256((size_t)(K / 8) * 2 * 32 * sizeof(u4));        // CK
  return total;
}

size_t TdecEngine::workspace_bytes(const std::vector<TdecGroupSpec>& groups, bool with_int16)
{
  size_t total = 0, ntiles = 0, hb_rows = 0;
  for (const TdecGroupSpec& g : groups) {
    const size_t nt = (g.ncb + TDEC_TILE_CB - 1) / TDEC_TILE_CB;
    total += nt * tile_bytes(g.K, with_int16);
    ntiles += nt;
    hb_rows += nt * (size_t)(g.K / 8);
  }
  total += hb_rows * 32 * sizeof(uint16_t) + 256;                         // HB
  total += ntiles * (sizeof(uint32_t) * 4 + 32 * sizeof(u4) + 32 * sizeof(LaneMap) + TDEC_TILE_CB * sizeof(CbStatus) +
                     sizeof(TileDesc) + 32 * sizeof(MoveRec) + 32 * sizeof(uint32_t));             // fmt, mask, pref, S2T, lanes, status, descriptors, moves
  total += groups.size() * (sizeof(TileGroup) + sizeof(GroupPlan));
  int max_K = 0;
  for (const TdecGroupSpec& g : groups) max_K = std::max(max_K, g.K);
  total += std::min<size_t>(ntiles, 256) * (size_t)(max_K / 8) * 64 * sizeof(u4); // low-latency pass: one alpha checkpoint slot per thread block
  return total + 20 * 256 + 4096;
}

size_t TdecEngine::workspace_bytes(int K, uint32_t ncb)
{
  std::vector<TdecGroupSpec> g(1);
  g[0] = TdecGroupSpec{K, 0, 0, ncb, 0, 0, 0};
  return workspace_bytes(g);
}

static std::mutex                                 g_pool_mutex;
static std::map<int, std::vector<TdecWorkspace*>> g_pool_free;

TdecWorkspace* workspace_acquire(int device)
{
  std::lock_guard<std::mutex> lk(g_pool_mutex);
  auto&                       v = g_pool_free[device];
  if (!v.empty()) {
    TdecWorkspace* w = v.back();
    v.pop_back();
    return w;
  }
  return new (std::nothrow) TdecWorkspace();
}

void workspace_release(int device, TdecWorkspace* w)
{
  if (!w) return;
  std::lock_guard<std::mutex> lk(g_pool_mutex);
  g_pool_free[device].push_back(w); // kept for the next decode on this device (its descriptors stay cached); never freed
}

void TdecWorkspace::release()
{
  arena.release();
  stage.release();
  if (uploaded) cudaEventDestroy(uploaded);
  uploaded = nullptr;
  if (h_err) cudaFreeHost(h_err);
  h_err = nullptr;
  cached.clear();
  cached_generation = ~0ull;
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

// Tiles in order of descending K (longest first: the block scheduler hands out CTAs in index order, so the short tiles of a
// mixed batch fill the tail of every pass), every group's tiles consecutive.
int TdecEngine::prepare(TdecWorkspace& w, const std::vector<TdecGroupSpec>& groups, cudaStream_t stream)
{
  const bool with_int16 = !w.int16_on_demand || w.have_int16;
  if (w.arena.reserve(workspace_bytes(groups, with_int16)) != B200_SUCCESS) return B200_ERROR;
  if (w.int16_on_demand && !w.h_err) {
    B200_CUDA_TRY(cudaHostAlloc(&w.h_err, sizeof(uint32_t), cudaHostAllocDefault));
    *w.h_err = 0;
  }
  if (w.cached_generation == w.arena.generation && w.cached == groups && w.cached_int16 == with_int16 && w.plan.v.tiles != nullptr) {
    return B200_SUCCESS; // same batch shape in the same memory: the descriptors on the device are still right
  }
  std::vector<uint32_t> order(groups.size());
  for (uint32_t i = 0; i < order.size(); i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return groups[a].K > groups[b].K; });
  size_t ntiles = 0, hb_rows = 0;
  int    max_K  = 0;
  for (const TdecGroupSpec& g : groups) {
    ntiles += (g.ncb + TDEC_TILE_CB - 1) / TDEC_TILE_CB;
    hb_rows += (size_t)((g.ncb + TDEC_TILE_CB - 1) / TDEC_TILE_CB) * (size_t)(g.K / 8);
    max_K = std::max(max_K, g.K);
  }
  if (ntiles == 0 || ntiles > 65535 || hb_rows * 32 > 0xFFFFFFFFull) {
    B200_LOG_ERROR("batch of %zu tiles is outside what one decode call handles (1..65535)", ntiles);
    return B200_ERROR_INVALID_INPUTS;
  }
  DeviceArena& a = w.arena;
  a.reset();
  TdecPlan& p     = w.plan;
  p               = TdecPlan{};
  p.max_K         = max_K;
  p.ngroups       = (uint32_t)groups.size();
  p.v.ntiles      = (int)ntiles;
  TileDesc* d_tiles = (TileDesc*)a.take(ntiles * sizeof(TileDesc));
  p.groups        = (TileGroup*)a.take(groups.size() * sizeof(TileGroup));
  p.plans         = (GroupPlan*)a.take(groups.size() * sizeof(GroupPlan));
  p.v.fmt         = (uint32_t*)a.take(ntiles * sizeof(uint32_t));
  p.v.err         = (uint32_t*)a.take(sizeof(uint32_t));
  p.v.ctl         = (uint32_t*)a.take(TDEC_CTL_WORDS * sizeof(uint32_t));
  p.v.ll_ck_slot  = (uint32_t)(max_K / 8) * 64u;
  p.v.ll_ck       = (u4*)a.take((size_t)std::min<size_t>(ntiles, (size_t)sm_count) * p.v.ll_ck_slot * sizeof(u4));
  p.mask          = (uint32_t*)a.take(ntiles * sizeof(uint32_t));
  p.pref          = (uint32_t*)a.take(ntiles * sizeof(uint32_t));
  p.move_counter  = (uint32_t*)a.take(sizeof(uint32_t));
  p.moves         = (MoveRec*)a.take(ntiles * 32 * sizeof(MoveRec));
  p.gsrc          = (uint32_t*)a.take(ntiles * 32 * sizeof(uint32_t));
  p.v.S2T         = (u4*)a.take(ntiles * 32 * sizeof(u4));
  p.v.lanes       = (LaneMap*)a.take(ntiles * 32 * sizeof(LaneMap));
  p.v.status      = (CbStatus*)a.take(ntiles * TDEC_TILE_CB * sizeof(CbStatus));
  p.v.HB          = (uint16_t*)a.take(hb_rows * 32 * sizeof(uint16_t));
  p.v.tiles       = d_tiles;
  if (!d_tiles || !p.groups || !p.plans || !p.v.fmt || !p.v.err || !p.v.ctl || !p.v.ll_ck || !p.mask || !p.pref || !p.move_counter || !p.moves || !p.gsrc || !p.v.S2T || !p.v.lanes ||
      !p.v.status || !p.v.HB) {
    B200_LOG_ERROR("decoder workspace too small");
    return B200_ERROR;
  }
  // the staging buffer may still be the source of the previous batch's descriptor copy
  if (!w.uploaded) B200_CUDA_TRY(cudaEventCreateWithFlags(&w.uploaded, cudaEventDisableTiming));
  else B200_CUDA_TRY(cudaEventSynchronize(w.uploaded));
  const size_t stage_bytes = ntiles * sizeof(TileDesc) + groups.size() * sizeof(TileGroup) + 256;
  if (w.stage.reserve(stage_bytes) != B200_SUCCESS) return B200_ERROR;
  w.stage.reset();
  TileDesc*  h_tiles  = (TileDesc*)w.stage.take(ntiles * sizeof(TileDesc));
  TileGroup* h_groups = (TileGroup*)w.stage.take(groups.size() * sizeof(TileGroup));
  if (!h_tiles || !h_groups) return B200_ERROR;
  uint32_t tile = 0, hb_row = 0;
  for (uint32_t oi = 0; oi < order.size(); oi++) {
    const TdecGroupSpec& g  = groups[order[oi]];
    const uint32_t       nt = (g.ncb + TDEC_TILE_CB - 1) / TDEC_TILE_CB;
    h_groups[oi]            = TileGroup{tile, nt};
    const CrcPow *cn = nullptr, *cp = nullptr;
    if (g.crc_kind != SRSRAN_B200_CRC_NONE) {
      if (ctx->crc_visit(g.cb_idx, g.crc_kind == SRSRAN_B200_CRC24A ? 0 : 1, &cn, &cp) != B200_SUCCESS) return B200_ERROR;
    }
    const int    K     = g.K;
    const size_t vrows = (size_t)((K + 4) / 4) * 32 * sizeof(u4), rows8 = (size_t)(K / 8 + 1) * 32 * sizeof(u4);
    for (uint32_t t = 0; t < nt; t++, tile++) {
      TileDesc& d = h_tiles[tile];
      d.S8        = (u4*)a.take(rows8);
      d.P08       = (u4*)a.take(rows8);
      d.P18       = (u4*)a.take(rows8);
      d.S         = with_int16 ? (u4*)a.take(vrows) : nullptr;
      d.P0        = with_int16 ? (u4*)a.take(vrows) : nullptr;
      d.P1        = with_int16 ? (u4*)a.take(vrows) : nullptr;
      d.E         = (uint32_t*)a.take((size_t)K * 32 * sizeof(uint32_t));
      d.CK        = (u4*)a.take((size_t)(K / 8) * 2 * 32 * sizeof(u4));
      if (!d.S8 || !d.P08 || !d.P18 || (with_int16 && (!d.S || !d.P0 || !d.P1)) || !d.E || !d.CK) {
        B200_LOG_ERROR("decoder workspace too small");
        return B200_ERROR;
      }
      d.qpp_fwd  = ctx->qpp_fwd(g.cb_idx);
      d.qpp_rev  = ctx->qpp_rev(g.cb_idx);
      d.crc_nat  = cn;
      d.crc_perm = cp;
      d.llr_off  = g.llr_off + (uint64_t)t * TDEC_TILE_CB * (3ull * K + 12ull);
      d.out_off  = g.out_off + (uint64_t)t * TDEC_TILE_CB * (uint64_t)(K / 8);
      d.K        = (uint32_t)K;
      d.hb_row0  = hb_row;
      d.cb0      = g.cb0 + t * TDEC_TILE_CB;
      d.nblk     = std::min<uint32_t>(TDEC_TILE_CB, g.ncb - t * TDEC_TILE_CB);
      d.group    = oi;
      d.pad[0] = d.pad[1] = d.pad[2] = 0;
      hb_row += (uint32_t)(K / 8);
    }
  }
  B200_CUDA_TRY(cudaMemcpyAsync(d_tiles, h_tiles, ntiles * sizeof(TileDesc), cudaMemcpyHostToDevice, stream));
  B200_CUDA_TRY(cudaMemcpyAsync(p.groups, h_groups, groups.size() * sizeof(TileGroup), cudaMemcpyHostToDevice, stream));
  B200_CUDA_TRY(cudaEventRecord(w.uploaded, stream));
  w.cached            = groups;
  w.cached_generation = a.generation;
  w.cached_int16      = with_int16;
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

  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) sm_count = prop.multiProcessorCount;
  if (max_cb_hint) {
    if (ws.arena.reserve(workspace_bytes(MAX_CB_LEN, max_cb_hint)) != B200_SUCCESS) {
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
    pipe_ws[i].release();
    pipe_io[i].release();
  }
  ws.release();
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
int TdecEngine::prof_get(double* ms_by_class, uint64_t* launches_by_class, int nclasses)
{
  if (nclasses > kProfClasses) nclasses = kProfClasses;
  for (int i = 0; i < nclasses; i++) {
    ms_by_class[i]       = 0;
    launches_by_class[i] = 0;
  }
  B200_CUDA_TRY(cudaDeviceSynchronize());
  if (spans.empty()) return B200_SUCCESS;
  std::vector<std::pair<double, double>> iv[kProfClasses];
  for (auto& s : spans) {
    float t0 = 0, t1 = 0;
    // event times relative to the first span's start (negative when a concurrent stream started earlier)
    if (cudaEventElapsedTime(&t0, spans[0].a, s.a) != cudaSuccess) {
      float r = 0;
      if (cudaEventElapsedTime(&r, s.a, spans[0].a) != cudaSuccess) continue;
      t0 = -r;
    }
    if (cudaEventElapsedTime(&t1, s.a, s.b) != cudaSuccess) continue;
    if (s.cls >= nclasses) continue;
    iv[s.cls].push_back({(double)t0, (double)t0 + (double)t1});
    launches_by_class[s.cls]++;
  }
  for (int c = 0; c < nclasses; c++) {
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

void TdecEngine::tile_layout(const std::vector<TdecGroupSpec>& groups, std::vector<uint32_t>& first_tile)
{
  std::vector<uint32_t> order(groups.size());
  for (uint32_t i = 0; i < order.size(); i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return groups[a].K > groups[b].K; });
  first_tile.assign(groups.size(), 0);
  uint32_t tile = 0;
  for (uint32_t oi = 0; oi < order.size(); oi++) {
    first_tile[order[oi]] = tile;
    tile += (groups[order[oi]].ncb + TDEC_TILE_CB - 1) / TDEC_TILE_CB;
  }
}

int TdecEngine::begin_batch(TdecWorkspace& w, const std::vector<TdecGroupSpec>& groups, cudaStream_t stream)
{
  int rc = prepare(w, groups, stream);
  if (rc != B200_SUCCESS) return rc;
  const TdecView& v = w.plan.v;
  B200_CUDA_TRY(cudaMemsetAsync(v.fmt, 0, (size_t)v.ntiles * sizeof(uint32_t), stream));
  B200_CUDA_TRY(cudaMemsetAsync(v.err, 0, sizeof(uint32_t), stream));
  return B200_SUCCESS;
}

// Everything on `stream`, all pointers device memory.
int TdecEngine::run_groups(TdecWorkspace&                    w,
                           const int16_t*                    llr_dev,
                           const std::vector<TdecGroupSpec>& groups,
                           uint32_t                          max_passes,
                           int                               early_stop,
                           uint8_t*                          out_dev,
                           uint8_t*                          crc_ok_dev,
                           uint8_t*                          npass_dev,
                           cudaStream_t                      stream,
                           const uint64_t*                   llr_offsets_dev,
                           bool                              offsets_aligned8,
                           bool                              tiles_preloaded)
{
  int rc = prepare(w, groups, stream);
  if (rc != B200_SUCCESS) return rc;
  TdecPlan& p    = w.plan;
  TdecView  v    = p.v;
  v.early_stop   = early_stop ? 1 : 0;
  v.max_pass     = (int)max_passes;
  bool aligned8  = llr_offsets_dev ? offsets_aligned8 : ((reinterpret_cast<uintptr_t>(llr_dev) & 7u) == 0);
  if (!llr_offsets_dev) {
    for (const TdecGroupSpec& g : groups) aligned8 = aligned8 && (g.llr_off % 4 == 0); // (3K+12) is a multiple of 4 already
  }
  // Re-packing the running lanes between passes only pays when the SMs hold several tiles each (then freed tiles shorten the
  // next pass); SRSLTE_B200_TDEC_NO_COMPACT switches it off (comparison runs).
  const bool  no_compact = getenv("SRSLTE_B200_TDEC_NO_COMPACT") != nullptr;
  const char* min_env    = getenv("SRSLTE_B200_TDEC_COMPACT_MIN_TILES"); // tests: re-pack small batches too
  const int   min_tiles  = min_env ? atoi(min_env) : 2 * sm_count;
  const bool  compact    = early_stop && !no_compact && v.ntiles >= min_tiles;
  // Which SISO kernel: batches of at most one tile per SM are latency-bound and go to the low-latency kernel (a whole SM per
  // tile); large early-stop batches start on the throughput kernel and switch when the re-packing finds that few tiles left
  // (decided on the device: both kernels are launched for every pass and one of them returns at once).
  const char* ll_env = getenv("SRSLTE_B200_TDEC_LL"); // "0": never, "1": always (comparison runs and tests)
  int         mode   = v.ntiles <= sm_count ? SISO_LOW_LATENCY : (compact ? SISO_AUTO : SISO_THROUGHPUT);
  if (ll_env && ll_env[0] == '0') mode = SISO_THROUGHPUT;
  if (ll_env && ll_env[0] == '1') mode = SISO_LOW_LATENCY;
  B200_CUDA_TRY(cudaMemsetAsync(v.ctl, 0, TDEC_CTL_WORDS * sizeof(uint32_t), stream));

  prof_begin(0, stream);
  launch_load_natural(v, p.max_K, llr_dev, llr_offsets_dev, aligned8, stream, tiles_preloaded);
  prof_end(stream);
  g_kernel_launches += tiles_preloaded ? 1 : 2;
  for (uint32_t ps = 0; ps < max_passes; ps++) {
    prof_begin(1, stream);
    v.split_percent = split_percent(ps == 0 ? 0 : ((ps & 1) ? 1 : 2)); // the checkpoints are per-pass scratch: each kind of pass splits where it balances
    launch_siso_pass(v, (int)ps, mode, sm_count, stream);
    prof_end(stream);
    g_kernel_launches += mode == SISO_AUTO ? 2 : 1;
    if (compact && ps + 1 < max_passes) {
      prof_begin(3, stream);
      launch_compact(v, p.groups, p.ngroups, p.mask, p.pref, p.plans, p.moves, p.move_counter, p.gsrc, min_env ? 0u : 4u, mode == SISO_AUTO ? (uint32_t)sm_count : 0u, sm_count, stream);
      prof_end(stream);
      g_kernel_launches += 4;
    }
  }
  prof_begin(2, stream);
  launch_decide(v, p.max_K, out_dev, crc_ok_dev, npass_dev, nullptr, stream);
  prof_end(stream);
  g_kernel_launches++;
  if (w.h_err) B200_CUDA_TRY(cudaMemcpyAsync(w.h_err, v.err, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
  B200_CUDA_TRY(cudaGetLastError());
  return B200_SUCCESS;
}

int TdecEngine::run_device(TdecWorkspace& w,
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
                           cudaStream_t   stream)
{
  std::vector<TdecGroupSpec> g(1);
  g[0] = TdecGroupSpec{K, cb_idx, crc_kind, ncb, 0, 0, 0};
  return run_groups(w, llr_dev, g, max_passes, early_stop, out_dev, crc_ok_dev, npass_dev, stream);
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
    if (llr8) { // widen into a scratch vector first (one extra pass over 3 bytes per value, ~6 % of an 8-pass decode)
      if (pipe_io[0].reserve((size_t)ncb * nllr * sizeof(int16_t) + 1024) != B200_SUCCESS) return B200_ERROR;
      pipe_io[0].reset();
      int16_t* wide = (int16_t*)pipe_io[0].take((size_t)ncb * nllr * sizeof(int16_t));
      launch_widen_i8(reinterpret_cast<const int8_t*>(llr), wide, (size_t)ncb * nllr, stream);
      g_kernel_launches++;
      return run_device(ws, wide, ncb, (int)K, cb_idx, max_passes, crc_kind, early_stop, out, crc_ok, npass, stream);
    }
    return run_device(ws, llr, ncb, (int)K, cb_idx, max_passes, crc_kind, early_stop, out, crc_ok, npass, stream);
  }

  // Host pointers: cut the batch into chunks and ping-pong two streams so the copy of chunk c+1 overlaps the decode
  // of chunk c.  6 input bytes per info bit cross PCIe here, which is what bounds this path.
  const size_t   nb    = K / 8;
  const uint32_t chunk = ncb < 2 * kPipeChunkCb ? (ncb + 1) / 2 : kPipeChunkCb;
  for (int i = 0; i < 2; i++) {
    if (pipe_io[i].reserve(chunk * (nllr * (sizeof(int16_t) + (llr8 ? 1 : 0)) + nb + 2) + 2048) != B200_SUCCESS) {
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
    rc = run_device(pipe_ws[s], d_llr, n, (int)K, cb_idx, max_passes, crc_kind, early_stop, d_out, d_ok, d_np, st);
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

// BASELINE config 3: code blocks of several lengths in one batch, one launch per pass over all of them.
int TdecEngine::run_mixed(const int16_t*  llr,
                          uint32_t        n_groups,
                          const uint32_t* K,
                          const uint32_t* ncb,
                          uint32_t        max_passes,
                          int             crc_kind,
                          int             early_stop,
                          uint8_t*        out,
                          uint8_t*        crc_ok,
                          uint8_t*        npass,
                          uint32_t        flags,
                          cudaStream_t    stream)
{
  if (!llr || !out || !K || !ncb || max_passes < 1 || max_passes > 255 || crc_kind < 0 || crc_kind > 2) {
    return B200_ERROR_INVALID_INPUTS;
  }
  if (flags & SRSRAN_B200_FLAG_LLR_INT8) {
    B200_LOG_ERROR("the int8 container is only offered by srsran_b200_tdec_run");
    return B200_ERROR_INVALID_INPUTS;
  }
  std::vector<TdecGroupSpec> groups;
  uint64_t                   llr_off = 0, out_off = 0;
  uint32_t                   cb0     = 0;
  for (uint32_t i = 0; i < n_groups; i++) {
    const int cb_idx = cb_index_exact(K[i]);
    if (cb_idx < 0) {
      B200_LOG_ERROR("Invalid CB length %u", K[i]); // turbodecoder.c:520
      return B200_ERROR;
    }
    if (ncb[i]) groups.push_back(TdecGroupSpec{(int)K[i], cb_idx, crc_kind, ncb[i], cb0, llr_off, out_off});
    llr_off += (uint64_t)ncb[i] * (3ull * K[i] + 12ull);
    out_off += (uint64_t)ncb[i] * (K[i] / 8);
    cb0 += ncb[i];
  }
  if (groups.empty()) return B200_SUCCESS;
  B200_CUDA_TRY(cudaSetDevice(ctx->device));
  if (flags & SRSRAN_B200_FLAG_DEVICE_PTRS) {
    return run_groups(ws, llr, groups, max_passes, early_stop, out, crc_ok, npass, stream);
  }
  // host pointers: one staged copy each way (the mixed entry is the decode loop's engine, not a PCIe pipeline)
  cudaStream_t st = pipe_stream[0];
  B200_CUDA_TRY(cudaStreamSynchronize(st));
  if (pipe_io[0].reserve(llr_off * sizeof(int16_t) + out_off + 2 * (size_t)cb0 + 2048) != B200_SUCCESS) return B200_ERROR;
  pipe_io[0].reset();
  int16_t* d_llr = (int16_t*)pipe_io[0].take(llr_off * sizeof(int16_t));
  uint8_t* d_out = (uint8_t*)pipe_io[0].take(out_off);
  uint8_t* d_ok  = (uint8_t*)pipe_io[0].take(cb0);
  uint8_t* d_np  = (uint8_t*)pipe_io[0].take(cb0);
  B200_CUDA_TRY(cudaMemcpyAsync(d_llr, llr, llr_off * sizeof(int16_t), cudaMemcpyHostToDevice, st));
  int rc = run_groups(pipe_ws[0], d_llr, groups, max_passes, early_stop, d_out, d_ok, d_np, st);
  if (rc != B200_SUCCESS) return rc;
  B200_CUDA_TRY(cudaMemcpyAsync(out, d_out, out_off, cudaMemcpyDeviceToHost, st));
  if (crc_ok) B200_CUDA_TRY(cudaMemcpyAsync(crc_ok, d_ok, cb0, cudaMemcpyDeviceToHost, st));
  if (npass) B200_CUDA_TRY(cudaMemcpyAsync(npass, d_np, cb0, cudaMemcpyDeviceToHost, st));
  B200_CUDA_TRY(cudaStreamSynchronize(st));
  return B200_SUCCESS;
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

int srsran_b200_tdec_run_mixed(srsran_b200_tdec_t* h,
                               const int16_t*      llr,
                               uint32_t            n_groups,
                               const uint32_t*     K,
                               const uint32_t*     ncb,
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
  return h->eng.run_mixed(llr, n_groups, K, ncb, max_passes, crc_kind, early_stop, out, crc_ok, npass, flags, (cudaStream_t)stream);
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
  return h->eng.prof_get(ms_by_class, launches_by_class, 3);
}

int srsran_b200_tdec_profile_spans(srsran_b200_tdec_t* h, float* ms, int* cls, int max_spans)
{
  if (!h || !ms || !cls || max_spans < 0) {
    return B200_ERROR_INVALID_INPUTS;
  }
  if (cudaDeviceSynchronize() != cudaSuccess) return B200_ERROR;
  int n = 0;
  for (auto& s : h->eng.spans) {
    if (n >= max_spans) break;
    float t = 0;
    if (cudaEventElapsedTime(&t, s.a, s.b) != cudaSuccess) continue;
    ms[n]  = t;
    cls[n] = s.cls;
    n++;
  }
  return n;
}

int srsran_b200_tdec_profile_get_ex(srsran_b200_tdec_t* h, double* ms_by_class, uint64_t* launches_by_class, int nclasses)
{
  if (!h || !ms_by_class || !launches_by_class || nclasses < 1) {
    return B200_ERROR_INVALID_INPUTS;
  }
  for (int i = TdecEngine::kProfClasses; i < nclasses; i++) {
    ms_by_class[i]       = 0;
    launches_by_class[i] = 0;
  }
  return h->eng.prof_get(ms_by_class, launches_by_class, nclasses);
}

} // extern "C"
