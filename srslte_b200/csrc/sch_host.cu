// Batched shared-channel receive processing: rate de-matching and the transport-block decode loop.
//
// Replaces, a batch of transport blocks at a time, what the reference does per code block on one CPU thread:
//   decode_tb     lib/src/phy/phch/sch.c:507-572   (filler check, TB CRC24A with the non-zero-parity rule)
//   decode_tb_cb  lib/src/phy/phch/sch.c:370-492   (E split with its off-by-one, de-match into the HARQ soft buffer, up to
//                                                   max_iterations passes with a CRC check after each, cb_crc bookkeeping)
// Code blocks of all transport blocks are pooled, grouped by (K, CRC kind) and each group runs as ONE batched decode.
#include <string.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <new>
#include <vector>

#include "../../include/srslte_b200.h"
#include "b200_runtime.h"
#include "lte_tables.h"
#include "rm_kernels.h"
#include "tdec_engine.h"
#include "tdec_kernels.h"

namespace b200 {

// Copies each code block's payload bytes from the per-group decision buffer into the transport block.  The reference
// writes K/8 bytes per block at data[c * rlen/8] in block order, so every block but the last of its TB effectively
// contributes rlen/8 bytes (its 3 CRC bytes are overwritten by the next block) -- reproduced without the overlap.
struct ScatterJob {
  uint64_t dst;    // byte offset in data
  uint64_t src;    // byte offset in the decision buffer of the batch
  uint32_t nbytes;
  uint32_t pad;
};

__global__ void sch_scatter_payload_kernel(const uint8_t* __restrict__ dec, uint8_t* __restrict__ data,
                                           const ScatterJob* __restrict__ jobs, uint32_t n)
{
  const uint32_t j = blockIdx.x;
  if (j >= n) return;
  const ScatterJob  job = jobs[j];
  const uint8_t*    src = dec + job.src;
  for (uint32_t i = threadIdx.x; i < job.nbytes; i += blockDim.x) data[job.dst + i] = src[i];
}

// One thread per transport block: CRC24A over tbs bits (byte table, crc.c:30-46,147-160) against the received parity.
struct TbCrcJob {
  uint64_t data_off;
  uint32_t tbs;
  uint32_t pad;
};

// a(x) * b(x) mod g(x) over GF(2), 24-bit operands
__device__ __forceinline__ uint32_t crc24_mulmod(uint32_t a, uint32_t b)
{
  uint32_t r = 0;
#pragma unroll 1
  for (int i = 23; i >= 0; i--) {
    r = (r & 0x800000u) ? ((r << 1) ^ CRC24A_POLY) : (r << 1);
    if ((b >> i) & 1u) r ^= a;
  }
  return r & 0xFFFFFFu;
}

// Transport block CRC24A (sch.c:543-559), one warp per transport block.  The block's bytes are cut into 32 chunks of L
// bytes aligned to the END of the block (the first chunks may be short or empty, which is harmless with a zero initial
// value); lane i computes the plain CRC of chunk i and the results are combined by linearity:
//   crc(block) = sum_i crc(chunk_i) * x^(8 L (31 - i))  mod g.
__global__ void __launch_bounds__(128) sch_tb_crc_kernel(const uint8_t* __restrict__ data, const TbCrcJob* __restrict__ jobs,
                                                         uint32_t n, uint8_t* __restrict__ ok)
{
  __shared__ uint32_t table[256];
  for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
    uint32_t r = i << 16;
    for (int b = 0; b < 8; b++) {
      r = (r & 0x800000u) ? ((r << 1) ^ CRC24A_POLY) : (r << 1);
    }
    table[i] = r & 0xFFFFFFu;
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t t    = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= n) return;
  const TbCrcJob j      = jobs[t];
  if (j.pad == 0) { // not a valid transport block of this batch (rejected inputs): nothing to read
    if (lane == 0) ok[t] = 0;
    return;
  }
  const uint8_t* p      = data + j.data_off;
  const int      nbytes = (int)(j.tbs / 8);
  const int      L      = (nbytes + 31) / 32;
  const int      end    = nbytes - (31 - (int)lane) * L; // one past this lane's last byte
  uint32_t       crc    = 0;
  for (int i = max(end - L, 0); i < end; i++) {
    crc = ((crc << 8) ^ table[((crc >> 16) ^ p[i]) & 0xFFu]) & 0xFFFFFFu;
  }
  // x^(8L) mod g by L table steps from 1, then this lane's multiplier (x^(8L))^(31-lane) by square and multiply
  uint32_t pw = 1;
  for (int i = 0; i < L; i++) pw = ((pw << 8) ^ table[(pw >> 16) & 0xFFu]) & 0xFFFFFFu;
  uint32_t mult = 1, sq = pw;
  for (uint32_t e = 31u - lane; e; e >>= 1) {
    if (e & 1u) mult = crc24_mulmod(mult, sq);
    sq = crc24_mulmod(sq, sq);
  }
  uint32_t part = crc24_mulmod(crc, mult);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part ^= __shfl_xor_sync(0xFFFFFFFFu, part, o);
  if (lane == 0) {
    const uint32_t rx = ((uint32_t)p[nbytes] << 16) | ((uint32_t)p[nbytes + 1] << 8) | (uint32_t)p[nbytes + 2];
    ok[t]             = (part == rx && part != 0) ? 1 : 0; // sch.c:553: parity must match AND be non-zero
  }
}

struct CbRec {
  uint32_t tb, c, K, cb_idx, E, rlen, crc_kind;
  uint64_t in_off, soft_off;
  bool     last_of_tb;
};

// Everything decode_batch derives from the transport block list alone: segmentation, E split, de-matching descriptors,
// decoder groups, payload assembly and CRC jobs -- on the host and, uploaded, on the device.  It is kept between calls: a
// receiver in steady state hands over the same list again (same grants, offsets and HARQ flags) and then none of it is
// rebuilt or uploaded; only a list that differs in any input field is planned anew.
struct SchPlan {
  bool                          valid = false;
  std::vector<srsran_b200_tb_t> key;   // the input fields of the list this plan was built for (outputs zeroed)
  uint64_t                      e_len = 0, soft_len = 0, data_len = 0;
  uint64_t                      meta_generation = ~0ull;
  // host side
  std::vector<CbRec>            cbs;
  std::vector<int32_t>          tb_result0; // verdict known from the inputs alone (rejected lists, empty blocks), else SRSRAN_ERROR
  std::vector<uint32_t>         tb_nof_cb;
  std::vector<uint8_t>          tb_valid;
  bool                          any_skip = false;
  std::vector<TdecGroupSpec>    specs;
  std::vector<uint32_t>         slot_of;    // position of code block i in the decoder's per-block arrays
  bool                          soft_offsets_aligned8 = true;
  uint64_t                      dec_bytes = 0;
  // device side (carved from SchEngine::meta)
  RmDescDev*  d_descs = nullptr;
  uint64_t*   d_offs  = nullptr;
  ScatterJob* d_jobs  = nullptr;
  TbCrcJob*   d_cj    = nullptr;
  uint8_t *   d_dec = nullptr, *d_ok = nullptr, *d_np = nullptr, *d_tbok = nullptr;
  uint32_t    max_E   = 0;       // longest received block of the list (sizes the fused kernel's staging)
  int32_t*    d_pairs = nullptr; // fused de-matching: de-matching job of every lane slot's two blocks (-1: none), [ntiles*32][2]
  uint32_t    ntiles  = 0;
};

// the input fields only, compared one by one: the caller's structs may carry anything in their output fields and padding
static bool tb_same_inputs(const srsran_b200_tb_t& a, const srsran_b200_tb_t& b)
{
  return a.tbs == b.tbs && a.Qm == b.Qm && a.rv == b.rv && a.nof_e_bits == b.nof_e_bits && a.e_offset == b.e_offset &&
         a.soft_offset == b.soft_offset && a.data_offset == b.data_offset && a.new_data == b.new_data && a.cb_crc_mask == b.cb_crc_mask;
}

struct SchEngine {
  DeviceContext* ctx = nullptr;
  TdecEngine     tdec;
  cudaStream_t   stream = nullptr;
  DeviceArena    io;   // staged e_bits / soft pool / data when the caller hands host memory
  DeviceArena    meta; // descriptor arrays, per-group decision buffers
  PinnedArena    hmeta; // page-locked staging of the descriptor arrays (decode_batch: truly asynchronous uploads)
  uint32_t       max_iterations = 10; // SRSRAN_PDSCH_MAX_TDEC_ITERS, sch.c:35
  size_t         last_ncb = 0;        // code blocks of the previous batch (sizes the bookkeeping vectors)
  cudaStream_t   after  = nullptr;    // srsran_b200_sch_decode_after: producer stream of the next decode_batch's device inputs
  bool           have_after = false;
  cudaEvent_t    after_ev = nullptr;
  cudaEvent_t    after_ext = nullptr; // srsran_b200_sch_decode_after_event: the caller's own event
  SchPlan        plan;                // of the last transport block list (reused when the next list repeats it)
  PinnedArena                   hres; // page-locked home of the per-batch verdicts (truly asynchronous copies back)
  uint8_t *                     h_tbok = nullptr, *tmp_ok = nullptr, *tmp_np = nullptr;
  std::vector<uint32_t>         iters_scratch;
  // a batch between decode_begin and decode_finish (decode_batch is the two back to back)
  struct Pending {
    bool              active = false, all_dev = false, soft_dev = false, al8 = false, preloaded = false, same = false;
    srsran_b200_tb_t* tbs    = nullptr;
    uint32_t          n_tb   = 0;
    const int16_t*    d_e    = nullptr;
    int16_t*          d_soft = nullptr;
    uint8_t*          d_data = nullptr;
    int16_t*          soft_pool = nullptr; // the caller's buffers (host or device)
    uint8_t*          data      = nullptr;
    uint64_t          soft_len = 0, data_len = 0;
    TdecWorkspace*    ws = nullptr;
    std::chrono::steady_clock::time_point t_0, t_1, t_3, t_4, t_5;
  } pend;
  int  enqueue_decode();           // decoder passes, payload, TB CRC, copies back: everything but the wait
  int  decode_begin(const int16_t* e_bits, uint64_t e_len, int16_t* soft_pool, uint64_t soft_len, uint8_t* data, uint64_t data_len,
                    srsran_b200_tb_t* tbs, uint32_t n_tb, uint32_t flags);
  int  decode_finish();
  void drop_pending()
  {
    if (pend.ws) workspace_release(ctx->device, pend.ws);
    pend.ws     = nullptr;
    pend.active = false;
  }

  int init(int device)
  {
    if (tdec.init(device, 0) != B200_SUCCESS) return B200_ERROR;
    ctx = tdec.ctx;
    B200_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    B200_CUDA_TRY(cudaEventCreateWithFlags(&after_ev, cudaEventDisableTiming));
    return B200_SUCCESS;
  }
  void destroy()
  {
    if (ctx) cudaSetDevice(ctx->device);
    if (stream) cudaStreamDestroy(stream);
    if (after_ev) cudaEventDestroy(after_ev);
    io.release();
    meta.release();
    hmeta.release();
    hres.release();
    if (pend.ws) drop_pending();
    tdec.destroy();
  }

  template <class T>
  int upload(DeviceArena& a, const std::vector<T>& h, T** d, cudaStream_t st, PinnedArena* stage = nullptr)
  {
    *d = (T*)a.take(h.size() * sizeof(T) + 16);
    if (!*d) return B200_ERROR;
    const void* src = h.data();
    if (stage) { // through page-locked memory: the copy is a plain enqueue instead of a stream synchronisation
      void* p = stage->take(h.size() * sizeof(T) + 16);
      if (p) {
        memcpy(p, h.data(), h.size() * sizeof(T));
        src = p;
      }
    }
    B200_CUDA_TRY(cudaMemcpyAsync(*d, src, h.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    return B200_SUCCESS;
  }

  int rm_batch(const int16_t* e_bits, uint64_t e_len, int16_t* soft_pool, uint64_t soft_len, const srsran_b200_rm_cb_t* cbs,
               uint32_t n, uint32_t flags, cudaStream_t user_stream);
  int build_plan(SchPlan& p, uint64_t e_len, uint64_t soft_len, uint64_t data_len, const srsran_b200_tb_t* tbs, uint32_t n_tb,
                 cudaStream_t st);
  int decode_batch(const int16_t* e_bits, uint64_t e_len, int16_t* soft_pool, uint64_t soft_len, uint8_t* data, uint64_t data_len,
                   srsran_b200_tb_t* tbs, uint32_t n_tb, uint32_t flags);
};

static int build_rm_descs(DeviceContext* ctx, const srsran_b200_rm_cb_t* cbs, uint32_t n, uint64_t e_len, uint64_t soft_len,
                          std::vector<RmDescDev>& out)
{
  out.resize(n);
  // consecutive jobs mostly share (cb_idx, rv): remember the last table instead of taking the context's lock per job
  uint32_t        last_idx = 0xFFFFFFFFu, last_rv = 0xFFFFFFFFu;
  const uint16_t* last_inv = nullptr;
  for (uint32_t i = 0; i < n; i++) {
    const srsran_b200_rm_cb_t& c = cbs[i];
    if (c.rv > 3 || c.cb_idx >= (uint32_t)NOF_CB_SIZES) {
      // rm_turbo.c:442-444
      fprintf(stderr, "Invalid inputs rv_idx=%u, cb_idx=%u\n", c.rv, c.cb_idx);
      return B200_ERROR_INVALID_INPUTS;
    }
    const uint32_t n_out = 3 * (uint32_t)cb_size(c.cb_idx) + 12;
    if (c.in_offset + c.E > e_len || c.soft_offset + n_out > soft_len) {
      B200_LOG_ERROR("de-matching job %u exceeds its buffers", i);
      return B200_ERROR_INVALID_INPUTS;
    }
    RmDescDev& d  = out[i];
    if (c.cb_idx != last_idx || c.rv != last_rv) {
      last_inv = ctx->rm_table((int)c.cb_idx, (int)c.rv);
      last_idx = c.cb_idx;
      last_rv  = c.rv;
    }
    d.inv         = last_inv;
    d.in_offset   = c.in_offset;
    d.soft_offset = c.soft_offset;
    d.E           = c.E;
    d.n_out       = n_out;
    d.flags       = c.new_data ? 1u : 0u;
    d.pad         = 0;
    if (!d.inv) return B200_ERROR;
  }
  return B200_SUCCESS;
}

int SchEngine::rm_batch(const int16_t* e_bits, uint64_t e_len, int16_t* soft_pool, uint64_t soft_len, const srsran_b200_rm_cb_t* cbs,
                        uint32_t n, uint32_t flags, cudaStream_t user_stream)
{
  if (!e_bits || !soft_pool || (!cbs && n)) return B200_ERROR_INVALID_INPUTS;
  if (n == 0) return B200_SUCCESS;
  B200_CUDA_TRY(cudaSetDevice(ctx->device));
  std::vector<RmDescDev> descs;
  int                    rc = build_rm_descs(ctx, cbs, n, e_len, soft_len, descs);
  if (rc != B200_SUCCESS) return rc;
  const bool   dev_ptrs = (flags & SRSRAN_B200_FLAG_DEVICE_PTRS) != 0;
  cudaStream_t st       = dev_ptrs ? user_stream : stream;
  if (meta.reserve(descs.size() * sizeof(RmDescDev) + 4096) != B200_SUCCESS) return B200_ERROR;
  meta.reset();
  plan.valid = false; // the decode loop's kept descriptors lived in this arena
  RmDescDev* d_descs = nullptr;
  if (upload(meta, descs, &d_descs, st) != B200_SUCCESS) return B200_ERROR;
  if (dev_ptrs) {
    rc = launch_rm_rx(e_bits, soft_pool, d_descs, n, st);
    g_kernel_launches++;
    // the descriptor upload came from a stack vector: make sure it has been consumed before returning
    B200_CUDA_TRY(cudaStreamSynchronize(st));
    return rc;
  }
  if (io.reserve((e_len + soft_len) * sizeof(int16_t) + 4096) != B200_SUCCESS) return B200_ERROR;
  io.reset();
  int16_t* d_e    = (int16_t*)io.take(e_len * sizeof(int16_t));
  int16_t* d_soft = (int16_t*)io.take(soft_len * sizeof(int16_t));
  B200_CUDA_TRY(cudaMemcpyAsync(d_e, e_bits, e_len * sizeof(int16_t), cudaMemcpyHostToDevice, st));
  B200_CUDA_TRY(cudaMemcpyAsync(d_soft, soft_pool, soft_len * sizeof(int16_t), cudaMemcpyHostToDevice, st));
  rc = launch_rm_rx(d_e, d_soft, d_descs, n, st);
  g_kernel_launches++;
  B200_CUDA_TRY(cudaMemcpyAsync(soft_pool, d_soft, soft_len * sizeof(int16_t), cudaMemcpyDeviceToHost, st));
  B200_CUDA_TRY(cudaStreamSynchronize(st));
  return rc;
}

int SchEngine::build_plan(SchPlan& p, uint64_t e_len, uint64_t soft_len, uint64_t data_len, const srsran_b200_tb_t* tbs, uint32_t n_tb,
                          cudaStream_t st)
{
  p.valid = false;
  p.cbs.clear();
  p.cbs.reserve(last_ncb + 64);
  p.tb_result0.assign(n_tb, B200_ERROR);
  p.tb_nof_cb.assign(n_tb, 0);
  p.tb_valid.assign(n_tb, 0);
  p.any_skip = false;
  std::vector<srsran_b200_rm_cb_t> rm;
  std::vector<TbCrcJob>            crc_jobs(n_tb);
  rm.reserve(last_ncb + 64);
  // ---- segmentation and the per-code-block E split of sch.c:392-406 ----------------------------------------------------
  for (uint32_t t = 0; t < n_tb; t++) {
    const srsran_b200_tb_t& tb = tbs[t];
    CbSegm                  s;
    if (tb.Qm == 0 || cb_segmentation(tb.tbs, s) != 0) {
      p.tb_result0[t] = B200_ERROR_INVALID_INPUTS;
      continue;
    }
    if (s.tbs == 0 || s.C == 0) {
      p.tb_result0[t] = B200_SUCCESS; // sch.c:517-519
      continue;
    }
    if (s.F) {
      fprintf(stderr, "Error filler bits are not supported. Use standard TBS\n"); // sch.c:521-524
      p.tb_result0[t] = B200_ERROR_INVALID_INPUTS;
      continue;
    }
    if (s.C > 32) { // SRSRAN_MAX_CODEBLOCKS, sch.c:382-385
      p.tb_result0[t] = B200_ERROR_INVALID_INPUTS;
      continue;
    }
    if (tb.e_offset + tb.nof_e_bits > e_len || tb.soft_offset + (uint64_t)s.C * SRSRAN_B200_SOFTBUFFER_SIZE > soft_len ||
        tb.data_offset + tb.tbs / 8 + 3 + MAX_CB_LEN / 8 > data_len) {
      B200_LOG_ERROR("transport block %u exceeds its buffers", t);
      p.tb_result0[t] = B200_ERROR_INVALID_INPUTS;
      continue;
    }
    p.tb_nof_cb[t]       = s.C;
    p.tb_valid[t]        = 1;
    crc_jobs[t].data_off = tb.data_offset;
    crc_jobs[t].tbs      = tb.tbs;
    crc_jobs[t].pad      = 1; // valid
    const uint32_t Gp    = tb.nof_e_bits / tb.Qm;
    const uint32_t gamma = Gp % s.C;
    const uint32_t n_e   = tb.Qm * (Gp / s.C);
    for (uint32_t c = 0; c < s.C; c++) {
      if (tb.cb_crc_mask & (1u << c)) { // sch.c:390: already decoded in an earlier transmission
        p.any_skip = true;
        continue;
      }
      CbRec r;
      r.tb     = t;
      r.c      = c;
      r.K      = c < s.C1 ? s.K1 : s.K2;           // sch.c:392 (decoder-side assignment)
      r.cb_idx = c < s.C1 ? s.K1_idx : s.K2_idx;
      r.rlen   = s.C == 1 ? r.K : r.K - 24;        // sch.c:395
      uint32_t rp = c * n_e, n_e2 = n_e;
      if (c > s.C - gamma) {                       // sch.c:403: '>' not '>=' -- the reference's off-by-one, kept
        n_e2 = n_e + tb.Qm;
        rp   = (s.C - gamma) * n_e + (c - (s.C - gamma)) * n_e2;
      }
      r.E          = n_e2;
      r.in_off     = tb.e_offset + rp;
      r.soft_off   = tb.soft_offset + (uint64_t)c * SRSRAN_B200_SOFTBUFFER_SIZE;
      r.crc_kind   = s.C > 1 ? SRSRAN_B200_CRC24B : SRSRAN_B200_CRC24A; // sch.c:437-444
      r.last_of_tb = (c == s.C - 1);
      if (r.in_off + r.E > e_len) {
        p.tb_result0[t] = B200_ERROR_INVALID_INPUTS;
        continue;
      }
      p.cbs.push_back(r);
      srsran_b200_rm_cb_t j;
      j.cb_idx      = r.cb_idx;
      j.rv          = tb.rv;
      j.E           = r.E;
      j.new_data    = tb.new_data;
      j.in_offset   = r.in_off;
      j.soft_offset = r.soft_off;
      rm.push_back(j);
    }
  }
  const std::vector<CbRec>& cbs = p.cbs;
  last_ncb                      = cbs.size();
  p.max_E                       = 0;
  for (const CbRec& r : cbs) p.max_E = std::max(p.max_E, r.E);
  // ---- de-matching descriptors, decoder groups (K, CRC kind), payload assembly -------------------------------------------
  std::vector<RmDescDev> descs;
  int                    rc = build_rm_descs(ctx, rm.data(), (uint32_t)rm.size(), e_len, soft_len, descs);
  if (rc != B200_SUCCESS) return rc;
  std::map<std::pair<uint32_t, uint32_t>, std::vector<uint32_t>> groups;
  for (uint32_t i = 0; i < cbs.size(); i++) groups[{cbs[i].K, cbs[i].crc_kind}].push_back(i);
  p.specs.clear();
  p.slot_of.assign(cbs.size(), 0);
  std::vector<uint64_t>   offs(cbs.size());
  std::vector<ScatterJob> jobs(cbs.size());
  p.soft_offsets_aligned8 = true;
  p.dec_bytes             = 0;
  {
    uint32_t slot = 0;
    for (auto& g : groups) {
      const uint32_t K = g.first.first;
      p.specs.push_back(TdecGroupSpec{(int)K, cb_index_exact(K), (int)g.first.second, (uint32_t)g.second.size(), slot, 0, p.dec_bytes});
      for (uint32_t i : g.second) {
        const CbRec& r          = cbs[i];
        p.slot_of[i]            = slot;
        offs[slot]              = r.soft_off;
        p.soft_offsets_aligned8 = p.soft_offsets_aligned8 && (r.soft_off % 4 == 0);
        jobs[slot].dst          = tbs[r.tb].data_offset + (uint64_t)r.c * (r.rlen / 8);
        jobs[slot].src          = p.dec_bytes;
        jobs[slot].nbytes       = r.last_of_tb ? r.K / 8 : r.rlen / 8;
        jobs[slot].pad          = 0;
        p.dec_bytes += K / 8;
        slot++;
      }
    }
  }
  // ---- upload ------------------------------------------------------------------------------------------------------------
  size_t meta_need = descs.size() * sizeof(RmDescDev) + cbs.size() * (sizeof(uint64_t) + sizeof(ScatterJob) + 8) +
                     n_tb * (sizeof(TbCrcJob) + 8) + p.dec_bytes + (cbs.size() + 64 * groups.size()) * 2 * sizeof(int32_t) + (size_t(2) << 20);
  if (meta.reserve(meta_need) != B200_SUCCESS) return B200_ERROR;
  meta.reset();
  // the previous call ended with a stream synchronisation, so its staged descriptors are free again
  if (hmeta.reserve(descs.size() * sizeof(RmDescDev) + cbs.size() * (sizeof(uint64_t) + sizeof(ScatterJob)) + n_tb * sizeof(TbCrcJob) +
                    (cbs.size() + 64 * groups.size()) * 2 * sizeof(int32_t) + (size_t(1) << 20)) != B200_SUCCESS) {
    return B200_ERROR;
  }
  hmeta.reset();
  p.d_descs = nullptr;
  p.d_offs  = nullptr;
  p.d_jobs  = nullptr;
  p.d_dec = p.d_ok = p.d_np = nullptr;
  if (!descs.empty() && upload(meta, descs, &p.d_descs, st, &hmeta) != B200_SUCCESS) return B200_ERROR;
  if (!cbs.empty()) {
    if (upload(meta, offs, &p.d_offs, st, &hmeta) != B200_SUCCESS || upload(meta, jobs, &p.d_jobs, st, &hmeta) != B200_SUCCESS) return B200_ERROR;
    p.d_dec = (uint8_t*)meta.take(p.dec_bytes);
    p.d_ok  = (uint8_t*)meta.take(cbs.size());
    p.d_np  = (uint8_t*)meta.take(cbs.size());
    if (!p.d_dec || !p.d_ok || !p.d_np) return B200_ERROR;
  }
  {
    // which de-matching job fills which lane slot of the decoder: code block i sits at position slot_of[i] - cb0 of its group
    std::vector<uint32_t> first_tile;
    TdecEngine::tile_layout(p.specs, first_tile);
    p.ntiles = 0;
    for (const TdecGroupSpec& g : p.specs) p.ntiles += (g.ncb + TDEC_TILE_CB - 1) / TDEC_TILE_CB;
    std::vector<int32_t> pairs((size_t)p.ntiles * 64, -1);
    uint32_t             gi = 0;
    for (auto& g : groups) {
      const TdecGroupSpec& sp = p.specs[gi];
      for (uint32_t i : g.second) {
        const uint32_t idx = p.slot_of[i] - sp.cb0;
        pairs[((size_t)first_tile[gi] + idx / TDEC_TILE_CB) * 64 + idx % TDEC_TILE_CB] = (int32_t)i; // [tile][lane][half] = [tile][block]
      }
      gi++;
    }
    p.d_pairs = nullptr;
    if (!pairs.empty() && upload(meta, pairs, &p.d_pairs, st, &hmeta) != B200_SUCCESS) return B200_ERROR;
  }
  if (upload(meta, crc_jobs, &p.d_cj, st, &hmeta) != B200_SUCCESS) return B200_ERROR;
  p.d_tbok = (uint8_t*)meta.take(n_tb);
  if (!p.d_tbok) return B200_ERROR;
  p.key.assign(tbs, tbs + n_tb);
  p.e_len           = e_len;
  p.soft_len        = soft_len;
  p.data_len        = data_len;
  p.meta_generation = meta.generation;
  p.valid           = true;
  return B200_SUCCESS;
}

int SchEngine::decode_batch(const int16_t* e_bits, uint64_t e_len, int16_t* soft_pool, uint64_t soft_len, uint8_t* data,
                            uint64_t data_len, srsran_b200_tb_t* tbs, uint32_t n_tb, uint32_t flags)
{
  const int rc = decode_begin(e_bits, e_len, soft_pool, soft_len, data, data_len, tbs, n_tb, flags);
  if (rc != B200_SUCCESS) return rc;
  return decode_finish();
}

// Everything of a batch except the wait: plan (reused when the list repeats), staging, de-matching, decoder launches, payload,
// transport-block CRC and the copies back.  Returns with the work queued on the engine's stream.
int SchEngine::decode_begin(const int16_t* e_bits, uint64_t e_len, int16_t* soft_pool, uint64_t soft_len, uint8_t* data,
                            uint64_t data_len, srsran_b200_tb_t* tbs, uint32_t n_tb, uint32_t flags)
{
  if (!e_bits || !soft_pool || !data || (!tbs && n_tb)) return B200_ERROR_INVALID_INPUTS;
  if (pend.active) {
    B200_LOG_ERROR("srsran_b200_sch_decode_begin: the previous batch has not been finished");
    return B200_ERROR_INVALID_INPUTS;
  }
  if (n_tb == 0) return B200_SUCCESS;
  B200_CUDA_TRY(cudaSetDevice(ctx->device));
  const bool   all_dev  = (flags & SRSRAN_B200_FLAG_DEVICE_PTRS) != 0;
  const bool   soft_dev = all_dev || (flags & SRSRAN_B200_FLAG_SOFT_ON_DEVICE) != 0;
  cudaStream_t st       = stream;
  if (have_after) { // the device inputs are produced on another stream: order this batch after what is queued there now
    B200_CUDA_TRY(cudaEventRecord(after_ev, after));
    B200_CUDA_TRY(cudaStreamWaitEvent(st, after_ev, 0));
    have_after = false;
  }
  if (after_ext) {
    B200_CUDA_TRY(cudaStreamWaitEvent(st, after_ext, 0));
    after_ext = nullptr;
  }

  static const bool timing = getenv("SRSLTE_B200_SCH_TIMING") != nullptr;
  auto              now    = [] { return std::chrono::steady_clock::now(); };
  pend                     = Pending();
  pend.t_0                 = now();
  // ---- the plan of this list: taken over from the previous call when nothing it depends on has changed -------------------
  SchPlan& p    = plan;
  bool     same = p.valid && p.key.size() == n_tb && p.e_len == e_len && p.soft_len == soft_len && p.data_len == data_len &&
              p.meta_generation == meta.generation;
  for (uint32_t t = 0; same && t < n_tb; t++) same = tb_same_inputs(tbs[t], p.key[t]);
  if (!same) {
    if (timing) {
      fprintf(stderr, "[sch timing] new plan: valid %d, size %zu/%u, lens %d%d%d, arena generation %llu/%llu\n", (int)p.valid, p.key.size(), n_tb,
              (int)(p.e_len == e_len), (int)(p.soft_len == soft_len), (int)(p.data_len == data_len), (unsigned long long)p.meta_generation,
              (unsigned long long)meta.generation);
    }
    int rc = build_plan(p, e_len, soft_len, data_len, tbs, n_tb, st);
    if (rc != B200_SUCCESS) {
      p.valid = false;
      return rc;
    }
  }
  const std::vector<CbRec>& cbs = p.cbs;
  for (uint32_t t = 0; t < n_tb; t++) {
    tbs[t].result         = p.tb_result0[t];
    tbs[t].nof_cb         = p.tb_nof_cb[t];
    tbs[t].avg_iterations = 0;
  }
  pend.t_1 = now();

  // ---- stage buffers ---------------------------------------------------------------------------------------------------
  const int16_t* d_e    = e_bits;
  int16_t*       d_soft = soft_pool;
  uint8_t*       d_data = data;
  if (!all_dev) {
    size_t need = e_len * sizeof(int16_t) + data_len + 8192 + (soft_dev ? 0 : soft_len * sizeof(int16_t));
    if (io.reserve(need) != B200_SUCCESS) return B200_ERROR;
    io.reset();
    int16_t* de = (int16_t*)io.take(e_len * sizeof(int16_t));
    d_data      = (uint8_t*)io.take(data_len);
    B200_CUDA_TRY(cudaMemcpyAsync(de, e_bits, e_len * sizeof(int16_t), cudaMemcpyHostToDevice, st));
    d_e = de;
    // the whole staged buffer travels back at the end: start from the caller's bytes when blocks decoded earlier must be
    // kept (sch.c:466-471), else from zeros, so that gaps, slack and rejected transport blocks never receive stale memory
    if (p.any_skip) B200_CUDA_TRY(cudaMemcpyAsync(d_data, data, data_len, cudaMemcpyHostToDevice, st));
    else B200_CUDA_TRY(cudaMemsetAsync(d_data, 0, data_len, st));
    if (!soft_dev) {
      d_soft = (int16_t*)io.take(soft_len * sizeof(int16_t));
      B200_CUDA_TRY(cudaMemcpyAsync(d_soft, soft_pool, soft_len * sizeof(int16_t), cudaMemcpyHostToDevice, st));
    }
  }
  // the verdicts come back into page-locked memory, so that their copies are plain enqueues
  if (hres.reserve((size_t)n_tb + 2 * cbs.size() + 256) != B200_SUCCESS) return B200_ERROR;
  hres.reset();
  h_tbok = (uint8_t*)hres.take(n_tb);
  tmp_ok = (uint8_t*)hres.take(cbs.size() + 1);
  tmp_np = (uint8_t*)hres.take(cbs.size() + 1);
  if (!h_tbok || !tmp_ok || !tmp_np) return B200_ERROR;
  memset(h_tbok, 0, n_tb);
  memset(tmp_ok, 0, cbs.size());
  memset(tmp_np, 0, cbs.size());

  // The decoder workspace is borrowed until decode_finish; it is carved without the int16 copies of the channel LLRs until a
  // batch needs them (TdecWorkspace::int16_on_demand).
  pend.ws = workspace_acquire(ctx->device);
  if (!pend.ws) return B200_ERROR;
  pend.ws->int16_on_demand = true;
  pend.all_dev   = all_dev;
  pend.soft_dev  = soft_dev;
  pend.same      = same;
  pend.tbs       = tbs;
  pend.n_tb      = n_tb;
  pend.d_e       = d_e;
  pend.d_soft    = d_soft;
  pend.d_data    = d_data;
  pend.soft_pool = soft_pool;
  pend.data      = data;
  pend.soft_len  = soft_len;
  pend.data_len  = data_len;

  // ---- rate de-matching of every pending code block -------------------------------------------------------------------
  // Two forms.  Default: de-match into the natural soft buffers, then the decoder's load kernel turns them into tiles (it owns a
  // whole tile and writes full 512-byte rows).  SRSLTE_B200_RM_FUSED=1: the de-matching kernel writes the decoder's int8 tiles
  // itself (one thread block per lane slot) -- bit-identical, one kernel and one read of every soft buffer less, but measured
  // SLOWER (1.40 ms against 0.67 + 0.59 ms for 53,248 blocks, profiles/README.md): a lane slot owns only 16 bytes of every tile
  // row, so its 2,300 row pieces are scattered partial-sector stores.  Kept for the record and for the tests.
  pend.al8              = p.soft_offsets_aligned8 && (reinterpret_cast<uintptr_t>(d_soft) & 7u) == 0;
  const char* fused_env = getenv("SRSLTE_B200_RM_FUSED");
  const bool  fused     = pend.al8 && !cbs.empty() && p.d_pairs != nullptr && fused_env != nullptr && fused_env[0] == '1';
  int         rc        = B200_SUCCESS;
  if (fused) {
    rc = tdec.begin_batch(*pend.ws, p.specs, st);
    if (rc == B200_SUCCESS && launch_rm_rx_tiles(d_e, d_soft, p.d_descs, p.d_pairs, pend.ws->plan.v, pend.ws->plan.max_K, p.max_E, st) != B200_SUCCESS) {
      rc = B200_ERROR;
    }
    g_kernel_launches++;
  } else if (!cbs.empty()) {
    if (launch_rm_rx(d_e, d_soft, p.d_descs, (uint32_t)cbs.size(), st, p.max_E) != B200_SUCCESS) rc = B200_ERROR;
    g_kernel_launches++;
  }
  pend.t_3       = now();
  pend.preloaded = fused;
  if (rc == B200_SUCCESS) rc = enqueue_decode();
  if (rc != B200_SUCCESS) {
    cudaStreamSynchronize(st);
    drop_pending();
    return rc;
  }
  pend.active = true;
  return B200_SUCCESS;
}

// ONE batched decode over all (K, CRC kind) groups (one launch per pass, tiles ordered by length), payload assembly, transport
// block CRC and the copies back
int SchEngine::enqueue_decode()
{
  SchPlan&                  p   = plan;
  const std::vector<CbRec>& cbs = p.cbs;
  cudaStream_t              st  = stream;
  auto                      now = [] { return std::chrono::steady_clock::now(); };
  if (!cbs.empty()) {
    int r = tdec.run_groups(*pend.ws, pend.d_soft, p.specs, max_iterations, 1, p.d_dec, p.d_ok, p.d_np, st, p.d_offs, pend.al8, pend.preloaded);
    if (r != B200_SUCCESS) return r;
    sch_scatter_payload_kernel<<<(unsigned)cbs.size(), 128, 0, st>>>(p.d_dec, pend.d_data, p.d_jobs, (uint32_t)cbs.size());
    g_kernel_launches++;
  }
  pend.t_4 = now();
  sch_tb_crc_kernel<<<(pend.n_tb + 3) / 4, 128, 0, st>>>(pend.d_data, p.d_cj, pend.n_tb, p.d_tbok);
  g_kernel_launches++;
  B200_CUDA_TRY(cudaMemcpyAsync(h_tbok, p.d_tbok, pend.n_tb, cudaMemcpyDeviceToHost, st));
  if (!cbs.empty()) {
    B200_CUDA_TRY(cudaMemcpyAsync(tmp_ok, p.d_ok, cbs.size(), cudaMemcpyDeviceToHost, st));
    B200_CUDA_TRY(cudaMemcpyAsync(tmp_np, p.d_np, cbs.size(), cudaMemcpyDeviceToHost, st));
  }
  if (!pend.all_dev) {
    B200_CUDA_TRY(cudaMemcpyAsync(pend.data, pend.d_data, pend.data_len, cudaMemcpyDeviceToHost, st));
    if (!pend.soft_dev) B200_CUDA_TRY(cudaMemcpyAsync(pend.soft_pool, pend.d_soft, pend.soft_len * sizeof(int16_t), cudaMemcpyDeviceToHost, st));
  }
  pend.t_5 = now();
  return B200_SUCCESS;
}

// Waits for the batch queued by decode_begin and writes the results into its transport-block list
int SchEngine::decode_finish()
{
  if (!pend.active) return B200_SUCCESS; // an empty list, or nothing begun
  static const bool timing = getenv("SRSLTE_B200_SCH_TIMING") != nullptr;
  auto              now    = [] { return std::chrono::steady_clock::now(); };
  cudaSetDevice(ctx->device);
  cudaStream_t st = stream;
  int          rc = B200_SUCCESS;
  if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) rc = B200_ERROR;
  if (rc == B200_SUCCESS && pend.ws->h_err && *pend.ws->h_err) {
    // some soft values did not fit the int8 tiles and this decoder workspace was carved without the int16 copies
    // (int16_on_demand): carve them from now on and decode the batch again -- the soft buffers still hold its input
    pend.ws->have_int16 = true;
    *pend.ws->h_err     = 0;
    pend.preloaded      = false; // the workspace is carved anew: the tiles are loaded again from the soft buffers
    rc                  = enqueue_decode();
    if (rc == B200_SUCCESS && (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess)) rc = B200_ERROR;
  }
  if (rc != B200_SUCCESS) {
    drop_pending();
    return rc;
  }
  auto t_6 = now();
  if (timing) {
    auto us = [](auto a, auto b) { return (double)std::chrono::duration_cast<std::chrono::nanoseconds>(b - a).count() / 1e3; };
    fprintf(stderr, "[sch timing] plan %s %.0f us, staging + de-matching launch %.0f, decode launches %.0f, tail launches %.0f, "
                    "wait for the device %.0f\n", pend.same ? "reused" : "built", us(pend.t_0, pend.t_1), us(pend.t_1, pend.t_3),
            us(pend.t_3, pend.t_4), us(pend.t_4, pend.t_5), us(pend.t_5, t_6));
  }
  SchPlan&                  p    = plan;
  const std::vector<CbRec>& cbs  = p.cbs;
  srsran_b200_tb_t*         tbs  = pend.tbs;
  const uint32_t            n_tb = pend.n_tb;
  iters_scratch.assign(n_tb, 0);
  for (size_t i = 0; i < cbs.size(); i++) {
    srsran_b200_tb_t& tb = tbs[cbs[i].tb];
    iters_scratch[cbs[i].tb] += tmp_np[p.slot_of[i]];
    if (tmp_ok[p.slot_of[i]]) tb.cb_crc_mask |= (1u << cbs[i].c);
  }
  for (uint32_t t = 0; t < n_tb; t++) {
    srsran_b200_tb_t& tb = tbs[t];
    if (!p.tb_valid[t] || tb.nof_cb == 0) continue;
    if (tb.result == B200_ERROR_INVALID_INPUTS) continue;
    tb.avg_iterations  = (float)iters_scratch[t] / (float)tb.nof_cb; // sch.c:490
    const uint32_t all = tb.nof_cb >= 32 ? 0xFFFFFFFFu : ((1u << tb.nof_cb) - 1u);
    tb.result          = ((tb.cb_crc_mask & all) == all && h_tbok[t]) ? B200_SUCCESS : B200_ERROR;
  }
  // the list's own cb_crc_mask fields have just changed: the kept plan stays valid only for a list that repeats the INPUT
  // masks (a retransmission with updated masks is planned anew, as it must be)
  drop_pending();
  return B200_SUCCESS;
}

} // namespace b200

using namespace b200;

struct srsran_b200_sch {
  SchEngine eng;
};

extern "C" {

int srsran_b200_sch_init(srsran_b200_sch_t** q, int device)
{
  if (!q) return B200_ERROR_INVALID_INPUTS;
  *q                   = nullptr;
  srsran_b200_sch_t* h = new (std::nothrow) srsran_b200_sch_t();
  if (!h) return B200_ERROR;
  if (h->eng.init(device) != B200_SUCCESS) {
    h->eng.destroy();
    delete h;
    return B200_ERROR;
  }
  *q = h;
  return B200_SUCCESS;
}

void srsran_b200_sch_free(srsran_b200_sch_t* q)
{
  if (q) {
    q->eng.destroy();
    delete q;
  }
}

void srsran_b200_sch_decode_after(srsran_b200_sch_t* q, void* producer_stream)
{
  if (q) {
    q->eng.after      = (cudaStream_t)producer_stream;
    q->eng.have_after = true;
  }
}

int srsran_b200_sch_decode_begin(srsran_b200_sch_t* q, const int16_t* e_bits, uint64_t e_len, int16_t* soft_pool, uint64_t soft_len, uint8_t* data,
                                 uint64_t data_len, srsran_b200_tb_t* tbs, uint32_t n_tb, uint32_t flags)
{
  if (!q) return B200_ERROR_INVALID_INPUTS;
  if (!(flags & SRSRAN_B200_FLAG_DEVICE_PTRS)) {
    B200_LOG_ERROR("srsran_b200_sch_decode_begin works on device buffers (host buffers would be read and written behind the caller's back)");
    return B200_ERROR_INVALID_INPUTS;
  }
  return q->eng.decode_begin(e_bits, e_len, soft_pool, soft_len, data, data_len, tbs, n_tb, flags);
}

int srsran_b200_sch_decode_finish(srsran_b200_sch_t* q)
{
  if (!q) return B200_ERROR_INVALID_INPUTS;
  return q->eng.decode_finish();
}

void srsran_b200_sch_decode_after_event(srsran_b200_sch_t* q, void* event)
{
  if (q) q->eng.after_ext = (cudaEvent_t)event;
}

void srsran_b200_sch_set_max_noi(srsran_b200_sch_t* q, uint32_t max_iterations)
{
  if (q) {
    q->eng.max_iterations = max_iterations ? max_iterations : 10; // sch.c:222-229
  }
}

int srsran_b200_rm_turbo_rx_batch(srsran_b200_sch_t*         q,
                                  const int16_t*             e_bits,
                                  uint64_t                   e_len,
                                  int16_t*                   soft_pool,
                                  uint64_t                   soft_len,
                                  const srsran_b200_rm_cb_t* cbs,
                                  uint32_t                   n,
                                  uint32_t                   flags,
                                  void*                      stream)
{
  if (!q) return B200_ERROR_INVALID_INPUTS;
  return q->eng.rm_batch(e_bits, e_len, soft_pool, soft_len, cbs, n, flags, (cudaStream_t)stream);
}

int srsran_b200_sch_decode_batch(srsran_b200_sch_t* q,
                                 const int16_t*     e_bits,
                                 uint64_t           e_len,
                                 int16_t*           soft_pool,
                                 uint64_t           soft_len,
                                 uint8_t*           data,
                                 uint64_t           data_len,
                                 srsran_b200_tb_t*  tbs,
                                 uint32_t           n_tb,
                                 uint32_t           flags)
{
  if (!q) return B200_ERROR_INVALID_INPUTS;
  return q->eng.decode_batch(e_bits, e_len, soft_pool, soft_len, data, data_len, tbs, n_tb, flags);
}

} // extern "C"
