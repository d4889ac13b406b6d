// Host-side engine of the batched turbo decoder (see tdec_host.cu).
#pragma once
#include <atomic>
#include <vector>

#include "b200_runtime.h"

namespace b200 {

extern std::atomic<uint64_t> g_kernel_launches;

// Code blocks of one length and CRC kind inside a batch.  A batch is a list of groups; the engine orders their tiles by
// descending K and runs ONE launch per pass over all of them (BASELINE config 3: all 188 sizes in one batch).
struct TdecGroupSpec {
  int      K;
  int      cb_idx;
  int      crc_kind; // SRSRAN_B200_CRC_*
  uint32_t ncb;
  uint32_t cb0;      // index of the group's first block in the per-block arrays (offset list, crc_ok, npass)
  uint64_t llr_off;  // int16 offset of the group's first natural vector inside llr (contiguous inputs; unused with an offset list)
  uint64_t out_off;  // byte offset of the group's first block in out (K/8 bytes per block, blocks consecutive)
  bool operator==(const TdecGroupSpec& o) const
  {
    return K == o.K && cb_idx == o.cb_idx && crc_kind == o.crc_kind && ncb == o.ncb && cb0 == o.cb0 && llr_off == o.llr_off && out_off == o.out_off;
  }
};

// What TdecEngine::prepare leaves on the device for one batch
struct TdecPlan {
  TdecView   v{};
  int        max_K   = 0;
  uint32_t   ngroups = 0;
  TileGroup* groups  = nullptr; // device
  uint32_t*  mask    = nullptr; // re-packing scratch, per tile
  uint32_t*  pref    = nullptr;
  GroupPlan* plans   = nullptr;
  MoveRec*   moves   = nullptr; // [ntiles*32]
  uint32_t*  move_counter = nullptr;
  uint32_t*  gsrc    = nullptr; // [ntiles*32] source slot of every lane slot during a re-packing, LANE_EMPTY otherwise
};

// Decoder workspace of one stream: device arrays, the page-locked staging of the tile descriptors and the last batch
// shape prepared in it (the same shape again -- the steady state of a receiver -- reuses the descriptors on the device).
struct TdecWorkspace {
  DeviceArena                arena;
  PinnedArena                stage;
  cudaEvent_t                uploaded = nullptr; // the staged descriptors have been copied
  std::vector<TdecGroupSpec> cached;
  uint64_t                   cached_generation = ~0ull; // arena generation the cached descriptors live in
  bool                       cached_int16 = true;
  TdecPlan                   plan;
  // The int16 copies of the channel LLRs (a tile uses them when a value does not fit int8) are 40 % of the workspace and
  // almost never touched.  With int16_on_demand the workspace is carved WITHOUT them until a batch needs them: the load
  // kernels then flag the batch (TdecView::err, mirrored into *h_err after the decode) and the caller, who synchronises
  // anyway, sets have_int16 and runs the decode again.  Only for synchronous callers (the transport-block decode loop).
  bool                       int16_on_demand = false;
  bool                       have_int16      = false;
  uint32_t*                  h_err           = nullptr; // page-locked
  void                       release();
};

// Decoder workspaces are scratch: valid only while one decode runs.  Callers that synchronise at the end of a decode (the
// transport-block decode loop) borrow one from a per-device pool instead of owning one, so that many receiver objects of a
// process (one per cell) share as many workspaces as decodes are in flight, not one each (~0.9 GB per 13,000 code blocks).
TdecWorkspace* workspace_acquire(int device);
void           workspace_release(int device, TdecWorkspace* w);

struct TdecEngine {
  // code blocks per pipeline chunk on the host-pointer path: ~300 MB of LLRs at K=6144, big enough to run PCIe at
  // full rate and to fill the GPU (128 tiles), small enough that two chunks in flight stay modest
  static constexpr uint32_t kPipeChunkCb = 8192;

  DeviceContext* ctx = nullptr;
  int            sm_count = 148;
  TdecWorkspace  ws;             // device-pointer path
  cudaStream_t   pipe_stream[2] = {nullptr, nullptr};
  TdecWorkspace  pipe_ws[2];     // host-pointer path: decoder workspace per stream
  DeviceArena    pipe_io[2];     // host-pointer path: staged inputs/outputs per stream

  // Optional per-kernel-class timing with CUDA events on the launching stream (bench.py's roofline numbers).
  static constexpr int kProfClasses = 4; // 0 load, 1 siso pass, 2 decide, 3 re-packing between passes
  struct ProfSpan {
    cudaEvent_t a, b;
    int         cls;
  };
  bool                  profiling = false;
  std::vector<ProfSpan> spans;
  void prof_begin(int cls, cudaStream_t st);
  void prof_end(cudaStream_t st);
  void prof_reset(bool enable);
  int  prof_get(double* ms_by_class, uint64_t* launches_by_class, int nclasses); // synchronises the device

  static size_t workspace_bytes(const std::vector<TdecGroupSpec>& groups, bool with_int16 = true);
  static size_t workspace_bytes(int K, uint32_t ncb);
  // carve the workspace for `groups`, build and upload the tile descriptors (or reuse the cached ones)
  int prepare(TdecWorkspace& w, const std::vector<TdecGroupSpec>& groups, cudaStream_t stream);
  // first tile of every group in the order prepare() lays the tiles out (descending K, stable)
  static void tile_layout(const std::vector<TdecGroupSpec>& groups, std::vector<uint32_t>& first_tile);
  // prepare() + the per-decode flags cleared: for a producer that fills the tiles itself (the fused de-matching kernel) and
  // then calls run_groups(.., tiles_preloaded = true)
  int begin_batch(TdecWorkspace& w, const std::vector<TdecGroupSpec>& groups, cudaStream_t stream);

  int  init(int device, uint32_t max_cb_hint);
  void destroy();

  // Everything on `stream`, all pointers device memory.  offsets (optional): int16 offset of every block's vector in llr_dev.
  int run_groups(TdecWorkspace&                    w,
                 const int16_t*                    llr_dev,
                 const std::vector<TdecGroupSpec>& groups,
                 uint32_t                          max_passes,
                 int                               early_stop,
                 uint8_t*                          out_dev,
                 uint8_t*                          crc_ok_dev,
                 uint8_t*                          npass_dev,
                 cudaStream_t                      stream,
                 const uint64_t*                   llr_offsets_dev  = nullptr,
                 bool                              offsets_aligned8 = false,
                 bool                              tiles_preloaded  = false);

  // one group of equal-K blocks, contiguous vectors
  int run_device(TdecWorkspace& w,
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
                 cudaStream_t   stream);

  int run(const int16_t* llr,
          uint32_t       ncb,
          uint32_t       K,
          uint32_t       max_passes,
          int            crc_kind,
          int            early_stop,
          uint8_t*       out,
          uint8_t*       crc_ok,
          uint8_t*       npass,
          uint32_t       flags,
          cudaStream_t   stream);

  // mixed code block lengths in one batch (include/srslte_b200.h: srsran_b200_tdec_run_mixed)
  int run_mixed(const int16_t*  llr,
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
                cudaStream_t    stream);
};

} // namespace b200
