// Host-side engine of the batched turbo decoder (see tdec_host.cu).
#pragma once
#include <atomic>
#include <vector>

#include "b200_runtime.h"

namespace b200 {

extern std::atomic<uint64_t> g_kernel_launches;

struct TdecEngine {
  // code blocks per pipeline chunk on the host-pointer path: ~300 MB of LLRs at K=6144, big enough to run PCIe at
  // full rate and to fill the GPU (128 tiles), small enough that two chunks in flight stay modest
  static constexpr uint32_t kPipeChunkCb = 8192;

  DeviceContext* ctx = nullptr;
  DeviceArena    arena;          // device-pointer path
  cudaStream_t   pipe_stream[2] = {nullptr, nullptr};
  DeviceArena    pipe_arena[2];  // host-pointer path: decoder workspace per stream
  DeviceArena    pipe_io[2];     // host-pointer path: staged inputs/outputs per stream

  // Optional per-kernel-class timing with CUDA events on the launching stream (bench.py's roofline numbers).
  struct ProfSpan {
    cudaEvent_t a, b;
    int         cls; // 0 load, 1 siso pass, 2 decide
  };
  bool                  profiling = false;
  std::vector<ProfSpan> spans;
  void prof_begin(int cls, cudaStream_t st);
  void prof_end(cudaStream_t st);
  void prof_reset(bool enable);
  int  prof_get(double* ms_by_class, uint64_t* launches_by_class); // synchronises the device

  static size_t workspace_bytes(int K, uint32_t ncb);
  int           carve(DeviceArena& a, int K, uint32_t ncb, TdecView& v) const;

  int  init(int device, uint32_t max_cb_hint);
  void destroy();

  int run_device(DeviceArena&   ws,
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
                 const uint64_t* llr_offsets_dev = nullptr, // optional: block cb's vector starts at llr_dev + offsets[cb]
                 bool            offsets_aligned8 = false,
                 bool            reset_ws = true);

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
};

} // namespace b200
