// Small host-side runtime shared by every subsystem of the library: error reporting in the reference's style
// (int return codes + a message on stderr, lib/include/srsran/phy/utils/debug.h:77-92; no exceptions cross the C
// ABI), a grow-only device arena, and the per-device context that owns the integer tables.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <map>
#include <mutex>
#include <vector>

#include "tdec_core.h"

#define B200_SUCCESS 0
#define B200_ERROR -1
#define B200_ERROR_INVALID_INPUTS -2

#define B200_LOG_ERROR(...)                                                                                            \
  do {                                                                                                                 \
    fprintf(stderr, "[srslte_b200] %s:%d: ", __FILE__, __LINE__);                                                      \
    fprintf(stderr, __VA_ARGS__);                                                                                      \
    fprintf(stderr, "\n");                                                                                             \
  } while (0)

#define B200_CUDA_TRY(expr)                                                                                            \
  do {                                                                                                                 \
    cudaError_t e__ = (expr);                                                                                          \
    if (e__ != cudaSuccess) {                                                                                          \
      B200_LOG_ERROR("CUDA error %s (%s) in %s", cudaGetErrorName(e__), cudaGetErrorString(e__), #expr);               \
      return B200_ERROR;                                                                                               \
    }                                                                                                                  \
  } while (0)

namespace b200 {

// cudaFuncSetAttribute applies to the CURRENT device: true the first time this is called for (flag word, current device),
// so that an object created on a second GPU of the same process sets its kernels' attributes there as well.
inline bool once_per_device(std::atomic<uint64_t>& done)
{
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t bit = 1ull << (dev & 63);
  return (done.fetch_or(bit) & bit) == 0;
}

// Grow-only device buffer: steady-state calls never touch cudaMalloc.
struct DeviceArena {
  void*    base = nullptr;
  size_t   cap  = 0;
  size_t   used = 0;
  uint64_t generation = 0; // bumped whenever the memory is (re)allocated: whatever was cached inside is gone

  int  reserve(size_t bytes);
  void reset() { used = 0; }
  // 256-byte aligned carve-out; nullptr if the arena was not reserved large enough
  void* take(size_t bytes);
  void  release();
};

// Grow-only PINNED host buffer for descriptor uploads: cudaMemcpyAsync from pageable memory first synchronises the stream,
// which serialises the host-side launches with the device; from page-locked memory it is a plain enqueue.  reset() once the
// stream that consumed the previous contents has been synchronised.
struct PinnedArena {
  void*  base = nullptr;
  size_t cap  = 0;
  size_t used = 0;

  int   reserve(size_t bytes);
  void  reset() { used = 0; }
  void* take(size_t bytes); // 64-byte aligned; nullptr when not reserved large enough
  void  release();
};

struct RmTableKey {
  int  cb_idx, rv;
  bool operator<(const RmTableKey& o) const { return cb_idx != o.cb_idx ? cb_idx < o.cb_idx : rv < o.rv; }
};

// One per (process, device).  Owns read-only tables; thread-safe creation, lock-free use afterwards.
struct DeviceContext {
  int device = 0;

  // QPP tables for all 188 sizes in one allocation; offsets in uint16 units, each 16-byte aligned
  uint16_t*           qpp_fwd_all = nullptr;
  uint16_t*           qpp_rev_all = nullptr;
  std::vector<size_t> qpp_off;

  // CRC syndrome weights in visiting order (TdecView::crc_nat / crc_perm), built per (size, polynomial) on first use
  struct CrcTables {
    CrcPow* nat  = nullptr;
    CrcPow* perm = nullptr;
  };
  std::mutex                        crc_mutex;
  std::map<RmTableKey, CrcTables>   crc_tables; // key = (cb_idx, kind) with kind 0 = CRC24A, 1 = CRC24B
  int crc_visit(int cb_idx, int kind, const CrcPow** nat, const CrcPow** perm);

  std::mutex                      rm_mutex;
  std::map<RmTableKey, uint16_t*> rm_gather; // device copies of the gather-form de-matching tables

  int init(int dev);
  const uint16_t* qpp_fwd(int cb_idx) const { return qpp_fwd_all + qpp_off[cb_idx]; }
  const uint16_t* qpp_rev(int cb_idx) const { return qpp_rev_all + qpp_off[cb_idx]; }
  // device pointer to inv[d] for (cb_idx, rv), built on first use
  const uint16_t* rm_table(int cb_idx, int rv);
};

DeviceContext* device_context(int device); // nullptr on failure

} // namespace b200
