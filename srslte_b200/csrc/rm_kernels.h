// Rate de-matching kernel interface (rm_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

// One code block's de-matching job, device-side view
struct RmDescDev {
  const uint16_t* inv;        // gather table for (cb_idx, rv): inv[d] = index of the received value landing on output d
  uint64_t        in_offset;  // int16 offset of this block's E received LLRs inside e_bits
  uint64_t        soft_offset;// int16 offset of this block's soft buffer (3K+12 values) inside soft_pool
  uint32_t        E;          // received LLRs (in_len of srsran_rm_turbo_rx_lut)
  uint32_t        n_out;      // 3K+12
  uint32_t        flags;      // bit0: soft buffer is logically zero (first transmission): do not read it
  uint32_t        pad;
};

int launch_rm_rx(const int16_t* e_bits_dev, int16_t* soft_pool_dev, const RmDescDev* descs_dev, uint32_t n, cudaStream_t stream);

} // namespace b200
