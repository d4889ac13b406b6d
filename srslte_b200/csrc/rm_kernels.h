// Rate de-matching kernel interface (rm_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tdec_core.h"

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

// max_E: the longest E of the list when known (sizes the staging), 0 = unknown
int launch_rm_rx(const int16_t* e_bits_dev, int16_t* soft_pool_dev, const RmDescDev* descs_dev, uint32_t n, cudaStream_t stream, uint32_t max_E = 0);

// De-matching fused with the decoder's tile layout: one thread block per lane slot of `v` (ntiles * 32).  pairs_dev: two int32
// per slot = index into descs_dev of the slot's low / high block, -1 for none.  Every soft buffer must start on an 8-byte
// boundary.  Leaves the tiles, S2T, lane map and block state of `v` as the decoder's own load kernels would.
int launch_rm_rx_tiles(const int16_t* e_bits_dev, int16_t* soft_pool_dev, const RmDescDev* descs_dev, const void* pairs_dev,
                       const TdecView& v, int max_K, uint32_t max_E, cudaStream_t stream);

} // namespace b200
