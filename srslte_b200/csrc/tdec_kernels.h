// Launchers of the turbo-decoder kernels (tdec_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tdec_core.h"

namespace b200 {

// max_K: the longest code block of the batch (sizes the grids and the shared memory)
void launch_load_natural(const TdecView& v,
                         int             max_K,
                         const int16_t*  llr_dev,
                         const uint64_t* offsets_dev, // optional per-block int16 offsets into llr_dev
                         bool            aligned8,    // every block vector starts on an 8-byte boundary
                         cudaStream_t    stream,
                         bool            int8_tiles_done = false); // only rebuild the tiles whose format flag is raised, in int16
constexpr int SISO_THROUGHPUT = 0, SISO_LOW_LATENCY = 1, SISO_AUTO = 2;
void launch_siso_pass(const TdecView& v, int pass_idx, int mode, int sm_count, cudaStream_t stream);
int  siso_resident_tiles_per_sm();
void launch_decide(const TdecView& v,
                   int             max_K,
                   uint8_t*        out_dev,
                   uint8_t*        crc_ok_dev,
                   uint8_t*        npass_dev,
                   uint8_t*        npass_run_dev,
                   cudaStream_t    stream);
// re-packs the lanes that still run into fewer tiles (two kernels, all decisions on the device)
void launch_compact(const TdecView& v, const TileGroup* groups_dev, uint32_t ngroups, uint32_t* mask_dev, uint32_t* pref_dev,
                    GroupPlan* plans_dev, MoveRec* moves_dev, uint32_t* move_counter_dev, uint32_t* gsrc_dev, uint32_t min_gain_tiles,
                    uint32_t ll_max_tiles, int sm_count, cudaStream_t stream);

// int8 LLR container -> int16 (sign extension), n values
void launch_widen_i8(const int8_t* in, int16_t* out, size_t n, cudaStream_t stream);

} // namespace b200
