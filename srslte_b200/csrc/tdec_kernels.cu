// CUDA kernels of the batched turbo decoder (sm_100a).  The arithmetic lives in tdec_core.h; this file maps it onto
// the grid and adds the two layout-conversion kernels at the ends of a decode.
//
//   tdec_load_natural_kernel   natural [cb][3K+12] int16 (turbodecoder_gen.c:238-258 order)  ->  S/P0/P1/S2T tiles
//   tdec_siso_pass_kernel      one SISO pass (turbodecoder_iter.h:72-144) for every still-active code block
//   tdec_decide_kernel         HB (visiting order of the last pass)  ->  packed bytes, natural order
//                              (turbodecoder.c:370-378 + turbodecoder_gen.c:260-277)
#include <cuda_runtime.h>

#include "tdec_core.h"
#include "tdec_kernels.h"

namespace b200 {

// ---------------------------------------------------------------------------------------------------------------
// One warp per tile of 64 code blocks, one block per warp: with the 65,536-block benchmark batch that is 1024
// independent warps, ~7 per SM, each a long serial recursion -- throughput comes from instruction-level
// parallelism inside a trellis step (8 independent state updates) and from register prefetch of the next window.
template <bool DEC2, bool FIRST>
__global__ void __launch_bounds__(32, 1) tdec_siso_pass_kernel(TdecView v, int pass_idx)
{
  const int tile = blockIdx.x;
  const int lane = threadIdx.x;
  if (tile >= v.ntiles) return;
  // whole tile finished (early stop): nothing to do for this warp
  const CbStatus* st  = v.status + ((size_t)tile * TDEC_TILE_CB + 2 * (size_t)lane);
  const bool      any = (st[0].active | st[1].active) != 0;
  if (__ballot_sync(0xFFFFFFFFu, any) == 0u) return;
  siso_pass_lane<DEC2, FIRST, 4>(v, tile, lane, pass_idx);
}

void launch_siso_pass(const TdecView& v, int pass_idx, cudaStream_t stream)
{
  dim3 grid((unsigned)v.ntiles), block(32);
  if (pass_idx == 0) {
    tdec_siso_pass_kernel<false, true><<<grid, block, 0, stream>>>(v, pass_idx);
  } else if (pass_idx & 1) {
    tdec_siso_pass_kernel<true, false><<<grid, block, 0, stream>>>(v, pass_idx);
  } else {
    tdec_siso_pass_kernel<false, false><<<grid, block, 0, stream>>>(v, pass_idx);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// natural -> tiled.  Block = (tile, chunk of 32 trellis rows).  Phase 1 stages the 64 blocks' 96 contiguous int16
// each through shared memory with 8-byte loads (a block's natural vector is only 8-byte aligned: (3K+12)*2 bytes);
// phase 2 emits one uint4 (4 rows x 2 blocks) per thread and stream.
constexpr int LOAD_ROWS = 32;
constexpr int LOAD_PITCH = 3 * LOAD_ROWS + 4; // int16 per staged block, +4 keeps 8-byte alignment and skews banks

__global__ void __launch_bounds__(256) tdec_load_natural_kernel(TdecView v, const int16_t* __restrict__ llr, uint32_t ncb)
{
  __shared__ __align__(16) int16_t sm[TDEC_TILE_CB * LOAD_PITCH];
  const int    tile  = blockIdx.y;
  const int    chunk = blockIdx.x;
  const int    K     = v.K;
  const int    k0    = chunk * LOAD_ROWS;
  const size_t nllr  = 3 * (size_t)K + 12;
  const int    tid   = threadIdx.x;

  if (k0 < K) {
    const int rows  = min(LOAD_ROWS, K - k0); // multiple of 8
    const int nvec  = rows * 3 / 4;           // 8-byte vectors per block in this chunk
    for (int idx = tid; idx < TDEC_TILE_CB * nvec; idx += 256) {
      const int      c  = idx / nvec, q = idx % nvec;
      const uint32_t cb = (uint32_t)tile * TDEC_TILE_CB + c;
      uint2          val = make_uint2(0u, 0u);
      if (cb < ncb) {
        val = __ldcs(reinterpret_cast<const uint2*>(llr + cb * nllr + 3 * (size_t)k0) + q);
      }
      *reinterpret_cast<uint2*>(&sm[c * LOAD_PITCH + 4 * q]) = val;
    }
    __syncthreads();
    const int lane = tid & 31;
    for (int r4 = tid >> 5; r4 < rows / 4; r4 += 8) {
      uint32_t w[3][4];
#pragma unroll
      for (int t = 0; t < 4; t++) {
#pragma unroll
        for (int s = 0; s < 3; s++) {
          const int o = 3 * (4 * r4 + t) + s;
          w[s][t]     = pack2(sm[(2 * lane) * LOAD_PITCH + o], sm[(2 * lane + 1) * LOAD_PITCH + o]);
        }
      }
      const size_t row = vec_row(v, tile, k0 / 4 + r4, lane);
      v.S[row]         = u4{w[0][0], w[0][1], w[0][2], w[0][3]};
      v.P0[row]        = u4{w[1][0], w[1][1], w[1][2], w[1][3]};
      v.P1[row]        = u4{w[2][0], w[2][1], w[2][2], w[2][3]};
    }
  } else {
    // the chunk past the payload carries the 12 tail values: rows K..K+3 of S/P0/P1 and S2T
    if (tid < 32) {
      const int lane = tid;
      uint32_t  w[4][4];
#pragma unroll
      for (int s = 0; s < 4; s++) {
#pragma unroll
        for (int t = 0; t < 4; t++) {
          int16_t a = 0, b = 0;
          const uint32_t cb0 = (uint32_t)tile * TDEC_TILE_CB + 2 * lane;
          if (cb0 < ncb) a = natural_pick(llr + cb0 * nllr, K, s, K + t);
          if (cb0 + 1 < ncb) b = natural_pick(llr + (cb0 + 1) * nllr, K, s, K + t);
          w[s][t] = pack2(a, b);
        }
      }
      const size_t row = vec_row(v, tile, K / 4, lane);
      v.S[row]         = u4{w[0][0], w[0][1], w[0][2], w[0][3]};
      v.P0[row]        = u4{w[1][0], w[1][1], w[1][2], w[1][3]};
      v.P1[row]        = u4{w[2][0], w[2][1], w[2][2], w[2][3]};
      v.S2T[(size_t)tile * 32 + lane] = u4{w[3][0], w[3][1], w[3][2], w[3][3]};
      // arm the per-block state for a fresh decode (srsran_tdec_new_cb, turbodecoder.c:510-525)
      const uint32_t cb0 = (uint32_t)tile * TDEC_TILE_CB + 2 * lane;
      v.status[cb0]      = CbStatus{(uint8_t)(cb0 < ncb), 0, 0, 0};
      v.status[cb0 + 1]  = CbStatus{(uint8_t)(cb0 + 1 < ncb), 0, 0, 0};
    }
  }
}

void launch_load_natural(const TdecView& v, const int16_t* llr_dev, uint32_t ncb, cudaStream_t stream)
{
  const int chunks = (v.K + LOAD_ROWS - 1) / LOAD_ROWS + 1; // +1: the tail chunk
  dim3      grid((unsigned)chunks, (unsigned)v.ntiles), block(256);
  tdec_load_natural_kernel<<<grid, block, 0, stream>>>(v, llr_dev, ncb);
}

// ---------------------------------------------------------------------------------------------------------------
// HB -> bytes.  Block = (tile, 32 output bytes per code block); decisions are transposed through shared memory so
// that both the HB reads (64 B rows) and the byte writes (32 B runs per block) are contiguous.
__global__ void __launch_bounds__(256) tdec_decide_kernel(TdecView v,
                                                          const uint16_t* __restrict__ qpp_rev,
                                                          uint8_t* __restrict__ out,
                                                          uint8_t* __restrict__ crc_ok,
                                                          uint8_t* __restrict__ npass,
                                                          uint8_t* __restrict__ npass_run,
                                                          uint32_t ncb)
{
  __shared__ uint8_t sm[TDEC_TILE_CB][33];
  const int tile = blockIdx.y;
  const int jb0  = blockIdx.x * 32;
  const int nb   = v.K / 8;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

  for (int j = wid; j < 32; j += 8) {
    const int jb = jb0 + j;
    if (jb < nb) {
      sm[2 * lane][j]     = decide_byte(v, qpp_rev, tile * TDEC_TILE_CB + 2 * lane, jb);
      sm[2 * lane + 1][j] = decide_byte(v, qpp_rev, tile * TDEC_TILE_CB + 2 * lane + 1, jb);
    }
  }
  __syncthreads();
  for (int c = wid; c < TDEC_TILE_CB; c += 8) {
    const uint32_t cb = (uint32_t)tile * TDEC_TILE_CB + c;
    const int      jb = jb0 + lane;
    if (cb < ncb && jb < nb) {
      out[(size_t)cb * nb + jb] = sm[c][lane];
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < TDEC_TILE_CB) {
    const uint32_t cb = (uint32_t)tile * TDEC_TILE_CB + threadIdx.x;
    if (cb < ncb) {
      const CbStatus s = v.status[cb];
      if (crc_ok) crc_ok[cb] = s.crc_ok;
      // the caller's loop counter (sch.c:431-432): pass at which the CRC matched, else the passes spent
      if (npass) npass[cb] = s.crc_ok ? s.npass_crc : s.npass_run;
      if (npass_run) npass_run[cb] = s.npass_run;
    }
  }
}

void launch_decide(const TdecView& v,
                   const uint16_t* qpp_rev_dev,
                   uint8_t*        out_dev,
                   uint8_t*        crc_ok_dev,
                   uint8_t*        npass_dev,
                   uint8_t*        npass_run_dev,
                   uint32_t        ncb,
                   cudaStream_t    stream)
{
  dim3 grid((unsigned)((v.K / 8 + 31) / 32), (unsigned)v.ntiles), block(256);
  tdec_decide_kernel<<<grid, block, 0, stream>>>(v, qpp_rev_dev, out_dev, crc_ok_dev, npass_dev, npass_run_dev, ncb);
}

} // namespace b200
