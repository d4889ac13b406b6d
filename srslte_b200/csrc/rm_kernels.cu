// Turbo rate de-matching on the GPU (gather form).
//
// Reference behaviour restated: srsran_rm_turbo_rx_lut_(input, output, in_len, cb_idx, rv, enable_input_tdec=false)
//   lib/src/phy/fec/turbo/rm_turbo.c:403-445 :  for i < in_len: output[deinter[i % out_len]] += input[i]   (int16 wrap)
// with deinter = the 188 x 4 tables of rm_turbo.c:175-248 (sub-block interleaver, bit collection, k0(rv), dummy bits).
//
// The reference scatters (one extract + add per received value, rm_turbo.c:702-705).  Here every OUTPUT element is owned
// by one thread: d -> i0 = inv[d] (inverse table, lte_tables.h) and the values i0, i0+N, i0+2N, ... < E that repetition
// folds onto it are summed.  No atomics, no write conflicts, the soft buffer is read and written once, coalesced; the
// received values of one code block are staged in shared memory so the scattered reads never leave the SM.
// int16 addition wraps and is associative, so the result is bit-identical to the reference's order of accumulation.
#include <cuda_runtime.h>

#include <algorithm>

#include "../../include/srslte_b200.h"
#include "b200_runtime.h"
#include "lte_tables.h"
#include "rm_kernels.h"

namespace b200 {

constexpr int RM_THREADS   = 256;
constexpr int RM_MAX_STAGE = 3 * MAX_CB_LEN + 12; // one wrap of the circular buffer, int16

__global__ void __launch_bounds__(RM_THREADS) rm_rx_gather_kernel(const int16_t* __restrict__ e_bits,
                                                                  int16_t* __restrict__ soft_pool,
                                                                  const RmDescDev* __restrict__ descs,
                                                                  uint32_t n)
{
  extern __shared__ __align__(16) int16_t stage[];
  const uint32_t b = blockIdx.x;
  if (b >= n) return;
  const RmDescDev  d   = descs[b];
  const uint32_t   N   = d.n_out;
  const int16_t*   in  = e_bits + d.in_offset;
  int16_t*         out = soft_pool + d.soft_offset;
  const uint16_t*  inv = d.inv;
  const bool       fresh = (d.flags & 1u) != 0;

  // accumulators for the outputs this thread owns live in the soft buffer itself: first wrap initialises, later
  // wraps (E > N, repetition) add on top.
  for (uint32_t base = 0; base < d.E || base == 0; base += N) {
    const uint32_t len = (d.E > base) ? min(N, d.E - base) : 0u;
    __syncthreads();
    // stage this wrap's received values; 'in' is only 2-byte aligned in general (rp offsets of sch.c:399-405)
    if ((reinterpret_cast<uintptr_t>(in + base) & 3u) == 0) {
      const uint32_t* in2 = reinterpret_cast<const uint32_t*>(in + base);
      uint32_t*       st2 = reinterpret_cast<uint32_t*>(stage);
      for (uint32_t i = threadIdx.x; i < len / 2; i += RM_THREADS) st2[i] = __ldcs(&in2[i]);
      if ((len & 1u) && threadIdx.x == 0) stage[len - 1] = in[base + len - 1];
    } else {
      for (uint32_t i = threadIdx.x; i < len; i += RM_THREADS) stage[i] = in[base + i];
    }
    __syncthreads();
    // two outputs per thread and step: n_out = 3K+12 is even and every soft buffer starts on a 4-byte boundary when its
    // offset is even (the decode loop's 18600-value slots are), so pairs move as aligned 32-bit words
    if (((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(inv)) & 7u) == 0 && (N & 3u) == 0) {
      // four outputs per thread and step (n_out = 3K+12 is a multiple of 4, the decode loop's soft-buffer slots start on 8-byte
      // boundaries): 8-byte table loads and 8-byte stores, half the instructions of the pair form below
      const uint2* inv4 = reinterpret_cast<const uint2*>(inv);
      uint2*       out4 = reinterpret_cast<uint2*>(out);
      for (uint32_t o4 = threadIdx.x; o4 < N / 4; o4 += RM_THREADS) {
        const uint2    ii = inv4[o4];
        const uint32_t i0 = ii.x & 0xFFFFu, i1 = ii.x >> 16, i2 = ii.y & 0xFFFFu, i3 = ii.y >> 16;
        uint2          w  = make_uint2(0u, 0u);
        if (base == 0) {
          if (!fresh) w = out4[o4];
        } else {
          if (i0 >= len && i1 >= len && i2 >= len && i3 >= len) continue;
          w = out4[o4];
        }
        int a = (int)(int16_t)(w.x & 0xFFFFu), b2 = (int)(int16_t)(w.x >> 16), c2 = (int)(int16_t)(w.y & 0xFFFFu), d2 = (int)(int16_t)(w.y >> 16);
        if (i0 < len) a += (int)stage[i0];
        if (i1 < len) b2 += (int)stage[i1];
        if (i2 < len) c2 += (int)stage[i2];
        if (i3 < len) d2 += (int)stage[i3];
        out4[o4] = make_uint2(((uint32_t)a & 0xFFFFu) | ((uint32_t)b2 << 16), ((uint32_t)c2 & 0xFFFFu) | ((uint32_t)d2 << 16));
      }
    } else if (((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(inv)) & 3u) == 0) {
      const uint32_t* inv2 = reinterpret_cast<const uint32_t*>(inv);
      uint32_t*       out2 = reinterpret_cast<uint32_t*>(out);
      for (uint32_t o2 = threadIdx.x; o2 < N / 2; o2 += RM_THREADS) {
        const uint32_t ii = inv2[o2];
        const uint32_t ia = ii & 0xFFFFu, ib = ii >> 16;
        uint32_t       w  = 0;
        if (base == 0) {
          if (!fresh) w = out2[o2];
        } else {
          if (ia >= len && ib >= len) continue;
          w = out2[o2];
        }
        int a = (int)(int16_t)(w & 0xFFFFu), b2 = (int)(int16_t)(w >> 16);
        if (ia < len) a += (int)stage[ia];
        if (ib < len) b2 += (int)stage[ib];
        out2[o2] = ((uint32_t)a & 0xFFFFu) | ((uint32_t)b2 << 16);
      }
    } else {
    for (uint32_t o = threadIdx.x; o < N; o += RM_THREADS) {
      const uint32_t i0 = inv[o];
      int            acc;
      if (base == 0) {
        acc = fresh ? 0 : (int)out[o];
      } else {
        if (i0 >= len) continue;
        acc = (int)out[o];
      }
      if (i0 < len) acc += (int)stage[i0];
      out[o] = (int16_t)(uint16_t)(unsigned)acc;
    }
    }
    if (len < N) break;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// De-matching FUSED with the decoder's input layout (rm_turbo.c:260-273,411-422 is the reference's version of the same
// idea: its de-interleaver tables can emit the layout its SIMD decoders read).  One thread block per LANE SLOT of the
// decoder (tile, lane) = two code blocks: each is de-matched exactly as above -- the combined soft values still go to the
// natural HARQ soft buffer, which later retransmissions need -- and the final values are also kept as int8 in shared
// memory, from which the block pair's 16 bytes of every S8 / P08 / P18 tile row, its tail row, the S2T entry, the lane map
// and the per-block state are written.  The separate load kernel (one more read of every soft buffer, 37 KB per block) is
// gone; a value that does not fit int8 raises the tile's format flag and only such tiles are re-read (tdec_load16_kernel).
struct RmPairDesc {
  int32_t desc[2]; // index of the de-matching job of the slot's low / high block, -1: no block
};

__global__ void __launch_bounds__(2 * RM_THREADS) rm_rx_tile_kernel(const int16_t* __restrict__ e_bits, int16_t* __restrict__ soft_pool,
                                                                const RmDescDev* __restrict__ descs, const RmPairDesc* __restrict__ pairs,
                                                                TdecView v, uint32_t stage_len)
{
  extern __shared__ __align__(16) int16_t stage[];
  const uint32_t  slot = blockIdx.x, tile = slot >> 5, lane = slot & 31u;
  const TileDesc& td   = v.tiles[tile];
  const uint32_t  K = td.K, N = 3u * K + 12u;
  const uint32_t  Np = (N + 7u) & ~7u; // each block's int8 vector starts on an 8-byte boundary
  // the two blocks of the slot are de-matched side by side, each by one half of the thread block with its own staging area
  int8_t*         nat8   = reinterpret_cast<int8_t*>(stage + 2 * stage_len); // [2][Np]; stage_len (multiple of 8) >= min(E, N) of every block
  const int       half   = (int)(threadIdx.x / RM_THREADS);
  const uint32_t  tid    = threadIdx.x % RM_THREADS;
  int16_t*        mystage = stage + (size_t)half * stage_len;
  auto half_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + half), "n"(RM_THREADS) : "memory"); };
  __shared__ int16_t tail16[2][12];
  __shared__ int     overflow;
  const RmPairDesc pr = pairs[slot];
  if (threadIdx.x == 0) overflow = 0;
  if (threadIdx.x < 24) tail16[threadIdx.x / 12][threadIdx.x % 12] = 0;
  int bad = 0;
  __syncthreads(); // the shared flags are initialised
  {
    int8_t* n8 = nat8 + (size_t)half * Np;
    if (pr.desc[half] < 0) {
      for (uint32_t i = tid; i < N / 4; i += RM_THREADS) reinterpret_cast<uint32_t*>(n8)[i] = 0u;
    } else {
    const RmDescDev d     = descs[pr.desc[half]];
    const int16_t*  in    = e_bits + d.in_offset;
    int16_t*        out   = soft_pool + d.soft_offset;
    const uint2*    inv4  = reinterpret_cast<const uint2*>(d.inv);
    uint2*          out4  = reinterpret_cast<uint2*>(out);
    const bool      fresh = (d.flags & 1u) != 0;
    for (uint32_t base = 0; base < d.E || base == 0; base += N) {
      const uint32_t len   = (d.E > base) ? min(N, d.E - base) : 0u;
      const bool     final = base + N >= d.E; // the values of this round are the block's soft values
      half_sync();
      if ((reinterpret_cast<uintptr_t>(in + base) & 3u) == 0) {
        const uint32_t* in2 = reinterpret_cast<const uint32_t*>(in + base);
        uint32_t*       st2 = reinterpret_cast<uint32_t*>(mystage);
        uint32_t i = tid;
        for (; i + 3 * RM_THREADS < len / 2; i += 4 * RM_THREADS) { // four loads in flight per thread
          const uint32_t x0 = __ldcs(&in2[i]), x1 = __ldcs(&in2[i + RM_THREADS]), x2 = __ldcs(&in2[i + 2 * RM_THREADS]),
                         x3 = __ldcs(&in2[i + 3 * RM_THREADS]);
          st2[i]                  = x0;
          st2[i + RM_THREADS]     = x1;
          st2[i + 2 * RM_THREADS] = x2;
          st2[i + 3 * RM_THREADS] = x3;
        }
        for (; i < len / 2; i += RM_THREADS) st2[i] = __ldcs(&in2[i]);
        if ((len & 1u) && tid == 0) mystage[len - 1] = in[base + len - 1];
      } else {
        for (uint32_t i = tid; i < len; i += RM_THREADS) mystage[i] = in[base + i];
      }
      half_sync();
      // the table entries of this thread's next quad are requested while the current one is worked on
      uint2 ii_next = tid < N / 4 ? __ldg(&inv4[tid]) : make_uint2(0u, 0u);
      for (uint32_t o4 = tid; o4 < N / 4; o4 += RM_THREADS) {
        const uint2    ii = ii_next;
        if (o4 + RM_THREADS < N / 4) ii_next = __ldg(&inv4[o4 + RM_THREADS]);
        const uint32_t i0 = ii.x & 0xFFFFu, i1 = ii.x >> 16, i2 = ii.y & 0xFFFFu, i3 = ii.y >> 16;
        const bool     touched = i0 < len || i1 < len || i2 < len || i3 < len;
        uint2          w  = make_uint2(0u, 0u);
        if (base == 0) {
          if (!fresh) w = out4[o4];
        } else {
          if (!touched && !final) continue;
          w = out4[o4];
        }
        int a = (int)(int16_t)(w.x & 0xFFFFu), b2 = (int)(int16_t)(w.x >> 16), c2 = (int)(int16_t)(w.y & 0xFFFFu), d2 = (int)(int16_t)(w.y >> 16);
        if (i0 < len) a += (int)mystage[i0];
        if (i1 < len) b2 += (int)mystage[i1];
        if (i2 < len) c2 += (int)mystage[i2];
        if (i3 < len) d2 += (int)mystage[i3];
        if (base == 0 || touched) {
          out4[o4] = make_uint2(((uint32_t)a & 0xFFFFu) | ((uint32_t)b2 << 16), ((uint32_t)c2 & 0xFFFFu) | ((uint32_t)d2 << 16));
        }
        if (final) {
          const int16_t vals[4] = {(int16_t)a, (int16_t)b2, (int16_t)c2, (int16_t)d2};
          uint32_t      packed  = 0;
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const uint32_t idx = 4u * o4 + (uint32_t)j;
            packed |= ((uint32_t)(uint8_t)(int8_t)vals[j]) << (8 * j);
            bool s2t = false; // encoder 2's systematic tail (natural index 3K+6+2t) stays int16 in S2T
            if (idx >= 3u * K) {
              tail16[half][idx - 3u * K] = vals[j];
              const uint32_t t = idx - 3u * K;
              s2t              = t >= 6u && ((t - 6u) & 1u) == 0u;
            }
            if (!s2t && (int16_t)(int8_t)vals[j] != vals[j]) bad = 1;
          }
          reinterpret_cast<uint32_t*>(n8)[o4] = packed;
        }
      }
      if (len < N) break;
    }
    }
  }
  if (bad) overflow = 1;
  __syncthreads();
  if (overflow && threadIdx.x == 0) {
    atomicOr(v.fmt + tile, 1u);
    if (td.S == nullptr) atomicOr(v.err, 1u);
  }
  // the pair's 16 bytes of every tile row: 8 trellis steps x 2 blocks per stream.  A window's 24 natural values of a block
  // (index 3k+s, k = 8w..8w+7) are 24 consecutive bytes: three 8-byte reads per block, then byte permutes
  const uint32_t nw = K / 8u;
  for (uint32_t w8 = threadIdx.x; w8 < nw; w8 += 2 * RM_THREADS) {
    uint32_t a[6], b[6];
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const uint2 x = *reinterpret_cast<const uint2*>(nat8 + 24u * w8 + 8u * j);
      const uint2 y = *reinterpret_cast<const uint2*>(nat8 + Np + 24u * w8 + 8u * j);
      a[2 * j] = x.x; a[2 * j + 1] = x.y; b[2 * j] = y.x; b[2 * j + 1] = y.y;
    }
#pragma unroll
    for (int s_ = 0; s_ < 3; s_++) {
      uint32_t wd[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int      i0 = 3 * (2 * q) + s_, i1 = i0 + 3; // byte index of steps 2q, 2q+1 of this stream inside the 24 bytes
        const uint32_t p0 = __byte_perm(a[i0 >> 2], b[i0 >> 2], (i0 & 3) | ((4 + (i0 & 3)) << 4));                 // (a_i0, b_i0) in bytes 0,1
        const uint32_t p1 = __byte_perm(a[i1 >> 2], b[i1 >> 2], (i1 & 3) | ((4 + (i1 & 3)) << 4));                 // (a_i1, b_i1) in bytes 0,1
        wd[q]             = __byte_perm(p0, p1, 0x5410);
      }
      u4* dst = s_ == 0 ? td.S8 : (s_ == 1 ? td.P08 : td.P18);
      dst[row8((int)w8, (int)lane)] = u4{wd[0], wd[1], wd[2], wd[3]};
    }
  }
  if (threadIdx.x == 0) {
    // tail row (trellis steps K, K+1, K+2; the fourth entry is zero), S2T, the lane map and the per-block state of a fresh decode
    uint32_t w16[4][4];
#pragma unroll
    for (int s_ = 0; s_ < 4; s_++) {
#pragma unroll
      for (int t = 0; t < 4; t++) {
        int16_t a0 = 0, b0 = 0;
        if (t < 3) {
          const int o = s_ == 0 ? 2 * t : (s_ == 1 ? 2 * t + 1 : (s_ == 2 ? 6 + 2 * t + 1 : 6 + 2 * t));
          a0          = tail16[0][o];
          b0          = tail16[1][o];
        }
        w16[s_][t] = pack2(a0, b0);
      }
    }
#pragma unroll
    for (int s_ = 0; s_ < 3; s_++) {
      uint32_t wd[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int t = 0; t < 4; t++) {
        const uint32_t lo = w16[s_][t] & 0xFFu, hi = (w16[s_][t] >> 16) & 0xFFu;
        wd[t >> 1] |= (lo | (hi << 8)) << (16 * (t & 1));
      }
      u4* dst = s_ == 0 ? td.S8 : (s_ == 1 ? td.P08 : td.P18);
      dst[row8((int)nw, (int)lane)] = u4{wd[0], wd[1], wd[2], wd[3]};
    }
    v.S2T[(size_t)tile * 32 + lane] = u4{w16[3][0], w16[3][1], w16[3][2], w16[3][3]};
    const LaneMap home = lane_home(td, (int)tile, (int)lane);
    v.lanes[(size_t)tile * 32 + lane] = home;
    v.status[home.st0]     = CbStatus{(uint8_t)(pr.desc[0] >= 0), 0, 0, 0};
    v.status[home.st0 + 1] = CbStatus{(uint8_t)(pr.desc[1] >= 0), 0, 0, 0};
  }
}

int launch_rm_rx_tiles(const int16_t* e_bits_dev, int16_t* soft_pool_dev, const RmDescDev* descs_dev, const void* pairs_dev,
                       const TdecView& v, int max_K, uint32_t max_E, cudaStream_t stream)
{
  static std::atomic<uint64_t> attr{0}; // function attributes are per device
  const size_t N         = 3 * (size_t)max_K + 12;
  const size_t stage_len = (std::min<size_t>(max_E, N) + 7) / 8 * 8; // received values staged per round: no more than the longest E
  const size_t smem      = 2 * stage_len * sizeof(int16_t) + 2 * ((N + 7) / 8 * 8);
  if (once_per_device(attr)) {
    B200_CUDA_TRY(cudaFuncSetAttribute(rm_rx_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(6 * (3 * MAX_CB_LEN + 12) + 64)));
  }
  rm_rx_tile_kernel<<<(unsigned)v.ntiles * 32u, 2 * RM_THREADS, smem, stream>>>(e_bits_dev, soft_pool_dev, descs_dev,
                                                                            reinterpret_cast<const RmPairDesc*>(pairs_dev), v, (uint32_t)stage_len);
  B200_CUDA_TRY(cudaGetLastError());
  return B200_SUCCESS;
}

int launch_rm_rx(const int16_t* e_bits_dev, int16_t* soft_pool_dev, const RmDescDev* descs_dev, uint32_t n, cudaStream_t stream, uint32_t max_E)
{
  if (n == 0) return B200_SUCCESS;
  static std::atomic<uint64_t> attr{0}; // function attributes are per device
  // a wrap stages min(n_out, E - base) values: when the caller knows the longest E of the list (punctured blocks: E is a third of
  // the circular buffer at 64QAM rate 0.87) the stage shrinks and eight blocks fit an SM instead of six
  size_t stage_len = RM_MAX_STAGE;
  if (max_E && max_E < (uint32_t)RM_MAX_STAGE) stage_len = ((size_t)max_E + 7) / 8 * 8;
  const size_t smem = stage_len * sizeof(int16_t);
  if (once_per_device(attr)) {
    B200_CUDA_TRY(cudaFuncSetAttribute(rm_rx_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(RM_MAX_STAGE * sizeof(int16_t))));
  }
  rm_rx_gather_kernel<<<n, RM_THREADS, smem, stream>>>(e_bits_dev, soft_pool_dev, descs_dev, n);
  B200_CUDA_TRY(cudaGetLastError());
  return B200_SUCCESS;
}

} // namespace b200
