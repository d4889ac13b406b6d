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

int launch_rm_rx(const int16_t* e_bits_dev, int16_t* soft_pool_dev, const RmDescDev* descs_dev, uint32_t n, cudaStream_t stream)
{
  if (n == 0) return B200_SUCCESS;
  static std::atomic<uint64_t> attr{0}; // function attributes are per device
  const size_t smem = RM_MAX_STAGE * sizeof(int16_t);
  if (once_per_device(attr)) {
    B200_CUDA_TRY(cudaFuncSetAttribute(rm_rx_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  rm_rx_gather_kernel<<<n, RM_THREADS, smem, stream>>>(e_bits_dev, soft_pool_dev, descs_dev, n);
  B200_CUDA_TRY(cudaGetLastError());
  return B200_SUCCESS;
}

} // namespace b200
