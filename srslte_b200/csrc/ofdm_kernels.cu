// OFDM receive on the GPU: CP removal, half-subcarrier shift, forward DFT, window-offset phase fix, FFT-shift with
// guard/DC removal and optional 1/sqrt(N), fused into one kernel.
//
// Behaviour restated from the reference (it calls FFTW there):
//   srsran_ofdm_rx_sf / ofdm_rx_slot   lib/src/phy/dft/ofdm.c:387-422,453-466
//   plan geometry                      lib/src/phy/dft/ofdm.c:126-166  (window start = cp1 + l*(N+cp2) - window_offset_n)
//   shift / window-offset tables       lib/src/phy/dft/ofdm.c:130-138,334-362
//
// One thread group per OFDM symbol runs a Stockham auto-sort FFT with register radix-16/8/4/3/2 butterflies:
//   pass 1 reads the N window samples straight from the subframe buffer (float2, coalesced), multiplied on the fly by the
//          shift table (which only depends on the position inside the FFT window, so it is N entries, not 15N);
//   middle passes exchange through padded shared memory (ping-pong, one __syncthreads per pass);
//   the last pass writes straight to the output grid: bin -> resource element with the FFT-shift, guard and DC drop of
//          ofdm.c:410-411, times the per-element window-offset/normalisation factor.
// So a subframe costs 15N*8 bytes of reads (only 14N of them touched) and 14*12*nof_prb*8 bytes of writes: HBM-bound.
// Unlike the reference the caller's input buffer is NOT modified (ofdm.c:455-457 multiplies the shift in place, which
// makes its srsran_ofdm_rx_sf non-idempotent).
#include <cuda_runtime.h>
#include <stdlib.h>

#include "b200_runtime.h"
#include "ofdm_kernels.h"

namespace b200 {

__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b)
{
  return make_float2(a.x + b.x, a.y + b.y);
}
__device__ __forceinline__ float2 csub(float2 a, float2 b)
{
  return make_float2(a.x - b.x, a.y - b.y);
}
// multiply by -i (forward transform quarter turn)
__device__ __forceinline__ float2 mul_mi(float2 a)
{
  return make_float2(a.y, -a.x);
}

// int16 I/Q sample -> float2 (exact: the scale is a power of two)
__device__ __forceinline__ float2 iq16_to_f2(uint32_t w, float scale)
{
  return make_float2((float)(int16_t)(w & 0xFFFFu) * scale, (float)(int16_t)(w >> 16) * scale);
}

template <int R>
__device__ __forceinline__ void dft_small(float2* u);

template <>
__device__ __forceinline__ void dft_small<2>(float2* u)
{
  float2 a = u[0], b = u[1];
  u[0]     = cadd(a, b);
  u[1]     = csub(a, b);
}

template <>
__device__ __forceinline__ void dft_small<3>(float2* u)
{
  // X1,2 = u0 - (u1+u2)/2 -+ i*(sqrt3/2)*(u1-u2)   (e^{-2 pi i/3} = -1/2 - i sqrt3/2)
  const float s  = 0.86602540378443864676f;
  float2      t  = cadd(u[1], u[2]);
  float2      d  = csub(u[1], u[2]);
  float2      m  = make_float2(u[0].x - 0.5f * t.x, u[0].y - 0.5f * t.y);
  float2      r  = make_float2(s * d.y, -s * d.x); // -i*s*d
  u[0]           = cadd(u[0], t);
  u[1]           = cadd(m, r);
  u[2]           = csub(m, r);
}

template <>
__device__ __forceinline__ void dft_small<5>(float2* u)
{
  // Winograd-style radix 5 (transform de-precoding sizes 12 L with L = 2^a 3^b 5^c, dft_precoding.c:88-95)
  const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f; // cos(2 pi/5), cos(4 pi/5)
  const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;  // sin(2 pi/5), sin(4 pi/5)
  const float2 t1 = cadd(u[1], u[4]), t2 = cadd(u[2], u[3]), t3 = csub(u[1], u[4]), t4 = csub(u[2], u[3]);
  const float2 a1 = make_float2(u[0].x + c1 * t1.x + c2 * t2.x, u[0].y + c1 * t1.y + c2 * t2.y);
  const float2 a2 = make_float2(u[0].x + c2 * t1.x + c1 * t2.x, u[0].y + c2 * t1.y + c1 * t2.y);
  const float2 b1 = make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
  const float2 b2 = make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
  u[0]            = make_float2(u[0].x + t1.x + t2.x, u[0].y + t1.y + t2.y);
  u[1]            = cadd(a1, mul_mi(b1)); // a1 - i b1
  u[4]            = csub(a1, mul_mi(b1));
  u[2]            = cadd(a2, mul_mi(b2));
  u[3]            = csub(a2, mul_mi(b2));
}

template <>
__device__ __forceinline__ void dft_small<4>(float2* u)
{
  float2 a = cadd(u[0], u[2]), b = csub(u[0], u[2]);
  float2 c = cadd(u[1], u[3]), d = mul_mi(csub(u[1], u[3]));
  u[0]     = cadd(a, c);
  u[1]     = cadd(b, d);
  u[2]     = csub(a, c);
  u[3]     = csub(b, d);
}

template <>
__device__ __forceinline__ void dft_small<8>(float2* u)
{
  // 2 x radix-4 on even/odd, then combine with W8^k
  const float h = 0.70710678118654752440f;
  float2      e[4] = {u[0], u[2], u[4], u[6]}, o[4] = {u[1], u[3], u[5], u[7]};
  dft_small<4>(e);
  dft_small<4>(o);
  o[1] = make_float2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));   // * e^{-i pi/4}
  o[2] = mul_mi(o[2]);
  o[3] = make_float2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));  // * e^{-3i pi/4}
#pragma unroll
  for (int k = 0; k < 4; k++) {
    u[k]     = cadd(e[k], o[k]);
    u[k + 4] = csub(e[k], o[k]);
  }
}

template <>
__device__ __forceinline__ void dft_small<16>(float2* u)
{
  // 4 x 4: columns (stride 4), twiddle W16^(r*c), rows
  const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
  float2      col[4][4];
#pragma unroll
  for (int c = 0; c < 4; c++) {
#pragma unroll
    for (int r = 0; r < 4; r++) col[c][r] = u[c + 4 * r];
    dft_small<4>(col[c]);
  }
  // W16^m = (cos(m pi/8), -sin(m pi/8))
  const float2 w1 = make_float2(c1, -s1), w2 = make_float2(h, -h), w3 = make_float2(s1, -c1);
  const float2 w4 = make_float2(0.f, -1.f), w6 = make_float2(-h, -h), w9 = make_float2(-c1, s1);
  col[1][1] = cmul(col[1][1], w1);
  col[1][2] = cmul(col[1][2], w2);
  col[1][3] = cmul(col[1][3], w3);
  col[2][1] = cmul(col[2][1], w2);
  col[2][2] = cmul(col[2][2], w4);
  col[2][3] = cmul(col[2][3], w6);
  col[3][1] = cmul(col[3][1], w3);
  col[3][2] = cmul(col[3][2], w6);
  col[3][3] = cmul(col[3][3], w9);
#pragma unroll
  for (int r = 0; r < 4; r++) {
    float2 row[4] = {col[0][r], col[1][r], col[2][r], col[3][r]};
    dft_small<4>(row);
#pragma unroll
    for (int c = 0; c < 4; c++) u[r + 4 * c] = row[c];
  }
}

__device__ __forceinline__ int pad_idx(int i)
{
  return i + (i >> 4);
}

// Map DFT bin -> output resource element index, or -1 when the bin is a guard / the dropped DC (ofdm.c:410-411)
__device__ __forceinline__ int bin_to_re(int bin, int N, int R, int dc)
{
  const int half = R >> 1;
  if (bin >= N - half) return bin - (N - half);
  if (bin >= dc && bin < dc + half) return half + bin - dc;
  return -1;
}

template <int RADIX>
__device__ __forceinline__ void fft_pass(const OfdmPlanDev& p,
                                         int               Ns,
                                         bool              first,
                                         bool              last,
                                         const float2* __restrict__ gin, // window start in global memory (first pass)
                                         const float2*     sin_,          // shared source (other passes)
                                         float2*           sout,          // shared destination (all but last pass)
                                         float2* __restrict__ gout,       // symbol's output row (last pass)
                                         int               t,
                                         int               tps,
                                         const float2* __restrict__ eq_h = nullptr, // PUSCH mode: the slot's channel estimates
                                         float             eq_n0 = 0.f)
{
  const int N = p.N, T = N / RADIX;
  for (int j = t; j < T; j += tps) {
    float2 u[RADIX];
#pragma unroll
    for (int q = 0; q < RADIX; q++) {
      const int idx = j + q * T;
      if (first) {
        float2 v = p.iq16 ? iq16_to_f2(reinterpret_cast<const uint32_t*>(gin)[idx], p.iq_scale) : gin[idx];
        if (p.shift) v = cmul(v, p.shift[idx]);
        if (eq_h) { // precoding.c:224-262: (y conj(h)) / (|h|^2 [+ noise when noise > 0])
          const float2 h  = eq_h[idx];
          float        hh = h.x * h.x + h.y * h.y;
          if (eq_n0 > 0.f) hh += eq_n0;
          v = make_float2((v.x * h.x + v.y * h.y) / hh, (v.y * h.x - v.x * h.y) / hh);
        }
        if (p.inverse) v.y = -v.y;
        u[q] = v;
      } else {
        u[q] = sin_[pad_idx(idx)];
      }
    }
    const int k = j % Ns;
    if (Ns > 1) {
      const int step = N / (Ns * RADIX);
#pragma unroll
      for (int q = 1; q < RADIX; q++) u[q] = cmul(u[q], p.W[q * k * step]);
    }
    dft_small<RADIX>(u);
    const int j0 = (j / Ns) * Ns * RADIX + k;
#pragma unroll
    for (int q = 0; q < RADIX; q++) {
      const int o = j0 + q * Ns;
      if (last) {
        const int re = p.generic ? o : bin_to_re(o, N, p.R, p.dc);
        if (re >= 0) {
          float2 v = u[q];
          if (p.ramp) v = cmul(v, p.ramp[re]);
          if (p.inverse) v.y = -v.y;
          if (p.gscale != 0.f) v = make_float2(v.x * p.gscale, v.y * p.gscale);
          gout[re] = v;
        }
      } else {
        sout[pad_idx(o)] = u[q];
      }
    }
  }
}

__global__ void __launch_bounds__(OFDM_THREADS) ofdm_rx_kernel(OfdmPlanDev p, const float2* __restrict__ in, float2* __restrict__ out,
                                                               uint32_t nsf)
{
  extern __shared__ __align__(16) float2 smem[];
  const int      tps     = p.tps;                 // threads per symbol
  const int      spb     = OFDM_THREADS / tps;    // symbols per block
  const int      graw    = threadIdx.x / tps;     // symbol slot inside the block
  const bool     spare   = graw >= spb;           // threads beyond the last full group only keep the barriers company
  const int      g       = spare ? 0 : graw;
  const int      t       = threadIdx.x % tps;
  const int      padN    = p.N + (p.N >> 4) + 1;
  float2*        bufA    = smem + (size_t)g * 2 * padN;
  float2*        bufB    = bufA + padN;
  const uint32_t nsymtot = p.generic ? nsf : nsf * (uint32_t)p.nsym;
  const int      half    = p.generic ? 1 : p.nsym / 2;

  for (uint32_t base = blockIdx.x * spb; base < nsymtot; base += gridDim.x * spb) {
    const uint32_t sidx   = base + g;
    const bool     active = !spare && sidx < nsymtot;
    const uint32_t sf = (active && !p.generic) ? sidx / p.nsym : 0, l = (active && !p.generic) ? sidx % p.nsym : 0;
    const int      slot = (int)l / half, ls = (int)l % half;
    const size_t   woff = (size_t)sf * p.sf_sz + (size_t)slot * p.slot_sz + p.cp1 + (size_t)ls * (p.N + p.cp2) - p.noff;
    const float2*  gin  = p.generic ? in + (size_t)sidx * p.idist
                                    : (p.iq16 ? reinterpret_cast<const float2*>(reinterpret_cast<const uint32_t*>(in) + woff) : in + woff);
    float2*        gout = p.generic ? out + (size_t)sidx * p.odist : out + (size_t)sidx * p.R;
    const float2*  eq_h  = nullptr;
    float          eq_n0 = 0.f;
    if (p.generic == 2 && active) {
      const uint32_t psf = sidx / (uint32_t)p.pusch_nd, d = sidx % (uint32_t)p.pusch_nd;
      const int      lsym = p.pusch_l[d];
      gin   = in + ((size_t)psf * p.grid_nsym + lsym) * p.grid_R + p.grid_off;
      gout  = out + (size_t)sidx * p.N;
      eq_h  = p.eq_ce + ((size_t)psf * 2 + (lsym >= p.grid_nsym / 2 ? 1 : 0)) * p.N;
      eq_n0 = p.eq_noise ? p.eq_noise[(size_t)psf * p.eq_noise_stride] : 0.f;
    }
    int            Ns   = 1;
    float2 *       src = bufA, *dst = bufB;
    for (int ps = 0; ps < p.npass; ps++) {
      const bool first = ps == 0, last = ps == p.npass - 1;
      if (active) {
        switch (p.radix[ps]) {
          case 16:
            fft_pass<16>(p, Ns, first, last, gin, src, dst, gout, t, tps, eq_h, eq_n0);
            break;
          case 8:
            fft_pass<8>(p, Ns, first, last, gin, src, dst, gout, t, tps, eq_h, eq_n0);
            break;
          case 4:
            fft_pass<4>(p, Ns, first, last, gin, src, dst, gout, t, tps, eq_h, eq_n0);
            break;
          case 3:
            fft_pass<3>(p, Ns, first, last, gin, src, dst, gout, t, tps, eq_h, eq_n0);
            break;
          case 5:
            fft_pass<5>(p, Ns, first, last, gin, src, dst, gout, t, tps, eq_h, eq_n0);
            break;
          default:
            fft_pass<2>(p, Ns, first, last, gin, src, dst, gout, t, tps, eq_h, eq_n0);
            break;
        }
      }
      Ns *= p.radix[ps];
      __syncthreads();
      float2* tmp = src;
      src         = dst;
      dst         = tmp;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Specialised kernel for the power-of-two LTE sizes (N = 128 .. 2048): everything that the generic kernel above
// derives at run time (pass sequence, strides, index divisions) is a compile-time constant, N/16 threads hold 16 points
// each in registers through every pass, and the passes exchange IN PLACE through one padded shared buffer (all loads of
// a pass, barrier, all stores), which halves the shared memory per symbol and doubles the resident warps.
template <int N, int RADIX, int NS, bool FIRST, bool LAST, bool IQ16 = false>
__device__ __forceinline__ void pass_ct(const OfdmPlanDev& p, const float2* __restrict__ gin, float2* buf, float2* __restrict__ gout, int t,
                                        bool active)
{
  constexpr int TPS = N / 16, T = N / RADIX, ITER = 16 / RADIX;
  float2        u[ITER][RADIX];
  if (active) {
    // Table values reach the multipliers through ONE load per butterfly plus a recurrence (the tables are bigger than
    // what is left of L1 beside the shared-memory carve-out, so every table load is an L2 round trip):
    // shift[j + q*T] = shift[j] * c^q with c = shift[noff + T] (the table is a pure phasor, 1 at index noff).
    float2 c1 = make_float2(1.f, 0.f);
    if (FIRST && p.shift) c1 = p.shift[p.noff + T];
#pragma unroll
    for (int it = 0; it < ITER; it++) {
      const int j = t + it * TPS;
      float2    sh = make_float2(1.f, 0.f);
      if (FIRST && p.shift) sh = p.shift[j];
#pragma unroll
      for (int q = 0; q < RADIX; q++) {
        const int idx = j + q * T;
        if (FIRST) { // gin = this symbol's window, already staged in shared memory (natural order)
          float2 v = IQ16 ? iq16_to_f2(reinterpret_cast<const uint32_t*>(gin)[idx], p.iq_scale) : gin[idx];
          if (p.shift) {
            v  = cmul(v, sh);
            sh = cmul(sh, c1);
          }
          u[it][q] = v;
        } else {
          u[it][q] = buf[pad_idx(idx)];
        }
      }
    }
  }
  if (!FIRST) __syncthreads(); // in place: everybody has read its points before anybody overwrites them
  if (active) {
#pragma unroll
    for (int it = 0; it < ITER; it++) {
      const int j = t + it * TPS;
      const int k = j % NS;
      if (NS > 1) { // twiddles W^(q k step) = w1^q
        constexpr int step = N / (NS * RADIX);
        const float2  w1   = p.W[k * step];
        float2        w    = w1;
#pragma unroll
        for (int q = 1; q < RADIX; q++) {
          u[it][q] = cmul(u[it][q], w);
          if (q + 1 < RADIX) w = cmul(w, w1);
        }
      }
      dft_small<RADIX>(u[it]);
      const int j0 = (j / NS) * NS * RADIX + k;
      // window-offset phase ramp x normalisation of bin o = j0 + q NS:  norm * exp(+2 pi i noff o / N) = rb * d^q with
      // rb = norm * conj(W[noff j0 mod N]) and d = conj(W[noff NS mod N]): two table loads per butterfly, not RADIX
      float2 rb = make_float2(1.f, 0.f), d1 = rb;
      if (LAST && p.ramp) {
        const float2 a = p.W[(p.noff * j0) % N], b = p.W[(p.noff * NS) % N];
        rb             = make_float2(a.x * p.norm, -a.y * p.norm);
        d1             = make_float2(b.x, -b.y);
      }
#pragma unroll
      for (int q = 0; q < RADIX; q++) {
        const int o = j0 + q * NS;
        if (LAST) {
          const int re = bin_to_re(o, N, p.R, p.dc);
          if (re >= 0) {
            float2 v = u[it][q];
            if (p.ramp) v = cmul(v, rb);
            __stcs(&gout[re], v);
          }
          if (p.ramp) rb = cmul(rb, d1);
        } else {
          buf[pad_idx(o)] = u[it][q];
        }
      }
    }
  }
  if (!LAST) __syncthreads();
}

__device__ __forceinline__ void cp_async8(float2* dst_smem, const float2* src)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)),
               "l"(__cvta_generic_to_global(src))
               : "memory");
}

// Persistent blocks; the window of the NEXT symbol is fetched with cp.async into a staging buffer while the current
// symbol's passes run, so the global-memory latency is off the critical path.
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)),
               "l"(__cvta_generic_to_global(src))
               : "memory");
}

template <int N, int R0, int R1, int R2, bool IQ16>
__global__ void __launch_bounds__(OFDM_THREADS, 6) ofdm_rx_kernel_ct(OfdmPlanDev p, const float2* __restrict__ in, float2* __restrict__ out,
                                                                     uint32_t nsf)
{
  extern __shared__ __align__(16) float2 smem[];
  constexpr int  TPS  = N / 16;
  constexpr int  SPB  = OFDM_THREADS / TPS;
  constexpr int  PADN = N + (N >> 4) + 1;
  const int      g    = threadIdx.x / TPS;
  const int      t    = threadIdx.x % TPS;
  float2*        buf   = smem + (size_t)g * (PADN + N);
  float2*        stage = buf + PADN;
  const uint32_t nsymtot = nsf * (uint32_t)p.nsym;
  const int      half    = p.nsym / 2;
  auto window = [&](uint32_t sidx) {
    const uint32_t sf = sidx / p.nsym, l = sidx % p.nsym;
    const int      slot = (int)l / half, ls = (int)l % half;
    return (size_t)sf * p.sf_sz + (size_t)slot * p.slot_sz + p.cp1 + (size_t)ls * (N + p.cp2) - p.noff; // in samples
  };
  auto prefetch = [&](uint32_t sidx) {
    if (sidx < nsymtot) {
      if (IQ16) { // 4-byte samples: the staging buffer holds them as they come, the first pass converts
        const uint32_t* gin = reinterpret_cast<const uint32_t*>(in) + window(sidx);
        uint32_t*       st4 = reinterpret_cast<uint32_t*>(stage);
#pragma unroll
        for (int q = 0; q < 16; q++) cp_async4(st4 + t + q * TPS, gin + t + q * TPS);
      } else {
        const float2* gin = in + window(sidx);
#pragma unroll
        for (int q = 0; q < 16; q++) cp_async8(stage + t + q * TPS, gin + t + q * TPS);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  prefetch(blockIdx.x * SPB + g);
  for (uint32_t base = blockIdx.x * SPB; base < nsymtot; base += gridDim.x * SPB) {
    const uint32_t sidx   = base + g;
    const bool     active = sidx < nsymtot;
    float2*        gout   = out + (size_t)sidx * p.R;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads(); // the staged window is complete and visible to the whole group
    pass_ct<N, R0, 1, true, false, IQ16>(p, stage, buf, gout, t, active);
    prefetch(sidx + gridDim.x * SPB); // the barrier that ended the first pass: every thread has read its staged points
    if (R2 > 1) {
      pass_ct<N, R1, R0, false, false>(p, stage, buf, gout, t, active);
      pass_ct<N, (R2 > 1 ? R2 : 2), R0 * R1, false, true>(p, stage, buf, gout, t, active);
    } else {
      pass_ct<N, R1, R0, false, true>(p, stage, buf, gout, t, active);
    }
    __syncthreads(); // the last pass's reads of buf are done before the next symbol's first pass overwrites it
  }
}

template <int N, int R0, int R1, int R2, bool IQ16>
static int launch_ct_t(const OfdmPlanDev& p, const float2* in_dev, float2* out_dev, uint32_t nsf, int sm_count, cudaStream_t stream)
{
  constexpr int    SPB  = OFDM_THREADS / (N / 16);
  constexpr size_t smem = (size_t)SPB * (2 * N + (N >> 4) + 1) * sizeof(float2); // work buffer + staged next window
  static std::atomic<uint64_t> attr_done{0}; // function attributes are per device
  if (once_per_device(attr_done)) {
    B200_CUDA_TRY(cudaFuncSetAttribute(ofdm_rx_kernel_ct<N, R0, R1, R2, IQ16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B200_CUDA_TRY(cudaFuncSetAttribute(ofdm_rx_kernel_ct<N, R0, R1, R2, IQ16>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       cudaSharedmemCarveoutMaxShared));
  }
  const uint32_t nsymtot = nsf * (uint32_t)p.nsym;
  uint32_t       blocks  = (nsymtot + SPB - 1) / SPB;
  const uint32_t cap     = (uint32_t)sm_count * 12u; // persistent: the resident blocks loop over the symbols
  if (blocks > cap) blocks = cap;
  ofdm_rx_kernel_ct<N, R0, R1, R2, IQ16><<<blocks, OFDM_THREADS, smem, stream>>>(p, in_dev, out_dev, nsf);
  B200_CUDA_TRY(cudaGetLastError());
  return B200_SUCCESS;
}

template <int N, int R0, int R1, int R2>
static int launch_ct(const OfdmPlanDev& p, const float2* in_dev, float2* out_dev, uint32_t nsf, int sm_count, cudaStream_t stream)
{
  return p.iq16 ? launch_ct_t<N, R0, R1, R2, true>(p, in_dev, out_dev, nsf, sm_count, stream)
                : launch_ct_t<N, R0, R1, R2, false>(p, in_dev, out_dev, nsf, sm_count, stream);
}


// ---------------------------------------------------------------------------------------------------------------
// Compile-time specialised batched DFT for the generic / PUSCH de-precoding modes (sizes 12 L_prb with factors 3 and 5):
// radices, strides and index divisions are constants, the twiddles of a butterfly come from ONE table load plus a
// recurrence, a thread keeps all its points of a pass in registers (<= 16) so the passes exchange in place through one
// padded shared buffer, several symbols share a block so that no thread idles, and the inputs of a block's NEXT symbols
// (samples and, in PUSCH mode, channel estimates) are fetched with cp.async into a staging buffer while the current ones
// are transformed.
template <int N, int RADIX, int NS, int TPS, bool FIRST, bool LAST>
__device__ __forceinline__ void dft_pass_ct(const OfdmPlanDev& p, const float2* gin, const float2* eq_h, float eq_n0, float2* buf,
                                            float2* __restrict__ gout, int t, bool active)
{
  constexpr int T = N / RADIX, ITER = (T + TPS - 1) / TPS;
  static_assert(ITER * RADIX <= 16, "a thread holds at most 16 points");
  float2 u[ITER][RADIX];
  if (active) {
#pragma unroll
    for (int it = 0; it < ITER; it++) {
      const int j = t + it * TPS;
      if (ITER * TPS == T || j < T) {
#pragma unroll
        for (int q = 0; q < RADIX; q++) {
          const int idx = j + q * T;
          if (FIRST) {
            float2 v = gin[idx];
            if (eq_h) { // precoding.c:224-262: (y conj(h)) / (|h|^2 [+ noise when noise > 0]); one rounded reciprocal
              const float2 h  = eq_h[idx];
              float        hh = h.x * h.x + h.y * h.y;
              if (eq_n0 > 0.f) hh += eq_n0;
              const float r = __frcp_rn(hh);
              v = make_float2((v.x * h.x + v.y * h.y) * r, (v.y * h.x - v.x * h.y) * r);
            }
            if (p.inverse) v.y = -v.y;
            u[it][q] = v;
          } else {
            u[it][q] = buf[pad_idx(idx)];
          }
        }
      }
    }
  }
  if (!FIRST) __syncthreads(); // in place: everybody has read its points before anybody overwrites them
  if (active) {
#pragma unroll
    for (int it = 0; it < ITER; it++) {
      const int j = t + it * TPS;
      if (ITER * TPS == T || j < T) {
        const int k = j % NS;
        if (NS > 1) {
          constexpr int step = N / (NS * RADIX);
          const float2  w1   = p.W[k * step];
          float2        w    = w1;
#pragma unroll
          for (int q = 1; q < RADIX; q++) {
            u[it][q] = cmul(u[it][q], w);
            if (q + 1 < RADIX) w = cmul(w, w1);
          }
        }
        dft_small<RADIX>(u[it]);
        const int j0 = (j / NS) * NS * RADIX + k;
#pragma unroll
        for (int q = 0; q < RADIX; q++) {
          const int o = j0 + q * NS;
          if (LAST) {
            float2 v = u[it][q];
            if (p.inverse) v.y = -v.y;
            if (p.gscale != 0.f) v = make_float2(v.x * p.gscale, v.y * p.gscale);
            __stcs(&gout[o], v);
          } else {
            buf[pad_idx(o)] = u[it][q];
          }
        }
      }
    }
  }
  if (!LAST) __syncthreads();
}

constexpr int dft_ct_tps(int N, int r0, int r1, int r2, int r3)
{
  int tps = 1;
  const int r[4] = {r0, r1, r2, r3};
  for (int i = 0; i < 4; i++) {
    if (r[i] > 1) {
      const int T = N / r[i], per = 16 / r[i], need = (T + per - 1) / per;
      if (need > tps) tps = need;
    }
  }
  return tps;
}

template <int N, int R0, int R1, int R2, int R3>
struct DftCt {
  static constexpr int TPS  = dft_ct_tps(N, R0, R1, R2, R3);
  static constexpr int SPB  = (256 / TPS) > 0 ? (256 / TPS) : 1;
  static constexpr int PADN = N + (N >> 4) + 1;
};

template <int N, int R0, int R1, int R2, int R3>
__global__ void __launch_bounds__(256, 2) dft_batch_kernel_ct(OfdmPlanDev p, const float2* __restrict__ in, float2* __restrict__ out, uint32_t ntot)
{
  using C = DftCt<N, R0, R1, R2, R3>;
  extern __shared__ __align__(16) float2 smem[];
  const int  graw  = threadIdx.x / C::TPS;
  const bool spare = graw >= C::SPB; // threads beyond the last full group only keep the barriers company
  const int  g     = spare ? 0 : graw;
  const int  t     = threadIdx.x % C::TPS;
  const bool eq    = p.generic == 2;
  float2*    buf   = smem + (size_t)g * (C::PADN + 2 * N);
  float2*    sy    = buf + C::PADN; // staged samples of the next symbol, natural order
  float2*    sh    = sy + N;        // staged channel estimates (PUSCH mode)
  auto source = [&](uint32_t sidx, const float2*& y, const float2*& h) {
    if (eq) {
      const uint32_t psf = sidx / (uint32_t)p.pusch_nd, d = sidx % (uint32_t)p.pusch_nd;
      const int      lsym = p.pusch_l[d];
      y = in + ((size_t)psf * p.grid_nsym + lsym) * p.grid_R + p.grid_off;
      h = p.eq_ce + ((size_t)psf * 2 + (lsym >= p.grid_nsym / 2 ? 1 : 0)) * N;
    } else {
      y = in + (size_t)sidx * p.idist;
      h = nullptr;
    }
  };
  auto prefetch = [&](uint32_t sidx) {
    if (!spare && sidx < ntot) {
      const float2 *y, *h;
      source(sidx, y, h);
      for (int i = t; i < N; i += C::TPS) {
        cp_async8(sy + i, y + i);
        if (eq) cp_async8(sh + i, h + i);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  prefetch(blockIdx.x * C::SPB + g);
  for (uint32_t base = blockIdx.x * C::SPB; base < ntot; base += gridDim.x * C::SPB) {
    const uint32_t sidx   = base + g;
    const bool     active = !spare && sidx < ntot;
    float2*        gout   = eq ? out + (size_t)sidx * N : out + (size_t)sidx * p.odist;
    float          eq_n0  = 0.f;
    if (eq && active && p.eq_noise) eq_n0 = p.eq_noise[(size_t)(sidx / (uint32_t)p.pusch_nd) * p.eq_noise_stride];
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads(); // the staged inputs are complete and visible to the whole group
    dft_pass_ct<N, R0, 1, C::TPS, true, false>(p, sy, eq ? sh : nullptr, eq_n0, buf, gout, t, active);
    prefetch(sidx + gridDim.x * C::SPB); // the barrier that ended the first pass: every thread has read its staged points
    if (R2 > 1) {
      dft_pass_ct<N, R1, R0, C::TPS, false, false>(p, sy, sh, eq_n0, buf, gout, t, active);
      if (R3 > 1) {
        dft_pass_ct<N, (R2 > 1 ? R2 : 2), R0 * R1, C::TPS, false, false>(p, sy, sh, eq_n0, buf, gout, t, active);
        dft_pass_ct<N, (R3 > 1 ? R3 : 2), R0 * R1 * (R2 > 1 ? R2 : 1), C::TPS, false, true>(p, sy, sh, eq_n0, buf, gout, t, active);
      } else {
        dft_pass_ct<N, (R2 > 1 ? R2 : 2), R0 * R1, C::TPS, false, true>(p, sy, sh, eq_n0, buf, gout, t, active);
      }
    } else {
      dft_pass_ct<N, R1, R0, C::TPS, false, true>(p, sy, sh, eq_n0, buf, gout, t, active);
    }
  }
}

template <int N, int R0, int R1, int R2, int R3>
static int launch_dft_ct(const OfdmPlanDev& p, const float2* in_dev, float2* out_dev, uint32_t ntot, int sm_count, cudaStream_t stream)
{
  using C = DftCt<N, R0, R1, R2, R3>;
  constexpr size_t smem = (size_t)C::SPB * (C::PADN + 2 * N) * sizeof(float2); // work buffer + staged next inputs per symbol
  static std::atomic<uint64_t> attr_done{0}; // function attributes are per device
  if (once_per_device(attr_done)) {
    B200_CUDA_TRY(cudaFuncSetAttribute(dft_batch_kernel_ct<N, R0, R1, R2, R3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B200_CUDA_TRY(cudaFuncSetAttribute(dft_batch_kernel_ct<N, R0, R1, R2, R3>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       cudaSharedmemCarveoutMaxShared));
  }
  uint32_t       blocks = (ntot + C::SPB - 1) / C::SPB;
  const uint32_t cap    = (uint32_t)sm_count * 8u;
  if (blocks > cap) blocks = cap;
  dft_batch_kernel_ct<N, R0, R1, R2, R3><<<blocks, 256, smem, stream>>>(p, in_dev, out_dev, ntot);
  B200_CUDA_TRY(cudaGetLastError());
  return B200_SUCCESS;
}

// the radix sequence must be the one fft_factorise() produces (the plan carries it; checked by the caller below)
static bool plan_is(const OfdmPlanDev& p, int r0, int r1, int r2, int r3)
{
  const int r[4] = {r0, r1, r2, r3};
  int       n    = 0;
  for (int i = 0; i < 4; i++) {
    if (r[i] > 1) {
      if (p.radix[i] != r[i]) return false;
      n++;
    }
  }
  return p.npass == n;
}

int launch_ofdm_rx(const OfdmPlanDev& p, const float2* in_dev, float2* out_dev, uint32_t nsf, int sm_count, cudaStream_t stream)
{
  if (nsf == 0) return B200_SUCCESS;
  if (p.generic && !p.shift && !p.ramp && getenv("SRSLTE_B200_DFT_GENERIC") == nullptr) {
    // allocation sizes of the wide LTE carriers: 100 / 75 / 50 / 25 / 15 / 6 PRB
    if (p.N == 1200 && plan_is(p, 16, 3, 5, 5)) return launch_dft_ct<1200, 16, 3, 5, 5>(p, in_dev, out_dev, nsf, sm_count, stream);
    if (p.N == 600 && plan_is(p, 8, 3, 5, 5)) return launch_dft_ct<600, 8, 3, 5, 5>(p, in_dev, out_dev, nsf, sm_count, stream);
    if (p.N == 300 && plan_is(p, 4, 3, 5, 5)) return launch_dft_ct<300, 4, 3, 5, 5>(p, in_dev, out_dev, nsf, sm_count, stream);
    if (p.N == 180 && plan_is(p, 4, 3, 3, 5)) return launch_dft_ct<180, 4, 3, 3, 5>(p, in_dev, out_dev, nsf, sm_count, stream);
    if (p.N == 72 && plan_is(p, 8, 3, 3, 1)) return launch_dft_ct<72, 8, 3, 3, 1>(p, in_dev, out_dev, nsf, sm_count, stream);
  }
  if (!p.generic && !p.inverse) {
    switch (p.N) {
      case 2048:
        return launch_ct<2048, 16, 16, 8>(p, in_dev, out_dev, nsf, sm_count, stream);
      case 1024:
        return launch_ct<1024, 16, 16, 4>(p, in_dev, out_dev, nsf, sm_count, stream);
      case 512:
        return launch_ct<512, 16, 8, 4>(p, in_dev, out_dev, nsf, sm_count, stream);
      case 256:
        return launch_ct<256, 16, 16, 1>(p, in_dev, out_dev, nsf, sm_count, stream);
      case 128:
        return launch_ct<128, 16, 8, 1>(p, in_dev, out_dev, nsf, sm_count, stream);
      default:
        break; // 384 / 768 / 1536 / forced sizes: the generic kernel
    }
  }
  const int      spb     = OFDM_THREADS / p.tps;
  const size_t   smem    = (size_t)spb * 2 * (p.N + (p.N >> 4) + 1) * sizeof(float2);
  const uint32_t nsymtot = p.generic ? nsf : nsf * (uint32_t)p.nsym;
  static size_t  attr_set = 0;
  if (smem > attr_set) {
    B200_CUDA_TRY(cudaFuncSetAttribute(ofdm_rx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = smem;
  }
  uint32_t blocks = (nsymtot + spb - 1) / spb;
  const uint32_t cap = (uint32_t)sm_count * 8u; // persistent: a few resident blocks per SM loop over the symbols
  if (blocks > cap) blocks = cap;
  ofdm_rx_kernel<<<blocks, OFDM_THREADS, smem, stream>>>(p, in_dev, out_dev, nsf);
  B200_CUDA_TRY(cudaGetLastError());
  return B200_SUCCESS;
}

} // namespace b200
