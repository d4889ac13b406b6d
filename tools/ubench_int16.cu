// Micro-benchmark: issue rate of the packed-int16 integer instructions the turbo decoder is built from.
// Prints warp-instructions per clock per SM for dependent-free streams of each opcode, so the decoder's
// "integer roofline"  has a measured denominator.  Build: nvcc -arch=sm_100a -O3 -o ubench tools/ubench_int16.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ADD2(a, b) asm volatile("add.s16x2 %0, %1, %2;" : "=r"(a) : "r"(a), "r"(b))
#define MAX2(a, b) asm volatile("max.s16x2 %0, %1, %2;" : "=r"(a) : "r"(a), "r"(b))
#define ADDMAX2(a, b, c) asm volatile("{.reg .b32 t; add.s16x2 t, %1, %2; max.s16x2 %0, t, %3;}" : "=r"(a) : "r"(a), "r"(b), "r"(c))
#define IADD(a, b) asm volatile("add.s32 %0, %1, %2;" : "=r"(a) : "r"(a), "r"(b))
#define IMAD(a, b, c) asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(a) : "r"(a), "r"(b), "r"(c))
#define LOP(a, b) asm volatile("xor.b32 %0, %1, %2;" : "=r"(a) : "r"(a), "r"(b))

constexpr int ILP = 8, ITER = 4096;

template <int MODE>
__global__ void k(uint32_t* out, uint32_t seed, long long* cycles)
{
  uint32_t r[ILP], b = seed | 1u, c = seed ^ 0x12345u;
#pragma unroll
  for (int i = 0; i < ILP; i++) r[i] = threadIdx.x * 7 + i + seed;
  long long t0 = clock64();
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      if (MODE == 0) ADD2(r[i], b);
      if (MODE == 1) MAX2(r[i], b);
      if (MODE == 2) ADDMAX2(r[i], b, c);
      if (MODE == 3) IADD(r[i], b);
      if (MODE == 4) IMAD(r[i], b, c);
      if (MODE == 5) LOP(r[i], b);
      if (MODE == 6) { if (i & 1) ADDMAX2(r[i], b, c); else IMAD(r[i], b, c); }   // alu + fma pipes interleaved
      if (MODE == 7) { if (i & 1) ADDMAX2(r[i], b, c); else ADD2(r[i], b); }
      if (MODE == 8) { if ((i & 3) == 3) IMAD(r[i], b, c); else ADDMAX2(r[i], b, c); }
    }
  }
  long long t1 = clock64();
  uint32_t  s  = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s ^= r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps_per_sm)
{
  int nsm = 0;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  uint32_t*  out;
  long long* cyc;
  int        threads = warps_per_sm * 32;
  cudaMalloc(&out, (size_t)nsm * threads * 4);
  cudaMalloc(&cyc, nsm * sizeof(long long));
  k<MODE><<<nsm, threads>>>(out, 3, cyc);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<nsm, threads>>>(out, 5, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[256];
  cudaMemcpy(h, cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < nsm; i++) avg += (double)h[i];
  avg /= nsm;
  double winstr = (double)ITER * ILP * warps_per_sm;
  printf("%-28s warps/SM=%2d  warp-instr/clk/SM=%.3f  (%.1f lanes/clk/SM)  kernel %.3f ms -> %.2f T lane-ops/s chip\n",
         name, warps_per_sm, winstr / avg, 32.0 * winstr / avg, ms, 32.0 * winstr * nsm / (ms * 1e-3) / 1e12);
  cudaFree(out);
  cudaFree(cyc);
}

int main()
{
  for (int w : {4, 8, 16, 32}) {
    run<0>("VIADD.16x2 (add.s16x2)", w);
    run<1>("VIMNMX.S16x2 (max.s16x2)", w);
    run<2>("VIADDMNMX.S16x2 (add+max)", w);
    run<3>("IADD3 (add.s32)", w);
    run<4>("IMAD (mad.lo.s32)", w);
    run<5>("LOP3 (xor)", w);
    run<6>("VIADDMNMX + IMAD 1:1", w);
    run<7>("VIADDMNMX + VIADD 1:1", w);
    run<8>("VIADDMNMX + IMAD 3:1", w);
  }
  return 0;
}
