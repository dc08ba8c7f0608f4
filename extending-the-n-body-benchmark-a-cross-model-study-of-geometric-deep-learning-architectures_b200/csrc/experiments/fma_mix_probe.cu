// Do packed-half / packed-fp32 FMAs and scalar FFMA share one pipe on sm_100a?  Per-SM warp-instruction rates of
// HFMA2, FFMA2 and FFMA alone and interleaved 1:1 on independent chains (8 chains of each kind per thread).
// One pipe with HFMA2 / FFMA2 holding it 2 clk and FFMA 1 clk predicts 2.67 warp-inst/clk/SM for a 1:1 HFMA2 + FFMA mix;
// a second (lite) pipe that takes the scalar FFMAs while the packed ones sit on the heavy pipe predicts ~4.
// nvcc -arch=sm_100a -o fma_mix_probe fma_mix_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ unsigned hfma2(unsigned a, unsigned b, unsigned c) { unsigned d; asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ unsigned imad(unsigned a, unsigned b, unsigned c) { unsigned d; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

// MODE: 0 HFMA2, 1 FFMA, 2 FFMA2, 3 HFMA2 + FFMA, 4 HFMA2 + FFMA2, 5 FFMA2 + FFMA, 6 HFMA2 + 2 FFMA, 7 HFMA2 + IMAD
template <int MODE>
__global__ void probe(float* out, long long* clk, int iters) {
  float a[8]; u64 p[8]; unsigned h[8], q[8];
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3f + i; p[i] = pk(a[i], -a[i]); h[i] = 0x3c003c00u + i; q[i] = threadIdx.x + i; }
  const u64 yy = pk(1.0001f, 1.0001f), zz = pk(0.5f, 0.25f);
  const unsigned hy = 0x3c003c01u, hz = 0x38003400u;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0 || MODE == 3 || MODE == 4 || MODE == 6 || MODE == 7) h[i] = hfma2(h[i], hy, hz);
      if (MODE == 1 || MODE == 3 || MODE == 5 || MODE == 6) a[i] = fma1(a[i], 1.0001f, 0.5f);
      if (MODE == 6) a[i] = fma1(a[i], 1.0002f, 0.25f);
      if (MODE == 2 || MODE == 4 || MODE == 5) p[i] = fma2(p[i], yy, zz);
      if (MODE == 7) q[i] = imad(q[i], 3u, 7u);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  float s = 0;
  for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); s += a[i] + lo + hi + (float)h[i] + (float)q[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

int main() {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&clk, 148 * sizeof(long long));
  const int iters = 4096;
  const char* names[] = {"HFMA2", "FFMA", "FFMA2", "HFMA2+FFMA (1:1)", "HFMA2+FFMA2 (1:1)", "FFMA2+FFMA (1:1)", "HFMA2+2 FFMA", "HFMA2+IMAD (1:1)"};
  const int per_iter[] = {8, 8, 8, 16, 16, 16, 24, 16};
  for (int mode = 0; mode < 8; ++mode)
    for (int warps = 4; warps <= 16; warps *= 2) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) probe<0><<<148, warps * 32>>>(out, clk, iters);
        if (mode == 1) probe<1><<<148, warps * 32>>>(out, clk, iters);
        if (mode == 2) probe<2><<<148, warps * 32>>>(out, clk, iters);
        if (mode == 3) probe<3><<<148, warps * 32>>>(out, clk, iters);
        if (mode == 4) probe<4><<<148, warps * 32>>>(out, clk, iters);
        if (mode == 5) probe<5><<<148, warps * 32>>>(out, clk, iters);
        if (mode == 6) probe<6><<<148, warps * 32>>>(out, clk, iters);
        if (mode == 7) probe<7><<<148, warps * 32>>>(out, clk, iters);
      }
      cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, clk, sizeof(c), cudaMemcpyDeviceToHost);
      const double inst = (double)iters * per_iter[mode] * warps;
      printf("%-20s warps/SM %2d: %8lld clk, %.2f warp-inst/clk/SM\n", names[mode], warps, c, inst / c);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
