// Measures the issue/pipe rate of packed fp32 (FFMA2 / FADD2 / FMUL2) against scalar FFMA on sm_100a:
// per-SM lane-operations per clock for 1..16 warps per SM sub-partition.  nvcc -arch=sm_100a -o ffma2_probe ffma2_probe.cu
#include <cuda_runtime.h>
#include <cstdio>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float tanhf_(float a) { float d; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(d) : "f"(a)); return d; }

template <int MODE>
__global__ void probe(float* out, long long* clk, int iters) {
  float x = threadIdx.x * 1e-3f, y = 1.0001f;
  float a[8];
  u64 p[8];
  for (int i = 0; i < 8; ++i) { a[i] = x + i; p[i] = pk(x + i, x - i); }
  const u64 yy = pk(y, y), zz = pk(0.5f, 0.25f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = fma1(a[i], y, 0.5f);
      if (MODE == 1) p[i] = fma2(p[i], yy, zz);
      if (MODE == 2) p[i] = add2(p[i], zz);
      if (MODE == 3) { p[i] = fma2(p[i], yy, zz); if ((i & 3) == 0) a[i] = tanhf_(a[i]); }  // 4 FFMA2 : 1 MUFU
      if (MODE == 4) { a[i] = fma1(a[i], y, 0.5f); if ((i & 3) == 0) a[(i + 1) & 7] = tanhf_(a[(i + 1) & 7]); }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  float s = 0;
  for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); s += a[i] + lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

int main() {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&clk, 148 * sizeof(long long));
  const int iters = 4096;
  const char* names[] = {"FFMA", "FFMA2", "FADD2", "FFMA2+MUFU(4:1)", "FFMA+MUFU(4:1)"};
  for (int mode = 0; mode < 5; ++mode)
    for (int warps = 4; warps <= 32; warps *= 2) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) probe<0><<<148, warps * 32>>>(out, clk, iters);
        if (mode == 1) probe<1><<<148, warps * 32>>>(out, clk, iters);
        if (mode == 2) probe<2><<<148, warps * 32>>>(out, clk, iters);
        if (mode == 3) probe<3><<<148, warps * 32>>>(out, clk, iters);
        if (mode == 4) probe<4><<<148, warps * 32>>>(out, clk, iters);
      }
      cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, clk, sizeof(c), cudaMemcpyDeviceToHost);
      const double inst = (double)iters * 8 * warps;  // warp-instructions of the main op per SM
      const double lanes = inst * 32 * ((mode == 1 || mode == 2 || mode == 3) ? 2 : 1);
      printf("%-18s warps/SM %2d: %8lld clk, %.2f warp-inst/clk/SM, %.1f fp32 lane-ops/clk/SM\n", names[mode], warps, c,
             inst / c, lanes / c);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
