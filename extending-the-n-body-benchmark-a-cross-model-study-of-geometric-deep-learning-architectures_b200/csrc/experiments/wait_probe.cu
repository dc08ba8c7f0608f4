// Latency of a SATISFIED mbarrier wait (the K3 roles spend 250-350 clk per wait in the timeline trace) in its variants,
// plus the per-SM rates of HFMA2 and tanh.approx.f16x2 (candidates for a half2 producer).
// nvcc -arch=sm_100a -o wait_probe wait_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void wait_probe(long long* out, int iters) {
  __shared__ unsigned long long bar;
  __shared__ volatile unsigned flag;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    flag = 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");  // phase 0 complete
  __syncthreads();
  const unsigned a = smem_u32(&bar);
  unsigned acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    unsigned done = 0;
    if (MODE == 0)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(a), "r"(0u) : "memory");
    if (MODE == 1)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(a), "r"(0u) : "memory");
    if (MODE == 2)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(a), "r"(0u) : "memory");
    if (MODE == 3) done = flag;
    if (MODE == 4)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(a), "r"(0u) : "memory");
    acc += done;
    if (!done) break;  // dependent branch, like the real wait loop
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = acc; }
}

template <int MODE>
__global__ void rate_probe(float* out, long long* clk, int iters) {
  __half2 h[8];
  float f[8];
  for (int i = 0; i < 8; ++i) { h[i] = __floats2half2_rn(threadIdx.x * 1e-3f + i, 0.5f - i); f[i] = threadIdx.x * 1e-3f + i; }
  const __half2 y = __floats2half2_rn(0.999f, 1.001f), z = __floats2half2_rn(0.01f, -0.01f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) h[i] = __hfma2(h[i], y, z);
      if (MODE == 1) { unsigned r, a = *reinterpret_cast<unsigned*>(&h[i]); asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(a)); *reinterpret_cast<unsigned*>(&h[i]) = r; }
      if (MODE == 2) { asm volatile("tanh.approx.f32 %0, %1;" : "=f"(f[i]) : "f"(f[i])); }
      if (MODE == 3) { h[i] = __hfma2(h[i], y, z); if ((i & 3) == 0) { unsigned r, a = *reinterpret_cast<unsigned*>(&h[i]); asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(a)); *reinterpret_cast<unsigned*>(&h[i]) = r; } }
      if (MODE == 4) {  // f32 pair -> f16x2 -> tanh -> two f32 (the epilogue variant)
        unsigned p, r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(f[i]), "f"(f[(i + 1) & 7]));
        asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(p));
        __half2 hr = *reinterpret_cast<__half2*>(&r);
        f[i] = __low2float(hr) + 0.25f; f[(i + 1) & 7] = __high2float(hr) * 0.5f;
      }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += __low2float(h[i]) + __high2float(h[i]) + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

int main() {
  long long* out; float* fo;
  cudaMalloc(&out, 148 * 2 * sizeof(long long)); cudaMalloc(&fo, 148 * 1024 * sizeof(float));
  long long h[2];
  const int iters = 1000;
  const char* names[] = {"try_wait+hint", "try_wait", "test_wait", "ld.volatile.shared", "try_wait.acquire.cta"};
  for (int mode = 0; mode < 5; ++mode)
    for (int threads = 32; threads <= 512; threads *= 4) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) wait_probe<0><<<1, threads>>>(out, iters);
        if (mode == 1) wait_probe<1><<<1, threads>>>(out, iters);
        if (mode == 2) wait_probe<2><<<1, threads>>>(out, iters);
        if (mode == 3) wait_probe<3><<<1, threads>>>(out, iters);
        if (mode == 4) wait_probe<4><<<1, threads>>>(out, iters);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      printf("%-22s threads %4d: %.1f clk per satisfied wait (ok %lld) %s\n", names[mode], threads, (double)h[0] / iters, h[1],
             cudaGetErrorString(cudaGetLastError()));
    }
  const char* rn[] = {"HFMA2", "tanh.f16x2", "tanh.f32", "HFMA2+tanh.f16x2(4:1)", "cvt+tanh.f16x2+unpack"};
  const int it2 = 4096;
  for (int mode = 0; mode < 5; ++mode)
    for (int warps = 4; warps <= 32; warps *= 2) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) rate_probe<0><<<148, warps * 32>>>(fo, out, it2);
        if (mode == 1) rate_probe<1><<<148, warps * 32>>>(fo, out, it2);
        if (mode == 2) rate_probe<2><<<148, warps * 32>>>(fo, out, it2);
        if (mode == 3) rate_probe<3><<<148, warps * 32>>>(fo, out, it2);
        if (mode == 4) rate_probe<4><<<148, warps * 32>>>(fo, out, it2);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h, out, sizeof(long long), cudaMemcpyDeviceToHost);
      const double instr = (double)it2 * 8 * warps;  // warp-instructions of the main kind per SM
      printf("%-24s warps/SM %2d: %.2f warp-instr/clk/SM (%.0f lane-ops/clk/SM) %s\n", rn[mode], warps, instr / h[0],
             instr * 32 / h[0], cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
