// Standalone probe (not part of the library): validates the tcgen05 plumbing the fused edge kernel relies on and
// measures the rates its design hinges on.
//   1. correctness of kind::f16 MMAs with A in TMEM (TS) or SMEM (SS), B in SMEM either MN-major or K-major with
//      128B swizzle, N = 16 sub-tiles addressed by byte offsets inside the swizzle atom;
//   2. MMA issue/throughput for N in {16,32,64,128}; tcgen05.ld bandwidth; MUFU.TANH and packed FFMA2 rates.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a umma_probe.cu -o umma_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <cmath>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// bounded wait: returns false on timeout instead of hanging the GPU
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (int it = 0; it < 2000000 && !done; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
  return done != 0;
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}

// smem matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
// layout type [61,64) (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor, kind::f16: D=f32, A=B=bf16
__host__ __device__ inline uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;   // c_format f32
  d |= 1u << 7;   // a_format bf16
  d |= 1u << 10;  // b_format bf16
  d |= (uint32_t)a_mn_major << 15;
  d |= (uint32_t)b_mn_major << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

#define TMEM_LD8(taddr, r)                                                                              \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"               \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) \
               : "r"(taddr))
#define TMEM_ST8(taddr, r)                                                                              \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"               \
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory")

constexpr int K = 64;        // contraction length of the correctness test
constexpr int NE = 64;       // edge columns in the B tile (one MN swizzle atom wide)
constexpr int DCOL = 256;    // accumulator column base

// mode 0: TS, B MN-major   mode 1: TS, B K-major   mode 2: SS (A K-major smem), B MN-major   mode 3: SS, B K-major
__global__ void __launch_bounds__(128) correctness_kernel(const uint32_t* __restrict__ a_packed,  // [128][K/2]
                                                         const __nv_bfloat16* __restrict__ b,      // [K][NE]
                                                         float* __restrict__ d_out,                // [128][NE]
                                                         int mode, int nsub, int* status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;              // 8 KB: MN-major [K rows][128 B] or K-major [NE rows][128 B]
  uint8_t* sA = smem + 8192;       // 16 KB: K-major [128 rows][128 B]
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) mbar_init(&bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

  // A (weights) -> TMEM, lane = row, 2 bf16 per column
  for (int c = 0; c < K / 2; c += 8) {
    uint32_t r[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) r[q] = a_packed[tid * (K / 2) + c + q];
    TMEM_ST8(tmem + lane_base + c, r);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  // A (weights) -> smem K-major SW128: row r at r*128, 16B chunk c at (c ^ (r%8))
  for (int idx = tid; idx < 128 * 8; idx += 128) {
    int r = idx >> 3, c = idx & 7;
    uint4 v = *reinterpret_cast<const uint4*>(a_packed + r * (K / 2) + c * 4);
    *reinterpret_cast<uint4*>(sA + r * 128 + ((c ^ (r & 7)) * 16)) = v;
  }
  // B -> smem
  if (mode == 0 || mode == 2) {  // MN-major: row k, chunk of 8 edges
    for (int idx = tid; idx < K * 8; idx += 128) {
      int k = idx >> 3, c = idx & 7;
      uint4 v = *reinterpret_cast<const uint4*>(b + k * NE + c * 8);
      *reinterpret_cast<uint4*>(sB + k * 128 + ((c ^ (k & 7)) * 16)) = v;
    }
  } else {  // K-major: row e, 64 k contiguous
    for (int idx = tid; idx < NE * K; idx += 128) {
      int e = idx / K, k = idx % K;
      int chunk = k >> 3;
      *reinterpret_cast<__nv_bfloat16*>(sB + e * 128 + ((chunk ^ (e & 7)) * 16) + (k & 7) * 2) = b[k * NE + e];
    }
  }
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int nper = NE / nsub;  // columns per MMA
  if (tid == 0) {
    const bool b_mn = (mode == 0 || mode == 2);
    const uint32_t idesc = make_idesc(128, nper, 0, b_mn ? 1 : 0);
    for (int sub = 0; sub < nsub; ++sub) {
      for (int s = 0; s < K / 16; ++s) {
        uint64_t bdesc;
        if (b_mn) bdesc = make_desc(smem_u32(sB) + s * 2048 + sub * nper * 2, 16, 1024);
        else      bdesc = make_desc(smem_u32(sB) + sub * nper * 128 + s * 32, 16, 1024);
        const uint32_t dt = tmem + DCOL + sub * nper;
        if (mode < 2) mma_ts(dt, tmem + s * 8, bdesc, idesc, s > 0);
        else          mma_ss(dt, make_desc(smem_u32(sA) + s * 32, 16, 1024), bdesc, idesc, s > 0);
      }
    }
    tc_commit(&bar);
  }
  bool ok = mbar_wait(&bar, 0);
  if (!ok && tid == 0) *status = -1;
  tc_fence_after();
  for (int c = 0; c < NE; c += 8) {
    uint32_t r[8];
    TMEM_LD8(tmem + lane_base + DCOL + c, r);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int q = 0; q < 8; ++q) d_out[tid * NE + c + q] = __uint_as_float(r[q]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// MMA rate: `iters` back-to-back MMAs of shape 128 x N x 16 (same operands), then commit + wait.
__global__ void __launch_bounds__(128) mma_rate_kernel(int N, int ts, int iters, long long* cycles, int* status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) mbar_init(&bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, N, 0, 1);
    const uint64_t bdesc = make_desc(smem_u32(smem), 16, 1024);
    const uint64_t adesc = make_desc(smem_u32(smem) + 16384, 16, 1024);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      // rotate over 8 K-slices so operand addresses change like in the real kernel
      const int s = i & 7;
      if (ts) mma_ts(tmem + 256, tmem + s * 8, bdesc + (uint64_t)(s * 128), idesc, 1);
      else    mma_ss(tmem + 256, adesc + (uint64_t)(s * 2), bdesc + (uint64_t)(s * 128), idesc, 1);
    }
    tc_commit(&bar);
    bool ok = mbar_wait(&bar, 0);
    long long t1 = clock64();
    *cycles = t1 - t0;
    if (!ok) *status = -2;
  }
  __syncthreads();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}


// MMA rate v2: fully unrolled issue (descriptors precomputed in registers), rotating over NACC accumulators of
// N columns each, so issue overhead, dependent-accumulate latency and tensor throughput can be told apart.
template <int NACC, int UNROLL>
__global__ void __launch_bounds__(128) mma_rate2_kernel(int N, int ts, int iters, long long* cycles, int* status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) mbar_init(&bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, N, 0, 1);
    uint64_t bdesc[UNROLL], adesc[UNROLL];
    uint32_t atm[UNROLL], dtm[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      bdesc[u] = make_desc(smem_u32(smem) + (u & 7) * 2048, 16, 1024);
      adesc[u] = make_desc(smem_u32(smem) + 16384 + (u & 3) * 32, 16, 1024);
      atm[u] = tmem + (u & 7) * 8;
      dtm[u] = tmem + 128 + (u % NACC) * N;
    }
    long long t0 = clock64();
    for (int i = 0; i < iters; i += UNROLL) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        if (ts) mma_ts(dtm[u], atm[u], bdesc[u], idesc, 1);
        else    mma_ss(dtm[u], adesc[u], bdesc[u], idesc, 1);
      }
    }
    tc_commit(&bar);
    bool ok = mbar_wait(&bar, 0);
    long long t1 = clock64();
    *cycles = t1 - t0;
    if (!ok) *status = -2;
  }
  __syncthreads();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int NACC>
static void run_rate2(int N, int ts, long long* d_cyc, int* d_status) {
  const int iters = 4096;
  if (128 + NACC * N > 512) return;
  CK(cudaFuncSetAttribute(mma_rate2_kernel<NACC, 24>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  CK(cudaMemset(d_status, 0, 4));
  mma_rate2_kernel<NACC, 24><<<1, 128, 64 * 1024>>>(N, ts, iters - iters % 24, d_cyc, d_status);
  CK(cudaDeviceSynchronize());
  long long cyc; int status;
  CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&status, d_status, 4, cudaMemcpyDeviceToHost));
  const int done = iters - iters % 24;
  printf("mma_rate2 %s N=%3d NACC=%d unrolled: %.1f clk/MMA (%.0f MAC/clk) status=%d\n", ts ? "TS" : "SS", N, NACC,
         (double)cyc / done, 128.0 * N * 16 * done / (double)cyc, status);
}


__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}

// MMA rate v3: the whole warp runs the (warp-uniform) issue loop, only the tcgen05.mma is under elect.sync, so the
// compiler can keep descriptors in uniform registers.
template <int NACC>
__global__ void __launch_bounds__(128) mma_rate3_kernel(int N, int ts, int iters, long long* cycles, int* status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) mbar_init(&bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0) {
    const uint32_t idesc = make_idesc(128, N, 0, 1);
    const uint32_t sbase = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
    const uint64_t b0 = make_desc(sbase, 16, 1024);
    const uint64_t a0 = make_desc(sbase + 16384, 16, 1024);
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 8 * NACC) {
#pragma unroll
      for (int s = 0; s < 8; ++s) {
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
          if (elect_one()) {
            if (ts) mma_ts(tmem + 128 + a * N, tmem + s * 8, b0 + (uint64_t)(s * 128), idesc, 1);
            else    mma_ss(tmem + 128 + a * N, a0 + (uint64_t)((s & 3) * 2), b0 + (uint64_t)(s * 128), idesc, 1);
          }
        }
      }
    }
    if (elect_one()) tc_commit(&bar);
    __syncwarp();
    bool ok = mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (tid == 0) { *cycles = t1 - t0; if (!ok) *status = -2; }
  }
  __syncthreads();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int NACC>
static void run_rate3(int N, int ts, long long* d_cyc, int* d_status) {
  if (128 + NACC * N > 512) return;
  const int iters = 8 * NACC * 64;
  CK(cudaFuncSetAttribute(mma_rate3_kernel<NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  CK(cudaMemset(d_status, 0, 4));
  mma_rate3_kernel<NACC><<<1, 128, 64 * 1024>>>(N, ts, iters, d_cyc, d_status);
  CK(cudaDeviceSynchronize());
  long long cyc; int status;
  CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&status, d_status, 4, cudaMemcpyDeviceToHost));
  printf("mma_rate3 %s N=%3d NACC=%d uniform-issue: %.1f clk/MMA (%.0f MAC/clk) status=%d\n", ts ? "TS" : "SS", N, NACC,
         (double)cyc / iters, 128.0 * N * 16 * iters / (double)cyc, status);
}


// MMA rate v4: NWARPS warps issue concurrently (each its own accumulator), to tell a tensor-pipe floor from a
// single-warp issue limit.
__global__ void __launch_bounds__(256) mma_rate4_kernel(int N, int nwarps, int iters, long long* cycles, int* status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[8];
  __shared__ uint32_t tmem_base_s;
  __shared__ long long tstart[8], tend[8];
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 1024 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid < 8) mbar_init(&bar[tid], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp < nwarps) {
    const uint32_t idesc = make_idesc(128, N, 0, 1);
    const uint32_t sbase = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
    const uint64_t b0 = make_desc(sbase, 16, 1024);
    const uint32_t dacc = tmem + 128 + warp * N;
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        if (elect_one()) mma_ts(dacc, tmem + s * 8, b0 + (uint64_t)(s * 128), idesc, 1);
      }
    }
    if (elect_one()) tc_commit(&bar[warp]);
    __syncwarp();
    bool ok = mbar_wait(&bar[warp], 0);
    long long t1 = clock64();
    if ((tid & 31) == 0) { tstart[warp] = t0; tend[warp] = t1; if (!ok) *status = -2; }
  }
  __syncthreads();
  if (tid == 0) {
    long long a = tstart[0], b = tend[0];
    for (int w = 1; w < nwarps; ++w) { a = min(a, tstart[w]); b = max(b, tend[w]); }
    *cycles = b - a;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// tcgen05.ld bandwidth: every warp streams x16 loads over 256 columns of its lane quadrant
__global__ void __launch_bounds__(256) tmem_ld_kernel(int iters, long long* cycles, float* sink) {
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t r[8], q[8];
    const uint32_t col = (uint32_t)((i * 16) & 255) + (warp >= 4 ? 256 : 0);
    TMEM_LD8(tmem + lane_base + col, r);
    TMEM_LD8(tmem + lane_base + col + 8, q);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += __uint_as_float(r[k] & 0x007fffffu) + __uint_as_float(q[k] & 0x007fffffu);
  }
  long long t1 = clock64();
  if (tid == 0) *cycles = t1 - t0;
  sink[blockIdx.x * blockDim.x + tid] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

__global__ void mufu_kernel(int iters, long long* cycles, float* sink) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 0.1f, x2 = x0 + 0.2f, x3 = x0 + 0.3f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x0));
    asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x1));
    asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x2));
    asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x3));
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) *cycles = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}

__global__ void ffma_kernel(int iters, int packed, long long* cycles, float* sink) {
  float a[8], b = 1.0001f, c = 0.5f;
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
  __syncthreads();
  long long t0 = clock64();
  if (packed) {
    unsigned long long av[4], bv, cv;
    for (int k = 0; k < 4; ++k) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(av[k]) : "f"(a[2 * k]), "f"(a[2 * k + 1]));
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(cv) : "f"(c));
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(av[k]) : "l"(bv), "l"(cv));
    }
    for (int k = 0; k < 4; ++k) asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * k]), "=f"(a[2 * k + 1]) : "l"(av[k]));
  } else {
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = fmaf(a[k], b, c);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) *cycles = t1 - t0;
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += a[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  // ---------------- correctness ----------------
  std::vector<float> A(128 * K), B(K * NE);
  srand(1);
  for (auto& v : A) v = bf16_round((rand() % 2001 - 1000) / 1000.0f);
  for (auto& v : B) v = bf16_round((rand() % 2001 - 1000) / 1000.0f);
  std::vector<uint32_t> a_packed(128 * K / 2);
  std::vector<__nv_bfloat16> b_bf(K * NE);
  for (int r = 0; r < 128; ++r)
    for (int k = 0; k < K; k += 2) {
      __nv_bfloat16 lo = __float2bfloat16(A[r * K + k]), hi = __float2bfloat16(A[r * K + k + 1]);
      a_packed[r * (K / 2) + k / 2] = (uint32_t)(*(uint16_t*)&lo) | ((uint32_t)(*(uint16_t*)&hi) << 16);
    }
  for (int i = 0; i < K * NE; ++i) b_bf[i] = __float2bfloat16(B[i]);
  std::vector<float> ref(128 * NE, 0.f);
  for (int m = 0; m < 128; ++m)
    for (int e = 0; e < NE; ++e) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[k * NE + e];
      ref[m * NE + e] = (float)s;
    }
  uint32_t* d_a; __nv_bfloat16* d_b; float* d_d; int* d_status; long long* d_cyc; float* d_sink;
  CK(cudaMalloc(&d_a, a_packed.size() * 4)); CK(cudaMalloc(&d_b, b_bf.size() * 2)); CK(cudaMalloc(&d_d, 128 * NE * 4));
  CK(cudaMalloc(&d_status, 4)); CK(cudaMalloc(&d_cyc, 8)); CK(cudaMalloc(&d_sink, 148 * 1024 * 4));
  CK(cudaMemcpy(d_a, a_packed.data(), a_packed.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_b, b_bf.data(), b_bf.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(correctness_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  CK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  const char* names[4] = {"TS A=tmem  B=MN-major", "TS A=tmem  B=K-major ", "SS A=smemK B=MN-major", "SS A=smemK B=K-major "};
  for (int mode = 0; mode < 4; ++mode)
    for (int nsub : {1, 2, 4}) {
      CK(cudaMemset(d_d, 0, 128 * NE * 4)); CK(cudaMemset(d_status, 0, 4));
      correctness_kernel<<<1, 128, 64 * 1024>>>(d_a, d_b, d_d, mode, nsub, d_status);
      CK(cudaDeviceSynchronize());
      std::vector<float> out(128 * NE); int status;
      CK(cudaMemcpy(out.data(), d_d, out.size() * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(&status, d_status, 4, cudaMemcpyDeviceToHost));
      double maxerr = 0;
      for (size_t i = 0; i < out.size(); ++i) maxerr = fmax(maxerr, fabs(out[i] - ref[i]));
      printf("correctness %s N/MMA=%2d: status=%d max_abs_err=%.3e %s\n", names[mode], NE / nsub, status, maxerr,
             (status == 0 && maxerr < 1e-3) ? "OK" : "FAIL");
    }
  // ---------------- rates ----------------
  for (int ts = 1; ts >= 0; --ts)
    for (int N : {16, 32, 64, 128, 256}) {
      if (256 + N > 512) continue;
      const int iters = 4096;
      CK(cudaMemset(d_status, 0, 4));
      mma_rate_kernel<<<1, 128, 64 * 1024>>>(N, ts, iters, d_cyc, d_status);
      CK(cudaDeviceSynchronize());
      long long cyc; int status;
      CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&status, d_status, 4, cudaMemcpyDeviceToHost));
      printf("mma_rate %s M=128 N=%3d K=16: %.1f clk/MMA (%.0f MAC/clk) status=%d\n", ts ? "TS" : "SS", N,
             (double)cyc / iters, 128.0 * N * 16 * iters / (double)cyc, status);
    }
  for (int ts = 1; ts >= 0; --ts)
    for (int N : {16, 32, 64, 128}) {
      run_rate2<1>(N, ts, d_cyc, d_status);
      run_rate2<2>(N, ts, d_cyc, d_status);
      run_rate2<3>(N, ts, d_cyc, d_status);
      run_rate2<6>(N, ts, d_cyc, d_status);
    }
  for (int ts = 1; ts >= 0; --ts)
    for (int N : {16, 32, 64, 128}) {
      run_rate3<1>(N, ts, d_cyc, d_status);
      run_rate3<2>(N, ts, d_cyc, d_status);
      run_rate3<6>(N, ts, d_cyc, d_status);
    }
  CK(cudaFuncSetAttribute(mma_rate4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  for (int N : {16, 32, 64})
    for (int nw : {1, 2, 4, 6}) {
      if (128 + nw * N > 512) continue;
      const int iters = 2048;
      CK(cudaMemset(d_status, 0, 4));
      mma_rate4_kernel<<<1, 256, 64 * 1024>>>(N, nw, iters, d_cyc, d_status);
      CK(cudaDeviceSynchronize());
      long long cyc; int status;
      CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&status, d_status, 4, cudaMemcpyDeviceToHost));
      printf("mma_rate4 TS N=%3d issuing warps=%d: %.1f clk per MMA overall (%.0f MAC/clk) status=%d\n", N, nw,
             (double)cyc / (iters * nw), 128.0 * N * 16 * iters * nw / (double)cyc, status);
    }
  for (int threads : {128, 256}) {
    const int iters = 4096;
    tmem_ld_kernel<<<1, threads>>>(iters, d_cyc, d_sink);
    CK(cudaDeviceSynchronize());
    long long cyc; CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
    printf("tmem_ld %d threads: %.1f B/clk/SM\n", threads, (double)threads * 16 * 4 * iters / (double)cyc);
  }
  for (int threads : {128, 512, 1024}) {
    const int iters = 4096;
    mufu_kernel<<<1, threads>>>(iters, d_cyc, d_sink);
    CK(cudaDeviceSynchronize());
    long long cyc; CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
    printf("mufu tanh %4d threads: %.2f ops/clk/SM\n", threads, (double)threads * 4 * iters / (double)cyc);
  }
  for (int packed = 0; packed < 2; ++packed)
    for (int threads : {512, 1024}) {
      const int iters = 4096;
      ffma_kernel<<<1, threads>>>(iters, packed, d_cyc, d_sink);
      CK(cudaDeviceSynchronize());
      long long cyc; CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
      printf("%s %4d threads: %.1f FMA/clk/SM\n", packed ? "fma.f32x2" : "ffma     ", threads,
             (double)threads * 8 * iters / (double)cyc);
    }
  printf("probe done\n");
  return 0;
}
