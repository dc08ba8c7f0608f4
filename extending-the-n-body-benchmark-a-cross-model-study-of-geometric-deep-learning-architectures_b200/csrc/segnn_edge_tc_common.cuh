// Shared device helpers of the tcgen05 edge kernels (segnn_edge_tc.cu, segnn_edge_tc_h2.cu): mbarrier / tcgen05 PTX
// wrappers, UMMA descriptors, packed-fp32 helpers, the optional timeline trace.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "segnn_common.cuh"

namespace segnn {
namespace tc {

constexpr int kRecv = 4;      // receivers per work item
constexpr int kSend = 8;      // senders per tile
constexpr int kCols = 32;     // edges (columns) per tile
constexpr int kGeoSlots = 8;  // geometry ring depth (geometry is written three tiles ahead, read until epilogue(t))
constexpr int kWarps = 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded spin: a protocol bug must never hang the GPU. On timeout the flag is raised and the kernel traps.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err_flag) {
  uint32_t done = 0;
#pragma unroll 1
  for (int it = 0; it < (1 << 22); ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)  // suspend-time hint: fewer polls steal issue slots
        : "memory");
    if (done) return;
  }
  if (err_flag) atomicExch(err_flag, 1);
  __trap();
}
// Same wait on a precomputed shared-memory address, as one PTX loop: the fast path (phase already complete) is
// mov + try_wait + branch; still bounded (traps after ~4M polls instead of hanging the GPU).
__device__ __forceinline__ void mbar_wait_a(uint32_t bar_addr, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .u32 c;\n\t"
      // first poll without a suspend-time hint: a satisfied wait costs 54 clk instead of 85 (experiments/wait_probe.cu)
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra SEGNN_DONE_%=;\n\t"
      "mov.u32 c, 0;\n"
      "SEGNN_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"
      "@p bra SEGNN_DONE_%=;\n\t"
      "add.u32 c, c, 1;\n\t"
      "setp.lt.u32 p, c, 4194304;\n\t"
      "@p bra SEGNN_WAIT_%=;\n\t"
      "trap;\n"
      "SEGNN_DONE_%=:\n\t}" ::"r"(bar_addr),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_barrier(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem desc]   (kind::f16: bf16 x bf16 -> fp32)
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// UMMA shared-memory matrix descriptor: MN-major, SWIZZLE_128B, 8-row k-groups 1024 B apart.
__device__ __forceinline__ uint64_t make_b_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;            // LBO (unused: the tile is one swizzle atom wide)
  d |= (uint64_t)(1024 >> 4) << 32;  // SBO
  d |= (uint64_t)1 << 46;            // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}
// instruction descriptor: D = f32, A = B = bf16 or f16, A K-major (TMEM), B MN-major, M = 128, N = 32
__device__ __forceinline__ uint32_t make_idesc(bool half) {
  uint32_t d = 0;
  d |= 1u << 4;
  if (!half) {
    d |= 1u << 7;   // A format: 0 = f16, 1 = bf16
    d |= 1u << 10;  // B format
  }
  d |= 1u << 16;
  d |= (uint32_t)(kCols >> 3) << 17;
  d |= (uint32_t)(128 >> 4) << 24;
  return d;
}

#define SEGNN_TMEM_LD8(taddr, r)                                                                                   \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                            \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])     \
               : "r"(taddr))
#define SEGNN_TMEM_ST8(taddr, r)                                                                                   \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),   \
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])                          \
               : "memory")

// one elected lane of a converged warp (the tcgen05.mma / commit instructions are issued by a single thread, but the
// surrounding code stays warp-uniform so descriptors live in uniform registers: profiles/r1_umma_probe.log)
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}

// two fp32 -> one 32-bit pair of 16-bit operands (bf16, or fp16 in SEGNN_MODE_FP16_TC)
template <bool HALF>
__device__ __forceinline__ uint32_t pack_pair(float lo, float hi) {
  uint32_t r;
  if (HALF)
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// gates with one MUFU each: sigmoid(x) = 0.5 tanh(x/2) + 0.5
__device__ __forceinline__ float silu_gate_fast(float x) {
  const float t = tanh_fast(0.5f * x);
  return (0.5f * kCSilu) * x * (1.0f + t);
}
__device__ __forceinline__ float sig_gate_fast(float x) {
  const float t = tanh_fast(0.5f * x);
  return fmaf(0.5f * kCSig, t, 0.5f * kCSig);
}


#define SEGNN_TMEM_LD4(taddr, r)                                                          \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"               \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])                           \
               : "r"(taddr))

__device__ __forceinline__ float2 bc2(float x) { return make_float2(x, x); }
__device__ __forceinline__ float2 tanh2(float2 x) { return make_float2(tanh_fast(x.x), tanh_fast(x.y)); }
__device__ __forceinline__ float2 lo2(float4 v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(float4 v) { return make_float2(v.z, v.w); }
__device__ __forceinline__ float2 u2f2(uint32_t a, uint32_t b) { return make_float2(__uint_as_float(a), __uint_as_float(b)); }

// ---- optional timeline tracing (compile with -DSEGNN_K3_TRACE; csrc/experiments/k3_trace.py reads the buffer) --------
#ifdef SEGNN_K3_TRACE
extern __device__ long long* g_k3_trace;  // [16 warps][64 tiles][8 events]
#define K3_TRACE(ev, t)                                                                           \
  do {                                                                                            \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (t) < 64u && g_k3_trace != nullptr)        \
      g_k3_trace[((threadIdx.x >> 5) * 64 + (t)) * 8 + (ev)] = clock64();                          \
  } while (0)
#else
#define K3_TRACE(ev, t) \
  do {                  \
  } while (0)
#endif

struct TileCursor {
  int item;
  int sb;
  uint32_t t;
  int g, i0;  // graph and first receiver of the item (MMA warp only; recomputed once per item)
};


}  // namespace tc
}  // namespace segnn
