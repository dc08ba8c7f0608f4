// K3 (bf16 tensor-core mode): fused SEGNN edge layer on tcgen05.
//
// Orientation: OUTPUT CHANNELS are the MMA M dimension (TMEM lanes), EDGES are the MMA N dimension (TMEM columns).
//   D_tile[128 lanes = channel w][32 cols = edges] += A_tile[128 x K] (message_layer_2 weights, resident in TMEM,
//   tcgen05.mma "TS" form) x B[K x 32] (per-edge features: gate(message_layer_1) output, bf16 in shared memory,
//   MN-major, 128B swizzle).
// With channels on the lanes, the producer (thread = channel), the tensor core output and the epilogue (thread =
// lane = channel) all agree on the mapping: node projections are read coalesced, the B tile is written with one
// 8-byte store per plane and sender, the gate needs no cross-lane traffic and the sum over senders is a plain
// in-register accumulation (no atomics, no shuffles).
//
// Six accumulator tiles per 32-edge tile: s (K = 2n: [s' | v'.a1]), g (same K), T1 = W_sv^T s' (K = n),
// D_k = W_vv^T v'_k (K = n, k = x,y,z); message = (silu(s), sigmoid(g) * (a1_k * T1 + D_k)).
// The six dependent-accumulate chains are interleaved because a tcgen05.mma that accumulates into the tile of
// its predecessor waits ~67 clk (measured, profiles/r1_umma_probe.log).
//
// Warp roles (one CTA per SM, persistent over work items = (graph, block of 4 receivers)):
//   4 producer groups (n threads each): group q owns senders 2q, 2q+1 of the 8-sender block and all 4 receivers;
//   2 epilogue groups (128 threads each, lane quadrant = warp % 4): group e owns tile columns [16e, 16e+16);
//   1 MMA warp (warp-uniform loop, one elected lane issues; it also owns the TMEM allocation and issues one
//   cp.async.bulk per tile: the 8 consecutive sender rows of the Q projection, 36 KB at n = 96, double buffered;
//   the copy for tile t+2 is issued when the producers have signalled tile t complete).
// Tile column c = sender_local * 4 + receiver_local.
#include <cuda_bf16.h>

#include "segnn_common.cuh"

namespace segnn {
namespace tc {

constexpr int kRecv = 4;      // receivers per work item
constexpr int kSend = 8;      // senders per tile
constexpr int kCols = 32;     // edges (columns) per tile
constexpr int kGeoSlots = 8;  // geometry ring depth (tile t+1 is written while tile t computes)
constexpr int kMaxNodesPerGraph = 2048;  // positions of one graph are staged in shared memory

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
// instruction descriptor: D = f32, A = B = bf16, A K-major (TMEM), B MN-major, M = 128, N = 32
__device__ __forceinline__ uint32_t make_idesc() {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 1u << 7;
  d |= 1u << 10;
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

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
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

// Warp layout (NW = n / 32): producers [0, 4 NW); epilogue group e at [4 NW + 4 e, 4 NW + 4 e + NW) so that
// warp % 4 (the TMEM lane quadrant a warp may access) equals the channel block; the MMA warp sits in the unused
// quadrant slot of epilogue group 0.
template <int NMUL>
constexpr int tc_num_warps() { return 4 * (NMUL / 32) + 4 + NMUL / 32; }

template <int NMUL>
__global__ void __launch_bounds__(tc_num_warps<NMUL>() * 32, 1)
    edge_layer_tc_kernel(const float* __restrict__ pos, const float* __restrict__ mass, int B, int N,
                         const float* __restrict__ pp, const float* __restrict__ qq,
                         const float* __restrict__ w_edge1,
                         const float* __restrict__ b2, const uint32_t* __restrict__ w2_tc,
                         const float* __restrict__ bn_mul, const float* __restrict__ bn_add,
                         float* __restrict__ agg, int* __restrict__ err_flag) {
  constexpr int n = NMUL;
  constexpr int NW = n / 32;            // warps per producer group
  constexpr int kProdWarps = 4 * NW;
  constexpr int kEpiWarp0 = kProdWarps;
  static_assert(NW < 4, "the MMA warp sits in the unused quadrant slot of epilogue group 0");
  constexpr int kMmaWarp = kEpiWarp0 + NW;
  constexpr int kQStageBytes = kSend * 4 * 3 * n * (int)sizeof(float);
  constexpr int kEpiThreads = 2 * NW * 32;
  constexpr int kWeightCols = 3 * n;    // TMEM columns of the weight image
  constexpr int kDBase = kWeightCols;   // accumulator tiles start here (6 x 32 columns)
  constexpr int n3 = 3 * n;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
  uint8_t* sB = smem;                                                  // 5n * 128 bytes
  float* sQ = reinterpret_cast<float*>(smem + 5 * n * 128);            // [2][8 senders][4 planes][3n]
  float4* sPos = reinterpret_cast<float4*>(smem + 5 * n * 128 + 2 * kQStageBytes);  // [N] (x, y, z, mass)
  float4* geoA = sPos + N;                                             // [slots][32] (ax, ay, az, valid)
  float2* geoB = reinterpret_cast<float2*>(geoA + kGeoSlots * kCols);  // [slots][32] (dist, m_i m_j)
  float* xch = reinterpret_cast<float*>(geoB + kGeoSlots * kCols);     // [4 recv][4 planes][n]
  uint64_t* bars = reinterpret_cast<uint64_t*>(xch + kRecv * 4 * n);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  uint64_t* full = bars;
  uint64_t* empty = bars + 2;
  uint64_t* dfull = bars + 4;
  uint64_t* dempty = bars + 5;
  uint64_t* qfull = bars + 6;   // [2] bulk copy landed (expect_tx)
  uint64_t* gfull = bars + 10;  // [8] geometry ring: one barrier per slot, one arrival per producer group

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&full[0], 4 * n);
    mbar_init(&full[1], 4 * n);
    mbar_init(&empty[0], 1);
    mbar_init(&empty[1], 1);
    mbar_init(dfull, 1);
    mbar_init(dempty, kEpiThreads);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qfull[i], 1);
    }
    for (int i = 0; i < kGeoSlots; ++i) mbar_init(&gfull[i], 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // ---- message_layer_2 weights -> TMEM (lane = output channel, 2 bf16 of K per column) ------------------------
  if (warp >= kEpiWarp0 && warp < kEpiWarp0 + NW) {  // lanes >= n are never read back
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t* src = w2_tc + (size_t)row * kWeightCols;
    for (int c = 0; c < kWeightCols; c += 8) {
      uint32_t r[8];
      const uint4 v0 = *reinterpret_cast<const uint4*>(src + c);
      const uint4 v1 = *reinterpret_cast<const uint4*>(src + c + 4);
      r[0] = v0.x; r[1] = v0.y; r[2] = v0.z; r[3] = v0.w;
      r[4] = v1.x; r[5] = v1.y; r[6] = v1.z; r[7] = v1.w;
      SEGNN_TMEM_ST8(tmem + lane_base + c, r);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int recv_blocks = (N + kRecv - 1) / kRecv;
  const int send_blocks = (N + kSend - 1) / kSend;
  const long long items = (long long)B * recv_blocks;

  if (warp < kProdWarps) {
    // ============================ producers: message_layer_1 (hoisted) + gate -> B tile ========================
    const int q = warp / NW;                  // group = sender pair
    const int w = (warp % NW) * 32 + lane;    // channel
    const float wd0s = w_edge1[w], wd0g = w_edge1[n + w], wm0s = w_edge1[2 * n + w], wm0g = w_edge1[3 * n + w],
                wd1 = w_edge1[4 * n + w], wm1 = w_edge1[5 * n + w];
    // geometry of this group's 8 columns (2 senders x 4 receivers) of sender block sb, one thread per column
    auto write_geometry = [&](int sb, int i0, int slot) {
      if (w < 8) {
        const int sl = 2 * q + (w >> 2), r = w & 3;
        const int jj = sb * kSend + sl, ii = i0 + r;
        const float4 ps = sPos[min(jj, N - 1)], pr = sPos[min(ii, N - 1)];
        float ux, uy, uz, len;
        unit_vec(ps.x - pr.x, ps.y - pr.y, ps.z - pr.z, ux, uy, uz, len);
        const bool valid = (jj < N) && (ii < N) && (jj != ii);
        const int c = sl * 4 + r;
        geoA[slot * kCols + c] = make_float4(kY1 * ux, kY1 * uy, kY1 * uz, valid ? 1.0f : 0.0f);
        geoB[slot * kCols + c] = make_float2(len, ps.w * pr.w);
      }
      named_barrier(1 + q, n);
      if (w == 0) mbar_arrive(&gfull[slot]);  // publishes the group's 8 geometry entries to the epilogue
    };
    uint32_t t = 0;
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
      const long long g = item / recv_blocks;
      const int i0 = (int)(item - g * recv_blocks) * kRecv;
      const long long base = g * N;
      // stage the graph's positions and masses (all producer groups together)
      named_barrier(5, 4 * n);
      for (int i = q * n + w; i < N; i += 4 * n) {
        const long long node = base + i;
        sPos[i] = make_float4(pos[node * 3 + 0], pos[node * 3 + 1], pos[node * 3 + 2], mass[node]);
      }
      named_barrier(5, 4 * n);
      // receiver-side projections of the 4 receivers (clamped: invalid receivers are masked by the epilogue)
      float p0s[kRecv], p0g[kRecv], p1[kRecv], p0sk[kRecv][3], p0gk[kRecv][3], p1k[kRecv][3];
#pragma unroll
      for (int r = 0; r < kRecv; ++r) {
        const long long node = base + min(i0 + r, N - 1);
        const float* pr = pp + node * 4 * n3;
        p0s[r] = pr[w];
        p0g[r] = pr[n + w];
        p1[r] = pr[2 * n + w];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float* prk = pr + (1 + k) * n3;
          p0sk[r][k] = prk[w];
          p0gk[r][k] = prk[n + w];
          p1k[r][k] = prk[2 * n + w];
        }
      }
      write_geometry(0, i0, t & (kGeoSlots - 1));
      for (int sb = 0; sb < send_blocks; ++sb, ++t) {
        const int st = t & 1, slot = t & (kGeoSlots - 1);
        const int nvalid = min(kSend, N - sb * kSend);
        mbar_wait(&empty[st], ((t >> 1) & 1) ^ 1, err_flag);
        mbar_wait(&qfull[st], (t >> 1) & 1, err_flag);
        const float* qs = sQ + st * (kQStageBytes / 4);
        uint2 packed[2][5];
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          const int sl = 2 * q + s2;
          const float* qr = qs + min(sl, nvalid - 1) * 4 * n3;  // senders past the graph end reuse a valid row
          const float q0s = qr[w], q0g = qr[n + w], q1 = qr[2 * n + w];
          float q0sk[3], q0gk[3], q1k[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const float* qk = qr + (1 + k) * n3;
            q0sk[k] = qk[w];
            q0gk[k] = qk[n + w];
            q1k[k] = qk[2 * n + w];
          }
          float o_s[kRecv], o_d[kRecv], o_x[kRecv], o_y[kRecv], o_z[kRecv];
#pragma unroll
          for (int r = 0; r < kRecv; ++r) {
            const float4 ga = geoA[slot * kCols + sl * 4 + r];
            const float2 gb = geoB[slot * kCols + sl * 4 + r];
            float zs = (p0s[r] + q0s) + ga.x * (p0sk[r][0] + q0sk[0]) + ga.y * (p0sk[r][1] + q0sk[1]) +
                       ga.z * (p0sk[r][2] + q0sk[2]) + gb.x * wd0s + gb.y * wm0s;
            float zg = (p0g[r] + q0g) + ga.x * (p0gk[r][0] + q0gk[0]) + ga.y * (p0gk[r][1] + q0gk[1]) +
                       ga.z * (p0gk[r][2] + q0gk[2]) + gb.x * wd0g + gb.y * wm0g;
            const float tt = (p1[r] + q1) + gb.x * wd1 + gb.y * wm1;
            const float gg = sig_gate_fast(zg);
            const float vx = gg * (ga.x * tt + (p1k[r][0] + q1k[0]));
            const float vy = gg * (ga.y * tt + (p1k[r][1] + q1k[1]));
            const float vz = gg * (ga.z * tt + (p1k[r][2] + q1k[2]));
            o_s[r] = silu_gate_fast(zs);
            o_d[r] = ga.x * vx + ga.y * vy + ga.z * vz;
            o_x[r] = vx;
            o_y[r] = vy;
            o_z[r] = vz;
          }
          packed[s2][0] = make_uint2(pack_bf16x2(o_s[0], o_s[1]), pack_bf16x2(o_s[2], o_s[3]));
          packed[s2][1] = make_uint2(pack_bf16x2(o_d[0], o_d[1]), pack_bf16x2(o_d[2], o_d[3]));
          packed[s2][2] = make_uint2(pack_bf16x2(o_x[0], o_x[1]), pack_bf16x2(o_x[2], o_x[3]));
          packed[s2][3] = make_uint2(pack_bf16x2(o_y[0], o_y[1]), pack_bf16x2(o_y[2], o_y[3]));
          packed[s2][4] = make_uint2(pack_bf16x2(o_z[0], o_z[1]), pack_bf16x2(o_z[2], o_z[3]));
        }
        // 2 senders x 4 receivers = 8 consecutive columns = one 16-byte chunk per plane row (conflict-free with
        // the 128B swizzle: 8 consecutive rows hit 8 distinct chunks)
        {
          const int chunk = st * 4 + q;
#pragma unroll
          for (int p = 0; p < 5; ++p) {
            const int row = p * n + w;
            *reinterpret_cast<uint4*>(sB + row * 128 + ((chunk ^ (row & 7)) << 4)) =
                make_uint4(packed[0][p].x, packed[0][p].y, packed[1][p].x, packed[1][p].y);
          }
        }
        proxy_fence();
        mbar_arrive(&full[st]);
        if (sb + 1 < send_blocks) write_geometry(sb + 1, i0, (t + 1) & (kGeoSlots - 1));
      }
    }
  } else if (warp == kMmaWarp) {
    // ============================ MMA issuer ==================================================================
    // The whole warp runs this warp-uniform loop; only the tcgen05 instructions sit under elect.sync. A lean issue
    // sequence matters: one warp sustains ~1 MMA / 24 clk, the tensor pipe needs one 128x32x16 MMA / 16 clk.
    const uint32_t idesc = make_idesc();
    const uint32_t sB_addr = __shfl_sync(0xffffffffu, smem_u32(sB), 0);
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint64_t bdesc0 = make_b_desc(sB_addr);
    const uint32_t d0 = tm + kDBase;
    // bulk copy of the 8 sender rows of tile (item, sb) into Q stage st (one elected lane)
    auto load_q = [&](long long item, int sb, int st) {
      if (item < items && elect_one()) {
        const long long g = item / recv_blocks;
        const int nvalid = min(kSend, N - sb * kSend);
        const uint32_t bytes = (uint32_t)nvalid * 4 * n3 * (uint32_t)sizeof(float);
        const float* src = qq + (g * N + (long long)sb * kSend) * 4 * n3;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&qfull[st])), "r"(bytes)
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(sQ + st * (kQStageBytes / 4))),
                     "l"(src), "r"(bytes), "r"(smem_u32(&qfull[st]))
                     : "memory");
      }
      __syncwarp();
    };
    // (item, sb) of the tile two ahead of the current one
    long long pf_item = blockIdx.x;
    int pf_sb = 0;
    auto advance_pf = [&]() {
      if (++pf_sb == send_blocks) {
        pf_sb = 0;
        pf_item += gridDim.x;
      }
    };
    load_q(pf_item, pf_sb, 0);
    advance_pf();
    load_q(pf_item, pf_sb, 1);
    advance_pf();
    uint32_t t = 0;
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
      for (int sb = 0; sb < send_blocks; ++sb, ++t) {
        const int st = t & 1;
        mbar_wait(&full[st], (t >> 1) & 1, err_flag);
        load_q(pf_item, pf_sb, st);  // producers are done with Q stage st: refill it for tile t + 2
        advance_pf();
        mbar_wait(dempty, (t & 1) ^ 1, err_flag);
        tc_fence_after();
        const uint64_t bst = bdesc0 + (uint64_t)(st * (64 >> 4));  // column half of the swizzled rows
#pragma unroll
        for (int s = 0; s < 2 * n / 16; ++s) {
          // rows [16 s, 16 s + 16) of the (s', dot) planes; 16 rows = 2048 bytes
          const uint64_t b_sd = bst + (uint64_t)(s * (2048 >> 4));
          if (elect_one()) {
            mma_ts(d0 + 0 * kCols, tm + 0 * n + s * 8, b_sd, idesc, s > 0);
            mma_ts(d0 + 1 * kCols, tm + 1 * n + s * 8, b_sd, idesc, s > 0);
            if (s < n / 16) {
              mma_ts(d0 + 2 * kCols, tm + 2 * n + s * 8, b_sd, idesc, s > 0);
#pragma unroll
              for (int k = 0; k < 3; ++k) {
                const uint64_t b_v = bst + (uint64_t)((((2 + k) * n + 16 * s) * 128) >> 4);
                mma_ts(d0 + (3 + k) * kCols, tm + 2 * n + n / 2 + s * 8, b_v, idesc, s > 0);
              }
            }
          }
        }
        if (elect_one()) {
          tc_commit(&empty[st]);
          tc_commit(dfull);
        }
        __syncwarp();
      }
    }
  } else {
    // ============================ epilogue: gate + aggregation ================================================
    if (warp >= kEpiWarp0 + NW && warp < kEpiWarp0 + 4) goto done;  // unused quadrant slots of group 0
    const int eg = (warp - kEpiWarp0) >> 2;   // column half
    const int quad = warp & 3;
    const int w = quad * 32 + lane;           // channel = TMEM lane
    const bool act = w < n;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const float b2s = act ? b2[w] : 0.f, b2g = act ? b2[n + w] : 0.f;
    float bm_s = 1.f, bm_v = 1.f, ba_s = 0.f;
    if (act && bn_mul != nullptr) {
      bm_s = bn_mul[w];
      bm_v = bn_mul[n + w];
      ba_s = bn_add[w];
    }
    uint32_t t = 0;
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
      const long long g = item / recv_blocks;
      const int i0 = (int)(item - g * recv_blocks) * kRecv;
      float acc[kRecv][4];
#pragma unroll
      for (int r = 0; r < kRecv; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

      for (int sb = 0; sb < send_blocks; ++sb, ++t) {
        const int slot = t & (kGeoSlots - 1);
        mbar_wait(&gfull[slot], (t >> 3) & 1, err_flag);  // acquires the producers' geometry writes
        mbar_wait(dfull, t & 1, err_flag);
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c0 = eg * 16 + h * 8;
          uint32_t d[6][8];
#pragma unroll
          for (int tile = 0; tile < 6; ++tile) SEGNN_TMEM_LD8(tmem + lane_base + kDBase + tile * kCols + c0, d[tile]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (h == 1) {  // accumulators are in registers: hand the TMEM tiles back to the MMA warp
            tc_fence_before();
            mbar_arrive(dempty);
          }
          if (act) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 ga = geoA[slot * kCols + c0 + j];
              const int r = j & 3;
              const float ms = silu_gate_fast(__uint_as_float(d[0][j]) + b2s);
              const float gt = sig_gate_fast(__uint_as_float(d[1][j]) + b2g) * ga.w;
              const float t1 = __uint_as_float(d[2][j]);
              acc[r][0] = fmaf(ga.w, ms, acc[r][0]);
              acc[r][1] = fmaf(gt, fmaf(ga.x, t1, __uint_as_float(d[3][j])), acc[r][1]);
              acc[r][2] = fmaf(gt, fmaf(ga.y, t1, __uint_as_float(d[4][j])), acc[r][2]);
              acc[r][3] = fmaf(gt, fmaf(ga.z, t1, __uint_as_float(d[5][j])), acc[r][3]);
            }
          }
        }
      }
      // combine the two column halves and write the receivers' aggregates (eval BatchNorm folded)
      if (eg == 1 && act) {
#pragma unroll
        for (int r = 0; r < kRecv; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) xch[(r * 4 + c) * n + w] = acc[r][c];
      }
      named_barrier(6, kEpiThreads);
      if (eg == 0 && act) {
#pragma unroll
        for (int r = 0; r < kRecv; ++r) {
          if (i0 + r < N) {
            float* o = agg + (g * N + i0 + r) * 4 * n;
            o[w] = fmaf(acc[r][0] + xch[(r * 4 + 0) * n + w], bm_s, ba_s);
            o[n + w] = (acc[r][1] + xch[(r * 4 + 1) * n + w]) * bm_v;
            o[2 * n + w] = (acc[r][2] + xch[(r * 4 + 2) * n + w]) * bm_v;
            o[3 * n + w] = (acc[r][3] + xch[(r * 4 + 3) * n + w]) * bm_v;
          }
        }
      }
      named_barrier(6, kEpiThreads);
    }
  }

done:
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

// ---- weight image: [128 lanes][3n columns] of bf16 pairs ------------------------------------------------------
__global__ void pack_w2_kernel(const float* __restrict__ ss, const float* __restrict__ vs,
                               const float* __restrict__ sv, const float* __restrict__ vv, int n,
                               uint32_t* __restrict__ out) {
  const int cols = 3 * n;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 128 * cols; idx += gridDim.x * blockDim.x) {
    const int w = idx / cols, c = idx - w * cols;
    float v[2] = {0.f, 0.f};
    if (w < n) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (c < 2 * n) {  // s tile (c < n) / g tile: K = [s' (n) | dot (n)], output column w or n + w
          const int oc = c < n ? w : n + w;
          const int k = 2 * (c < n ? c : c - n) + h;
          v[h] = k < n ? ss[(size_t)k * 2 * n + oc] : vs[(size_t)(k - n) * 2 * n + oc];
        } else if (c < 2 * n + n / 2) {  // T1 tile: K = s'
          const int k = 2 * (c - 2 * n) + h;
          v[h] = sv[(size_t)k * n + w];
        } else {  // D_k tiles: K = v'_k
          const int k = 2 * (c - 2 * n - n / 2) + h;
          v[h] = vv[(size_t)k * n + w];
        }
      }
    }
    const __nv_bfloat16 lo = __float2bfloat16(v[0]), hi = __float2bfloat16(v[1]);
    out[idx] = (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
  }
}

template <int NMUL>
static int launch_tc(const float* pos, const float* mass, int B, int N, const float* pp, const float* qq,
                     const float* w_edge1,
                     const float* b2, const void* w2_tc, const float* bn_mul, const float* bn_add, float* agg,
                     int* err_flag, cudaStream_t stream) {
  constexpr int threads = tc_num_warps<NMUL>() * 32;
  const size_t smem = 1024 + (size_t)5 * NMUL * 128 + (size_t)2 * kSend * 4 * 3 * NMUL * sizeof(float) +
                      (size_t)N * sizeof(float4) + kGeoSlots * kCols * (sizeof(float4) + sizeof(float2)) +
                      (size_t)kRecv * 4 * NMUL * sizeof(float) + 24 * sizeof(uint64_t) + 16;
  auto kern = edge_layer_tc_kernel<NMUL>;
  {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) {
      set_error("edge_layer_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(err));
      return SEGNN_E_CUDA;
    }
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long items = (long long)B * ((N + kRecv - 1) / kRecv);
  const unsigned grid = (unsigned)(items < sms ? items : sms);
  kern<<<grid, threads, smem, stream>>>(pos, mass, B, N, pp, qq, w_edge1, b2, (const uint32_t*)w2_tc, bn_mul, bn_add, agg,
                                        err_flag);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error("edge_layer_tc: launch: %s", cudaGetErrorString(err));
    return SEGNN_E_CUDA;
  }
  return SEGNN_OK;
}

}  // namespace tc

int edge_layer_tc(const float* pos, const float* mass, int B, int N, int n, const float* pp, const float* qq,
                  const float* w_edge1,
                  const float* b2, const void* w2_tc, const float* bn_mul, const float* bn_add, float* agg,
                  cudaStream_t stream) {
  if (N > tc::kMaxNodesPerGraph) {
    set_error("edge_layer_tc: N=%d nodes per graph exceeds the tensor-core kernel's staging limit (%d)", N,
              tc::kMaxNodesPerGraph);
    return SEGNN_E_UNSUPPORTED;
  }
  if (n == 32) return tc::launch_tc<32>(pos, mass, B, N, pp, qq, w_edge1, b2, w2_tc, bn_mul, bn_add, agg, nullptr, stream);
  if (n == 64) return tc::launch_tc<64>(pos, mass, B, N, pp, qq, w_edge1, b2, w2_tc, bn_mul, bn_add, agg, nullptr, stream);
  if (n == 96) return tc::launch_tc<96>(pos, mass, B, N, pp, qq, w_edge1, b2, w2_tc, bn_mul, bn_add, agg, nullptr, stream);
  set_error("edge_layer_tc: tensor-core mode is built for hidden multiplicity n in {32, 64, 96} (hidden_features "
            "64/128/192), got n=%d", n);
  return SEGNN_E_UNSUPPORTED;
}

int64_t pack_w2_tc(const float* ss, const float* vs, const float* sv, const float* vv, int n, void* out,
                   cudaStream_t stream) {
  if (n != 32 && n != 64 && n != 96) {
    set_error("segnn_pack_w2_tc: n must be 32, 64 or 96 (got %d)", n);
    return SEGNN_E_UNSUPPORTED;
  }
  const int64_t bytes = (int64_t)128 * 3 * n * 4;
  if (out == nullptr) return bytes;
  if (!ss || !vs || !sv || !vv) {
    set_error("segnn_pack_w2_tc: null weight block");
    return SEGNN_E_INVALID;
  }
  tc::pack_w2_kernel<<<(128 * 3 * n + 255) / 256, 256, 0, stream>>>(ss, vs, sv, vv, n, (uint32_t*)out);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error("segnn_pack_w2_tc: launch: %s", cudaGetErrorString(err));
    return SEGNN_E_CUDA;
  }
  return bytes;
}

}  // namespace segnn
