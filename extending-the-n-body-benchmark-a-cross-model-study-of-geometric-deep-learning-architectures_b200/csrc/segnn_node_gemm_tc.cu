// Node GEMM on tcgen05 (bf16 mode): y[node][c][:] = cat_K(x0[node][c], x1[node][c]) @ (c == 0 ? W_s : W_v).
//
// Rows (node, plane) are the MMA M dimension: a CTA owns 128-row tiles of ONE row class (scalar planes or vector
// planes; blockIdx.y) so that a single weight matrix stays resident in shared memory for the whole kernel.
//   A tile  [128 rows][K]   fp32 in HBM -> bf16, K-major, 128B swizzle, in shared memory (double buffered)
//   B       [n_out][K]      bf16 (pre-transposed weights), K-major, 128B swizzle, loaded once per CTA
//   D       [128][NC]       fp32 in TMEM, NC = output-column chunk (<= 192), double buffered
// Warp roles: 8 loader warps (consecutive lanes read consecutive 32-byte pieces of a row, 6 independent loads in flight
// each), 8 epilogue warps (TMEM lane quadrant = warp % 4, two warps per quadrant split the columns), 1 MMA warp.
// The kernel is HBM-bound (writes 4*n_out bytes per row against 2*K*n_out flops): the roofline is the copy rate.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "segnn_common.cuh"

namespace segnn {
namespace ngemm {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (int it = 0; it < (1 << 26); ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();  // protocol bug: fail loudly instead of hanging the GPU
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, SWIZZLE_128B: 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int N, int half) {
  uint32_t d = 0;
  d |= 1u << 4;   // D = f32
  if (!half) {
    d |= 1u << 7;   // A format: 0 = f16, 1 = bf16
    d |= 1u << 10;  // B format
  }
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(128 >> 4) << 24;
  return d;
}
__device__ __forceinline__ uint32_t pack_pair(float lo, float hi, int half) {
  uint32_t r;
  if (half)
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
#define SEGNN_NG_LD8(taddr, r)                                                                                     \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                            \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])     \
               : "r"(taddr))

#define SEGNN_NG_LD16(taddr, r)                                                                                    \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"((r)[0]), "=r"((r)[1]), "=r"((r)[2]), "=r"((r)[3]), "=r"((r)[4]), "=r"((r)[5]), "=r"((r)[6]),   \
                 "=r"((r)[7]), "=r"((r)[8]), "=r"((r)[9]), "=r"((r)[10]), "=r"((r)[11]), "=r"((r)[12]),              \
                 "=r"((r)[13]), "=r"((r)[14]), "=r"((r)[15])                                                         \
               : "r"(taddr))

constexpr int kLoadWarps = 8;
constexpr int kEpiWarps = 16;  // four per TMEM lane quadrant: the 32-column blocks of a chunk are dealt round-robin
constexpr int kThreads = (kLoadWarps + kEpiWarps + 1) * 32;  // 8 loader + 16 epilogue + 1 MMA warps
constexpr int kMmaWarp = kLoadWarps + kEpiWarps;

// rows of class 0: node r -> plane (r*4); class 1: row r -> node r/3, plane 1 + r%3
__device__ __forceinline__ long long plane_of(int cls, long long r) {
  return cls == 0 ? r * 4 : (r / 3) * 4 + 1 + (r % 3);
}
// pair16 output mode: four row classes (one per plane), row r of class c = node r, plane c
__device__ __forceinline__ long long plane_of4(int cls, long long r) { return r * 4 + cls; }

__global__ void __launch_bounds__(kThreads, 1)
    node_gemm_tc_kernel(const float* __restrict__ x0, const float* __restrict__ x1, int nodes, int n_in,
                        const __nv_bfloat16* __restrict__ wt_s, const __nv_bfloat16* __restrict__ wt_v,
                        const float* __restrict__ bias, int n_bias, int n_out, int nc, float* __restrict__ y0,
                        float* __restrict__ y1, int split, int ctas_cls0, int fp16_operands, int pair16, int in16) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int K = x1 ? 2 * n_in : n_in;
  const int katoms = (K + 63) / 64;
  const int a_bytes = katoms * 128 * 128;       // one A stage
  const int b_atom_bytes = n_out * 128;         // one K-atom of B
  uint8_t* sA = smem;                           // 2 stages
  uint8_t* sB = smem + 2 * a_bytes;             // katoms * n_out * 128
  float* sBias = reinterpret_cast<float*>(sB + katoms * b_atom_bytes);  // [n_out] (zeros past n_bias)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + n_out);
  uint64_t* afull = bars;       // [2] loaders -> MMA
  uint64_t* aempty = bars + 2;  // [2] MMA -> loaders
  uint64_t* dfull = bars + 4;   // [2] MMA -> epilogue
  uint64_t* dempty = bars + 6;  // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // the first ctas_cls0 CTAs own the scalar-plane rows, the rest the vector-plane rows (3x as many)
  // pair16: four row classes (class = plane, rows = nodes), CTAs dealt round-robin (the grid is a multiple of 4)
  const bool four = pair16 == 1;
  const int cls = four ? (int)(blockIdx.x & 3) : ((int)blockIdx.x < ctas_cls0 ? 0 : 1);
  const int cta = four ? (int)(blockIdx.x >> 2) : (cls == 0 ? blockIdx.x : blockIdx.x - ctas_cls0);
  const int cta_stride = four ? (int)(gridDim.x >> 2) : (cls == 0 ? ctas_cls0 : gridDim.x - ctas_cls0);
  const long long rows = (four || cls == 0) ? (long long)nodes : (long long)nodes * 3;
  const long long tiles = (rows + 127) / 128;
  const __nv_bfloat16* __restrict__ wt = cls == 0 ? wt_s : wt_v;
  const int nchunks = n_out / nc;

  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&afull[i], kLoadWarps * 32);
      mbar_init(&aempty[i], 1);
      mbar_init(&dfull[i], 1);
      mbar_init(&dempty[i], kEpiWarps);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // weights -> smem (bf16 [n_out][K] K-major in HBM): 16-byte pieces of 8 k
  {
    const int k8n = K / 8;
    for (int idx = tid; idx < n_out * k8n; idx += kThreads) {
      const int row = idx / k8n, k8 = idx - row * k8n;
      const uint4 v = *reinterpret_cast<const uint4*>(wt + (size_t)row * K + k8 * 8);
      *reinterpret_cast<uint4*>(sB + (k8 >> 3) * b_atom_bytes + row * 128 + (((k8 & 7) ^ (row & 7)) << 4)) = v;
    }
    for (int i = tid; i < n_out; i += kThreads) sBias[i] = (bias != nullptr && i < n_bias) ? bias[i] : 0.f;
    proxy_fence();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < kLoadWarps && in16) {
    // ===================== loaders, 16-bit rows: x0 / x1 already hold the operand format =====================
    // The producing kernels keep a 16-bit copy of the features next to the fp32 one (rounded exactly like the
    // conversion below would round them), so the tile is a list of 16-byte pieces (8 consecutive k of one row) that go
    // to their swizzled position unchanged.  A piece is 4 registers instead of 8: two batches of up to 6 pieces per
    // thread stay in flight, the loads of batch g + 1 are issued before batch g is stored, across tile boundaries
    // (the fp32 loader has one batch in flight and idles for a full memory latency per tile).
    const uint16_t* __restrict__ xh0 = reinterpret_cast<const uint16_t*>(x0);
    const uint16_t* __restrict__ xh1 = reinterpret_cast<const uint16_t*>(x1);
    const int k8 = K / 8;                       // pieces per row
    const int per_thread = (128 * k8) / 256;    // K % 16 == 0 -> exact
    const int nbpt = per_thread > 6 ? 2 : 1;    // batches per tile
    const int bs = (per_thread + nbpt - 1) / nbpt;
    const int row_step = 256 / k8, kp_step = 256 % k8;
    const uint32_t my_tiles = cta < tiles ? (uint32_t)((tiles - cta + cta_stride - 1) / cta_stride) : 0u;
    const uint32_t G = my_tiles * nbpt;
    auto issue = [&](uint32_t g, uint4 (&v)[6]) {
      const uint32_t t = nbpt == 2 ? g >> 1 : g;
      const int half = nbpt == 2 ? (int)(g & 1) : 0;
      const long long tile = cta + (long long)t * cta_stride;
      const int p0 = tid + 256 * half * bs;
      int row = p0 / k8, kp = p0 - row * k8;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        v[j] = make_uint4(0u, 0u, 0u, 0u);
        const long long gr = tile * 128 + row;
        if (j < bs && half * bs + j < per_thread && gr < rows) {
          const long long pl = four ? plane_of4(cls, gr) : plane_of(cls, gr);
          const int k = kp * 8;
          const uint16_t* src = k < n_in ? xh0 + pl * n_in + k : xh1 + pl * n_in + (k - n_in);
          asm volatile("ld.global.nc.v4.b32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(v[j].x), "=r"(v[j].y), "=r"(v[j].z), "=r"(v[j].w)
                       : "l"(src));
        }
        row += row_step;
        kp += kp_step;
        if (kp >= k8) {
          kp -= k8;
          ++row;
        }
      }
    };
    auto commit = [&](uint32_t g, const uint4 (&v)[6]) {
      const uint32_t t = nbpt == 2 ? g >> 1 : g;
      const int half = nbpt == 2 ? (int)(g & 1) : 0;
      const int ab = t & 1;
      if (half == 0) mbar_wait(&aempty[ab], ((t >> 1) & 1) ^ 1);
      uint8_t* dst = sA + ab * a_bytes;
      const int p0 = tid + 256 * half * bs;
      int row = p0 / k8, kp = p0 - row * k8;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        if (j < bs && half * bs + j < per_thread) {
          const int k = kp * 8;
          *reinterpret_cast<uint4*>(dst + (k >> 6) * (128 * 128) + row * 128 + ((((k & 63) >> 3) ^ (row & 7)) << 4)) = v[j];
        }
        row += row_step;
        kp += kp_step;
        if (kp >= k8) {
          kp -= k8;
          ++row;
        }
      }
      if (half == nbpt - 1) {
        proxy_fence();
        mbar_arrive(&afull[ab]);
      }
    };
    uint4 va[6], vb[6];
    if (G > 0) issue(0, va);
    for (uint32_t g = 0; g < G; g += 2) {
      if (g + 1 < G) issue(g + 1, vb);
      commit(g, va);
      if (g + 2 < G) issue(g + 2, va);
      if (g + 1 < G) commit(g + 1, vb);
    }
  } else if (warp < kLoadWarps) {
    // ===================== loaders: fp32 rows -> bf16 swizzled A tile =====================
    // The tile is a list of 32-byte pieces (8 consecutive k of one row); piece p = tid + 256 i, so consecutive lanes
    // read consecutive sectors of a row and the memory system sees full 128-byte requests (one row per lane made every
    // lane its own request).  (row, piece-in-row) advance incrementally: no division in the loop.
    const int k8 = K / 8;                       // pieces per row
    const int per_thread = (128 * k8) / 256;    // K % 16 == 0 -> exact
    const int row_step = 256 / k8, kp_step = 256 % k8;
    uint32_t t = 0;
    for (long long tile = cta; tile < tiles; tile += cta_stride, ++t) {
      const int ab = t & 1;
      mbar_wait(&aempty[ab], ((t >> 1) & 1) ^ 1);
      uint8_t* dst = sA + ab * a_bytes;
      int row = tid / k8, kp = tid - row * k8;
      for (int b0 = 0; b0 < per_thread; b0 += 6) {
        float v[6][8];
        int rr[6], kk[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          rr[i] = row;
          kk[i] = kp * 8;
          row += row_step;
          kp += kp_step;
          if (kp >= k8) {
            kp -= k8;
            ++row;
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) v[i][q] = 0.f;
          const long long gr = tile * 128 + rr[i];
          if (b0 + i < per_thread && gr < rows) {
            const long long pl = four ? plane_of4(cls, gr) : plane_of(cls, gr);
            const int k = kk[i];
            const float* src = k < n_in ? x0 + pl * n_in + k : x1 + pl * n_in + (k - n_in);
            asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(v[i][0]), "=f"(v[i][1]), "=f"(v[i][2]), "=f"(v[i][3]), "=f"(v[i][4]), "=f"(v[i][5]),
                           "=f"(v[i][6]), "=f"(v[i][7])
                         : "l"(src));
          }
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          if (b0 + i < per_thread) {
            const int k = kk[i], r = rr[i];  // 8 k = one 16-byte chunk of the swizzled row
            const uint4 o = make_uint4(pack_pair(v[i][0], v[i][1], fp16_operands), pack_pair(v[i][2], v[i][3], fp16_operands),
                                       pack_pair(v[i][4], v[i][5], fp16_operands), pack_pair(v[i][6], v[i][7], fp16_operands));
            *reinterpret_cast<uint4*>(dst + (k >> 6) * (128 * 128) + r * 128 + ((((k & 63) >> 3) ^ (r & 7)) << 4)) = o;
          }
        }
      }
      proxy_fence();
      mbar_arrive(&afull[ab]);
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc(nc, fp16_operands);
    const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
    uint32_t t = 0, dcount = 0;
    for (long long tile = cta; tile < tiles; tile += cta_stride, ++t) {
      const int ab = t & 1;
      mbar_wait(&afull[ab], (t >> 1) & 1);
      for (int c = 0; c < nchunks; ++c, ++dcount) {
        const int db = dcount & 1;
        mbar_wait(&dempty[db], ((dcount >> 1) & 1) ^ 1);
        tc_fence_after();
        if (lane == 0) {
          for (int s = 0; s < K / 16; ++s) {
            const uint32_t koff = (s >> 2) * (128 * 128) + (s & 3) * 32;
            const uint32_t boff = (s >> 2) * b_atom_bytes + (c * nc) * 128 + (s & 3) * 32;
            mma_ss(tmem + db * 256, make_desc(sA_addr + ab * a_bytes + koff), make_desc(sB_addr + boff), idesc, s > 0);
          }
          tc_commit(&dfull[db]);
          if (c == nchunks - 1) tc_commit(&aempty[ab]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue: TMEM -> HBM =====================
    // 8 warps, two per TMEM lane quadrant; the 32-column blocks of the accumulator chunks alternate between the two.
    // tcgen05.ld hands every thread one ROW, and storing from that layout makes every lane its own L2 request of one
    // 32-byte sector: the profile showed the L2 tag pipeline (one lookup per request) as the busiest unit and the
    // epilogue warps stalled behind their own stores.  Each quad of lanes therefore transposes its 4 rows x 4 sectors
    // with two rounds of xor-shuffles, after which lanes 4g..4g+3 hold the four consecutive sectors of ONE row: a
    // 256-bit store per lane then makes full 128-byte lines (4x fewer L2 requests), with no shared memory.
    const int e = warp - kLoadWarps;
    const int quad = e & 3, chalf = e >> 2;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const int q = lane & 3;
    const bool up2 = (q & 2) != 0, up1 = (q & 1) != 0;
    const int n_out1 = n_out - split;
    const int nb = nc / 32;  // 32-column blocks per chunk
    uint32_t dcount = 0;
    for (long long tile = cta; tile < tiles; tile += cta_stride) {
      // plane index of this lane's own row, then of the 4 rows of its quad (row base + m), -1 past the end
      const long long gr = tile * 128 + quad * 32 + lane;
      const long long pl_own = gr < rows ? (four ? plane_of4(cls, gr) : plane_of(cls, gr)) : -1;
      long long plm[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) plm[m] = __shfl_sync(0xffffffffu, pl_own, (lane & ~3) + m);
      for (int c = 0; c < nchunks; ++c, ++dcount) {
        const int db = dcount & 1;
        mbar_wait(&dfull[db], (dcount >> 1) & 1);
        tc_fence_after();
        bool arrived = false;
        for (int bk = 0; bk < nb; ++bk) {
          if (((c * nb + bk) & (kEpiWarps / 4 - 1)) != chalf) continue;
          uint32_t u[32];  // u[8 k + i]: sector k (columns 8k .. 8k+7 of the block) of this lane's row
          SEGNN_NG_LD16(tmem + lane_base + db * 256 + bk * 32, u);
          SEGNN_NG_LD16(tmem + lane_base + db * 256 + bk * 32 + 16, (u + 16));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (bk + kEpiWarps / 4 >= nb) {  // this warp's last block of the accumulator is in registers
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&dempty[db]);
            arrived = true;
          }
          const int col0 = c * nc + bk * 32;
          if (cls == 0 && col0 < n_bias) {  // warp-uniform; n_bias is a multiple of 32 here (checked by the launcher)
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4) {
              const float4 bv = *reinterpret_cast<const float4*>(sBias + col0 + 4 * q4);
              u[4 * q4 + 0] = __float_as_uint(__uint_as_float(u[4 * q4 + 0]) + bv.x);
              u[4 * q4 + 1] = __float_as_uint(__uint_as_float(u[4 * q4 + 1]) + bv.y);
              u[4 * q4 + 2] = __float_as_uint(__uint_as_float(u[4 * q4 + 2]) + bv.z);
              u[4 * q4 + 3] = __float_as_uint(__uint_as_float(u[4 * q4 + 3]) + bv.w);
            }
          }
          if (pair16 == 1) {
            // fp16 output, nodes interleaved in pairs: y[node / 2][plane][col][node & 1] (the layout the packed-half
            // edge kernel reads).  Lanes 2g, 2g+1 hold the two nodes of a pair (rows of a class are consecutive nodes,
            // tiles start at even rows): the even lane packs columns [0, 16) of both, the odd lane columns [16, 32),
            // 64 contiguous bytes each = one 128-byte line per pair of lanes.
            const bool odd = (lane & 1) != 0;
            uint32_t w16[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint32_t rcv = __shfl_xor_sync(0xffffffffu, odd ? u[i] : u[16 + i], 1);
              const float lo = __uint_as_float(odd ? rcv : u[i]), hi = __uint_as_float(odd ? u[16 + i] : rcv);
              asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w16[i]) : "f"(hi), "f"(lo));
            }
            if (pl_own >= 0) {
              const int col = col0 + (odd ? 16 : 0);
              const long long prow = (pl_own >> 3) * 4 + cls;  // (node / 2) * 4 + plane
              uint32_t* dst = col < split
                                  ? reinterpret_cast<uint32_t*>(y0) + prow * split + col
                                  : reinterpret_cast<uint32_t*>(y1) + prow * n_out1 + (col - split);
              asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(w16[0]), "r"(w16[1]),
                           "r"(w16[2]), "r"(w16[3]), "r"(w16[4]), "r"(w16[5]), "r"(w16[6]), "r"(w16[7])
                           : "memory");
              asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + 8), "r"(w16[8]), "r"(w16[9]),
                           "r"(w16[10]), "r"(w16[11]), "r"(w16[12]), "r"(w16[13]), "r"(w16[14]), "r"(w16[15])
                           : "memory");
            }
            continue;
          }
          // round 1 (lane ^ 2): slots {k, k + 2} -> slot = 2 * (row bit 1) + (sector bit 0), sector bit 1 = lane bit 1
#pragma unroll
          for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint32_t lo = u[8 * k + i], hi = u[8 * (k + 2) + i];
              const uint32_t rcv = __shfl_xor_sync(0xffffffffu, up2 ? lo : hi, 2);
              u[8 * k + i] = up2 ? rcv : lo;
              u[8 * (k + 2) + i] = up2 ? hi : rcv;
            }
          // round 2 (lane ^ 1): slots {2 rb, 2 rb + 1} -> slot = row offset m, sector = lane & 3
#pragma unroll
          for (int rb = 0; rb < 2; ++rb)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint32_t lo = u[8 * (2 * rb) + i], hi = u[8 * (2 * rb + 1) + i];
              const uint32_t rcv = __shfl_xor_sync(0xffffffffu, up1 ? lo : hi, 1);
              u[8 * (2 * rb) + i] = up1 ? rcv : lo;
              u[8 * (2 * rb + 1) + i] = up1 ? hi : rcv;
            }
          // u[8 m + i]: columns col0 + 8 q + i of row (quad base + m): lanes 4g..4g+3 write one 128-byte line
          const int col = col0 + 8 * q;
          if (pair16 == 2) {  // plain fp16 rows [plane][n_out] (outputs that only feed an attribute-combine pass)
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              if (plm[m] >= 0) {
                uint32_t h[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  asm("cvt.rn.f16x2.f32 %0, %1, %2;"
                      : "=r"(h[i])
                      : "f"(__uint_as_float(u[8 * m + 2 * i + 1])), "f"(__uint_as_float(u[8 * m + 2 * i])));
                __half* dst = reinterpret_cast<__half*>(y0) + plm[m] * n_out + col;
                asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3])
                             : "memory");
              }
            }
            continue;
          }
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            if (plm[m] >= 0) {
              float* dst = col < split ? y0 + plm[m] * split + col : y1 + plm[m] * n_out1 + (col - split);
              asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(u[8 * m + 0]),
                           "r"(u[8 * m + 1]), "r"(u[8 * m + 2]), "r"(u[8 * m + 3]), "r"(u[8 * m + 4]), "r"(u[8 * m + 5]),
                           "r"(u[8 * m + 6]), "r"(u[8 * m + 7])
                           : "memory");
            }
          }
        }
        if (!arrived && lane == 0) mbar_arrive(&dempty[db]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

__global__ void transpose_bf16_kernel(const float* __restrict__ w, int K, int n_out, int half,
                                      __nv_bfloat16* __restrict__ wt) {
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < K * n_out; idx += gridDim.x * blockDim.x) {
    const int o = idx / K, k = idx - o * K;
    if (half)
      reinterpret_cast<__half*>(wt)[idx] = __float2half_rn(w[(size_t)k * n_out + o]);
    else
      wt[idx] = __float2bfloat16(w[(size_t)k * n_out + o]);
  }
}

}  // namespace ngemm
}  // namespace segnn

using namespace segnn;

extern "C" {

int segnn_pack_node_weight_tc(const float* w, int K, int n_out, int operand, void* wt_bf16, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(w && wt_bf16 && K > 0 && n_out > 0, "bad arguments");
  SEGNN_CHECK_ARG(operand == SEGNN_OPERAND_BF16 || operand == SEGNN_OPERAND_FP16, "unknown operand format");
  ngemm::transpose_bf16_kernel<<<(K * n_out + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      w, K, n_out, operand == SEGNN_OPERAND_FP16 ? 1 : 0, (__nv_bfloat16*)wt_bf16);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

}  // extern "C"

static int node_gemm_tc_launch(const float* x0, const float* x1, int nodes, int n_in, const void* wt_s, const void* wt_v,
                               const float* bias, int n_bias, int n_out, float* y0, float* y1, int split, int operand,
                               int pair16, segnn_stream_t stream, int in16 = 0) {
  SEGNN_CHECK_ARG(nodes >= 0 && n_in >= 1 && n_out >= 1, "bad sizes");
  SEGNN_CHECK_ARG(operand == SEGNN_OPERAND_BF16 || operand == SEGNN_OPERAND_FP16, "unknown operand format");
  if (nodes == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(x0 && wt_s && wt_v && y0, "null pointer");
  const int K = x1 ? 2 * n_in : n_in;
  SEGNN_CHECK_ARG(n_in % 16 == 0 && K <= 192, "tensor-core node GEMM needs n_in % 16 == 0 and K <= 192");
  SEGNN_CHECK_ARG(n_out % 32 == 0, "tensor-core node GEMM needs n_out % 32 == 0");
  if (y1 == nullptr) split = n_out;
  SEGNN_CHECK_ARG(split % 8 == 0 && split > 0 && split <= n_out, "split must be a positive multiple of 8");
  SEGNN_CHECK_ARG(bias == nullptr || (n_bias >= 0 && n_bias <= n_out), "n_bias out of range");
  SEGNN_CHECK_ARG(((uintptr_t)y0 & 31) == 0 && ((uintptr_t)y1 & 31) == 0 && ((uintptr_t)x0 & 31) == 0 &&
                      ((uintptr_t)x1 & 31) == 0,
                  "inputs and outputs must be 32-byte aligned (256-bit loads and stores)");
  SEGNN_CHECK_ARG(n_out % 32 == 0 && split % 32 == 0 && n_bias % 32 == 0,
                  "tensor-core node GEMM needs n_out, split and n_bias to be multiples of 32");
  int nc = 0;  // accumulator chunk: its 32-column blocks alternate between the two epilogue warps of a lane quadrant
  for (int c = 192; c >= 32; c -= 32)
    if (n_out % c == 0) { nc = c; break; }
  const int katoms = (K + 63) / 64;
  const size_t smem = 1024 + (size_t)2 * katoms * 128 * 128 + (size_t)katoms * n_out * 128 +
                      (size_t)n_out * sizeof(float) + 8 * 8 + 16;
  if (smem > 227 * 1024) {
    set_error("segnn_node_gemm_tc: K=%d, n_out=%d needs %zu bytes of shared memory", K, n_out, smem);
    return SEGNN_E_UNSUPPORTED;
  }
  cudaError_t err = cudaFuncSetAttribute(ngemm::node_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
  if (err != cudaSuccess) {
    set_error("segnn_node_gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(err));
    return SEGNN_E_CUDA;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // one persistent CTA per SM: a quarter of them on the scalar-plane rows (1/4 of all rows), the rest on the
  // vector-plane rows
  const long long tiles0 = ((long long)nodes + 127) / 128, tiles1 = ((long long)nodes * 3 + 127) / 128;
  long long c0 = sms / 4 > 0 ? sms / 4 : 1;
  if (c0 > tiles0) c0 = tiles0;
  long long c1 = sms - c0;
  if (c1 > tiles1) c1 = tiles1;
  unsigned grid = (unsigned)(c0 + c1);
  if (pair16 == 2) SEGNN_CHECK_ARG(y1 == nullptr, "fp16 row output has no split");
  if (pair16 == 1) {  // four row classes of `nodes` rows each, CTAs dealt round-robin
    SEGNN_CHECK_ARG(nodes % 2 == 0, "pair-interleaved output needs an even node count (even graph size)");
    long long per_cls = sms / 4 > 0 ? sms / 4 : 1;
    if (per_cls > tiles0) per_cls = tiles0;
    grid = (unsigned)(4 * per_cls);
  }
  ngemm::node_gemm_tc_kernel<<<grid, ngemm::kThreads, smem, (cudaStream_t)stream>>>(
      x0, x1, nodes, n_in, (const __nv_bfloat16*)wt_s, (const __nv_bfloat16*)wt_v, bias, n_bias, n_out, nc, y0, y1,
      split, (int)c0, operand == SEGNN_OPERAND_FP16 ? 1 : 0, pair16, in16);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

extern "C" {

int segnn_node_gemm_tc(const float* x0, const float* x1, int nodes, int n_in, const void* wt_s, const void* wt_v,
                       const float* bias, int n_bias, int n_out, float* y0, float* y1, int split, int operand,
                       segnn_stream_t stream) {
  return node_gemm_tc_launch(x0, x1, nodes, n_in, wt_s, wt_v, bias, n_bias, n_out, y0, y1, split, operand, 0, stream);
}

int segnn_node_gemm_tc_out16(const float* x0, const float* x1, int nodes, int n_in, const void* wt_s, const void* wt_v,
                             int n_out, void* y, int operand, segnn_stream_t stream) {
  return node_gemm_tc_launch(x0, x1, nodes, n_in, wt_s, wt_v, nullptr, 0, n_out, (float*)y, nullptr, n_out, operand, 2,
                             stream);
}

int segnn_node_gemm_tc_pair16(const float* x0, const float* x1, int nodes, int n_in, const void* wt_s, const void* wt_v,
                              const float* bias, int n_bias, int n_out, void* y0, void* y1, int split, int operand,
                              segnn_stream_t stream) {
  return node_gemm_tc_launch(x0, x1, nodes, n_in, wt_s, wt_v, bias, n_bias, n_out, (float*)y0, (float*)y1, split, operand,
                             1, stream);
}

int segnn_node_gemm_tc_x16(const void* x0, const void* x1, int nodes, int n_in, const void* wt_s, const void* wt_v,
                           const float* bias, int n_bias, int n_out, void* y0, void* y1, int split, int operand,
                           int out_mode, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(out_mode == 1 || out_mode == 2, "out_mode: 1 = fp16 node pairs (pair16), 2 = fp16 rows (out16)");
  return node_gemm_tc_launch((const float*)x0, (const float*)x1, nodes, n_in, wt_s, wt_v, bias, n_bias, n_out,
                             (float*)y0, (float*)y1, split, operand, out_mode, stream, 1);
}

}  // extern "C"
