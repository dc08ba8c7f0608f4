// fp32-accurate GEMM on tcgen05: C[M][N] = A[M][K] * B[K][N], fp32 in and out, "3xTF32" error compensation.
//
// Why it exists: the generic-irreps path (lmax_h = 2, BASELINE configuration 3; O3TensorProduct.forward for arbitrary
// irreps) contracts the Clebsch-Gordan-expanded operand with the stacked path weights, a real dense contraction that
// must keep the 1e-5 parity budget of the fp32 mode.  Round 1 called a library SGEMM (CUDA cores) for it.  Here every
// fp32 operand value v is split into hi = tf32(v) and lo = tf32(v - hi), and the product is accumulated as
// hi*hi + hi*lo + lo*hi in fp32 in TMEM (kind::tf32 MMAs, K = 8 per instruction): the dropped lo*lo term is 2^-22 of
// the product, so the result is fp32-accurate (measured <= 2e-7 relative to a float64 GEMM, tests/test_gpu_gemm.py)
// at one third of the tf32 tensor rate, which is still several times the fp32 CUDA-core rate.
//
// Tiling: a CTA owns 128-row tiles of A (persistent over tiles) and one N block of <= 128 columns (blockIdx.y); K is
// streamed in chunks of 32 fp32 (= one 128-byte swizzle row) through a 3-stage ring:
//   A chunk  [128 rows][32]  fp32 in HBM -> (hi, lo) tf32, K-major, 128B swizzle, written by 8 loader warps
//   B chunk  [NP rows][32]   (hi, lo) images prepared once per call by gemm_prep_b_kernel (transposed, split, swizzled):
//                            one cp.async.bulk per chunk
//   D        [128][NP]       fp32 in TMEM, double buffered (2 x 128 columns)
// Warp roles: 8 loader warps, 1 MMA warp, 4 epilogue warps (TMEM lane quadrant = warp % 4).
#include "segnn_common.cuh"

namespace segnn {
namespace g3 {

constexpr int kBM = 128;      // rows per tile
constexpr int kKC = 32;       // fp32 elements of K per chunk (128 bytes)
constexpr int kStages = 3;
constexpr int kNMax = 128;    // columns per CTA
constexpr int kLoadWarps = 8;
constexpr int kEpiWarps = 4;
constexpr int kMmaWarp = kLoadWarps;
constexpr int kThreads = (kLoadWarps + 1 + kEpiWarps) * 32;
constexpr int kStgLd = 36;    // floats per row of an epilogue staging tile (32 + 4: 16-byte aligned rows)

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

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
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
// instruction descriptor, kind::tf32: D = f32, A = B = tf32, both K-major, M = 128
__device__ __forceinline__ uint32_t make_idesc(int N) {
  uint32_t d = 0;
  d |= 1u << 4;   // D format f32
  d |= 2u << 7;   // A format tf32
  d |= 2u << 10;  // B format tf32
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(kBM >> 4) << 24;
  return d;
}
__device__ __forceinline__ float tf32_rna(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

#define SEGNN_G3_LD16(taddr, r)                                                                                    \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"((r)[0]), "=r"((r)[1]), "=r"((r)[2]), "=r"((r)[3]), "=r"((r)[4]), "=r"((r)[5]), "=r"((r)[6]),   \
                 "=r"((r)[7]), "=r"((r)[8]), "=r"((r)[9]), "=r"((r)[10]), "=r"((r)[11]), "=r"((r)[12]),              \
                 "=r"((r)[13]), "=r"((r)[14]), "=r"((r)[15])                                                         \
               : "r"(taddr))

// B[K][N] (row-major, ldb) -> per N block and K chunk the two swizzled images the tensor core reads:
// out[(nb * chunks + c) * 2 + {hi, lo}][NP rows][128 bytes], row = output column, 16-byte piece p at (p ^ (row & 7)).
__global__ void gemm_prep_b_kernel(const float* __restrict__ B, int ldb, int K, int N, int NP, int chunks,
                                   float* __restrict__ out) {
  const int nblocks = (N + kNMax - 1) / kNMax;
  const long long total = (long long)nblocks * chunks * NP * kKC;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(idx % kKC);
    const int row = (int)((idx / kKC) % NP);
    const int c = (int)((idx / ((long long)kKC * NP)) % chunks);
    const int nb = (int)(idx / ((long long)kKC * NP * chunks));
    const int k = c * kKC + kk, col = nb * NP + row;
    const float v = (k < K && col < N) ? B[(long long)k * ldb + col] : 0.f;
    const float hi = tf32_rna(v), lo = tf32_rna(v - hi);
    const long long base = ((long long)(nb * chunks + c) * 2) * NP * kKC;
    const int off = row * kKC + ((((kk >> 2) ^ (row & 7)) << 2) | (kk & 3));
    out[base + off] = hi;
    out[base + (long long)NP * kKC + off] = lo;
  }
}

// NODE mode: the rows are the (node, plane) rows of planar node features [nodes][4][n_in] (segnn_node_gemm's contract):
// class 0 = the scalar plane of every node against W_s, class 1 = the three vector planes against W_v, both in one
// launch (the tile list is class 0's tiles followed by class 1's); K may be the concatenation of two sources (x0 | x1);
// the output goes to y0 [plane][split] | y1 [plane][n_out - split] with the bias on class 0.
struct NodeMap {
  const float *x0, *x1;
  int n_in;
  long long nodes;
  const float* bias;
  int n_bias, split;
  float *y0, *y1;
  long long img_stride;  // floats between the B images of the two classes
};

template <bool NODE>
__global__ void __launch_bounds__(kThreads, 1)
    gemm_tf32x3_kernel(const float* __restrict__ A, long long lda, long long M, int K, int N, int NP, int chunks,
                       const float* __restrict__ Bimg, float* __restrict__ C, long long ldc, const NodeMap nm) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int a_bytes = kBM * 128;      // one of (hi, lo)
  const int b_bytes = NP * 128;
  const int stage_bytes = 2 * a_bytes + 2 * b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * stage_bytes);
  uint64_t* full = bars;                 // [kStages] loaders (+ B copy) -> MMA
  uint64_t* empty = bars + kStages;      // [kStages] MMA -> loaders
  uint64_t* dfull = bars + 2 * kStages;  // [2] MMA -> epilogue
  uint64_t* dempty = dfull + 2;          // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nb = blockIdx.y;
  const int n0 = nb * NP;
  const int ncols = max(0, min(NP, N - n0));  // live columns of this block (<= NP)
  // NODE: tiles of class 0 (nodes rows) then tiles of class 1 (3 nodes rows)
  const long long tiles0 = NODE ? (nm.nodes + kBM - 1) / kBM : 0;
  const long long tiles = NODE ? tiles0 + (3 * nm.nodes + kBM - 1) / kBM : (M + kBM - 1) / kBM;

  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], kLoadWarps + 1);  // one arrival per loader warp + the expect_tx arrival of the B copy
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&dfull[i], 1);
      mbar_init(&dempty[i], kEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < kLoadWarps) {
    // ===================== loaders: fp32 rows -> (hi, lo) tf32 swizzled chunks; lane 0 of warp 0 copies B ==========
    // The global loads of a chunk are issued two chunks before it is converted (three register sets, the (tile, chunk)
    // sequence flattened so that the prefetch crosses tile boundaries): ~48 KB per SM in flight instead of 16 KB, which
    // is what lets a GEMM with K of one to six chunks stream its rows at HBM speed.
    const float* bsrc0 = Bimg + (long long)nb * chunks * 2 * NP * kKC;
    const bool vec_ok = (lda & 3) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0;
    const bool node_vec_ok = NODE && (nm.n_in & 7) == 0 && (reinterpret_cast<uintptr_t>(nm.x0) & 15) == 0 &&
                             (nm.x1 == nullptr || (reinterpret_cast<uintptr_t>(nm.x1) & 15) == 0);
    const long long my_tiles = tiles > blockIdx.x ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long total = my_tiles * chunks;
    constexpr int kPieces = (kBM * 4) / (kLoadWarps * 32);
    auto issue = [&](long long it, float4 (&r)[2 * kPieces]) {
      const long long tile = blockIdx.x + (it / chunks) * gridDim.x;
      const int k0 = (int)(it % chunks) * kKC;
#pragma unroll
      for (int i = 0; i < kPieces; ++i) {
        const int p = tid + i * (kLoadWarps * 32);
        const int row = p >> 2, q = p & 3;
        const int k = k0 + 8 * q;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
        if (NODE) {
          const bool c1 = tile >= tiles0;
          const long long gr = (c1 ? tile - tiles0 : tile) * kBM + row;
          if (gr < (c1 ? 3 * nm.nodes : nm.nodes)) {
            const long long plane = c1 ? (gr / 3) * 4 + 1 + gr % 3 : gr * 4;
            if (node_vec_ok && k + 8 <= K) {  // n_in % 8 == 0: a piece never straddles the two sources
              const float* src = k < nm.n_in ? nm.x0 + plane * nm.n_in + k : nm.x1 + plane * nm.n_in + (k - nm.n_in);
              a0 = *reinterpret_cast<const float4*>(src);
              a1 = *reinterpret_cast<const float4*>(src + 4);
            } else {
              float v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int kk = k + j;
                v[j] = kk >= K ? 0.f
                               : (kk < nm.n_in ? nm.x0[plane * nm.n_in + kk] : nm.x1[plane * nm.n_in + (kk - nm.n_in)]);
              }
              a0 = make_float4(v[0], v[1], v[2], v[3]);
              a1 = make_float4(v[4], v[5], v[6], v[7]);
            }
          }
          r[2 * i] = a0;
          r[2 * i + 1] = a1;
          continue;
        }
        const long long gr = tile * kBM + row;
        if (gr < M) {
          const float* src = A + gr * lda + k;
          if (vec_ok && k + 8 <= K) {
            a0 = *reinterpret_cast<const float4*>(src);
            a1 = *reinterpret_cast<const float4*>(src + 4);
          } else {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (k + j < K) ? src[j] : 0.f;
            a0 = make_float4(v[0], v[1], v[2], v[3]);
            a1 = make_float4(v[4], v[5], v[6], v[7]);
          }
        }
        r[2 * i] = a0;
        r[2 * i + 1] = a1;
      }
    };
    auto consume = [&](long long it, const float4 (&r)[2 * kPieces]) {
      const int s = (int)(it % kStages);
      const int c = (int)(it % chunks);
      mbar_wait(&empty[s], (uint32_t)((it / kStages) & 1) ^ 1);
      uint8_t* st = smem + s * stage_bytes;
      if (tid == 0) {
        const float* bsrc = bsrc0;
        if (NODE && blockIdx.x + (it / chunks) * gridDim.x >= tiles0) bsrc += nm.img_stride;  // W_v images
        const uint32_t bytes = 2u * (uint32_t)b_bytes;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(bytes)
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(st + 2 * a_bytes)),
                     "l"(bsrc + (long long)c * 2 * NP * kKC), "r"(bytes), "r"(smem_u32(&full[s]))
                     : "memory");
      }
#pragma unroll
      for (int i = 0; i < kPieces; ++i) {
        const int p = tid + i * (kLoadWarps * 32);
        const int row = p >> 2, q = p & 3;
        const float v[8] = {r[2 * i].x, r[2 * i].y, r[2 * i].z, r[2 * i].w,
                            r[2 * i + 1].x, r[2 * i + 1].y, r[2 * i + 1].z, r[2 * i + 1].w};
        // hi = tf32(v) rounded to nearest on the integer pipe (cvt.rna.tf32 is an XU-pipe instruction, 16 lanes per
        // clock and SM); lo = v - hi is exact and the tensor core ignores its low 13 mantissa bits (2^-22 of v)
        float hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          hi[j] = __uint_as_float((__float_as_uint(v[j]) + 0x1000u) & 0xFFFFE000u);
          lo[j] = v[j] - hi[j];
        }
        uint8_t* rowp = st + row * 128;
        const int c0 = ((2 * q) ^ (row & 7)) << 4, c1 = ((2 * q + 1) ^ (row & 7)) << 4;
        *reinterpret_cast<float4*>(rowp + c0) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(rowp + c1) = make_float4(hi[4], hi[5], hi[6], hi[7]);
        *reinterpret_cast<float4*>(rowp + a_bytes + c0) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<float4*>(rowp + a_bytes + c1) = make_float4(lo[4], lo[5], lo[6], lo[7]);
      }
      proxy_fence();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
    };
    float4 r0[2 * kPieces], r1[2 * kPieces], r2[2 * kPieces];
    if (0 < total) issue(0, r0);
    if (1 < total) issue(1, r1);
    for (long long it = 0; it < total; it += 3) {
      if (it + 2 < total) issue(it + 2, r2);
      consume(it, r0);
      if (it + 1 < total) {
        if (it + 3 < total) issue(it + 3, r0);
        consume(it + 1, r1);
      }
      if (it + 2 < total) {
        if (it + 4 < total) issue(it + 4, r1);
        consume(it + 2, r2);
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer: per chunk 4 K-steps x (hi*lo, lo*hi, hi*hi) =================================
    const uint32_t idesc = make_idesc(NP);
    uint32_t it = 0, dcount = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++dcount) {
      const int db = dcount & 1;
      mbar_wait(&dempty[db], ((dcount >> 1) & 1) ^ 1);
      for (int c = 0; c < chunks; ++c, ++it) {
        const int s = it % kStages;
        mbar_wait(&full[s], (it / kStages) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + s * stage_bytes);
          const uint32_t sa_lo = sa + a_bytes, sb = sa + 2 * a_bytes, sb_lo = sb + b_bytes;
          const uint32_t d = tmem + db * kNMax;
#pragma unroll
          for (int ks = 0; ks < kKC / 8; ++ks) {
            const uint32_t o = ks * 32;
            mma_tf32(d, make_desc(sa + o), make_desc(sb_lo + o), idesc, (c > 0 || ks > 0) ? 1u : 0u);
            mma_tf32(d, make_desc(sa_lo + o), make_desc(sb + o), idesc, 1u);
            mma_tf32(d, make_desc(sa + o), make_desc(sb + o), idesc, 1u);
          }
          tc_commit(&empty[s]);
          if (c == chunks - 1) tc_commit(&dfull[db]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers (lane = row) -> per-warp staging tile -> HBM ===================
    // A lane holds 32 consecutive columns of ITS row; storing them directly makes every lane of a store instruction hit
    // a different 128-byte line (32 partial-sector requests per instruction, the limiter of the first version).  The
    // 32 x 32 block goes through shared memory instead and leaves as 128-byte row segments, four rows per instruction.
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const bool vec_ok = (ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0 && (n0 & 3) == 0;
    float* stg = reinterpret_cast<float*>(smem + kStages * stage_bytes + 256) + quad * (32 * kStgLd);
    const int sub = lane >> 3, c4 = lane & 7;
    uint32_t dcount = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++dcount) {
      const int db = dcount & 1;
      mbar_wait(&dfull[db], (dcount >> 1) & 1);
      tc_fence_after();
      for (int col = 0; col < NP; col += 32) {
        uint32_t u[32];
        const bool two = col + 16 < NP;
        SEGNN_G3_LD16(tmem + lane_base + db * kNMax + col, u);
        if (two) SEGNN_G3_LD16(tmem + lane_base + db * kNMax + col + 16, (u + 16));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (col + 32 >= NP) {  // the last block of the accumulator is in registers
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&dempty[db]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j < 4 || two)
            *reinterpret_cast<float4*>(stg + lane * kStgLd + 4 * j) =
                make_float4(__uint_as_float(u[4 * j]), __uint_as_float(u[4 * j + 1]), __uint_as_float(u[4 * j + 2]),
                            __uint_as_float(u[4 * j + 3]));
        }
        __syncwarp();
        const int ccol = col + 4 * c4;
        if (ccol < ncols && (c4 < 4 || two)) {
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) {
            const int row = rr * 4 + sub;
            if (NODE) {
              const bool c1 = tile >= tiles0;
              const long long gr = (c1 ? tile - tiles0 : tile) * kBM + quad * 32 + row;
              if (gr < (c1 ? 3 * nm.nodes : nm.nodes)) {
                const long long plane = c1 ? (gr / 3) * 4 + 1 + gr % 3 : gr * 4;
                const float4 v = *reinterpret_cast<const float4*>(stg + row * kStgLd + 4 * c4);
                const float vv[4] = {v.x, v.y, v.z, v.w};
                const int c = n0 + ccol;  // global output column
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                  const int cc = c + jj;
                  if (ccol + jj < ncols) {
                    float o = vv[jj];
                    if (!c1 && nm.bias != nullptr && cc < nm.n_bias) o += nm.bias[cc];
                    if (cc < nm.split) nm.y0[plane * nm.split + cc] = o;
                    else nm.y1[plane * (N - nm.split) + (cc - nm.split)] = o;
                  }
                }
              }
              continue;
            }
            const long long gr = tile * kBM + quad * 32 + row;
            if (gr < M) {
              const float4 v = *reinterpret_cast<const float4*>(stg + row * kStgLd + 4 * c4);
              float* dst = C + gr * ldc + n0 + ccol;
              if (vec_ok && ccol + 4 <= ncols) {
                *reinterpret_cast<float4*>(dst) = v;
              } else {
                const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                  if (ccol + jj < ncols) dst[jj] = vv[jj];
              }
            }
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// N is split into ceil(N / 128) column blocks of equal width (a multiple of 16): N = 192 -> 2 x 96, not 128 + 64
static inline int padded_n(int N) {
  const int nblocks = (N + kNMax - 1) / kNMax;
  const int blk = (N + nblocks - 1) / nblocks;
  return (blk + 15) & ~15;
}

}  // namespace g3
}  // namespace segnn

using namespace segnn;

extern "C" {

int64_t segnn_gemm_tf32x3_workspace(int K, int N) {
  if (K < 1 || N < 1) return -1;
  const int chunks = (K + g3::kKC - 1) / g3::kKC;
  const int nblocks = (N + g3::kNMax - 1) / g3::kNMax;
  return (int64_t)nblocks * chunks * 2 * g3::padded_n(N) * g3::kKC * (int64_t)sizeof(float);
}

int segnn_gemm_tf32x3(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int K, int N, float* C,
                      int64_t ldc, float* workspace, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(M >= 0 && K >= 1 && N >= 1 && lda >= K && ldb >= N && ldc >= N, "bad sizes");
  if (M == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(A && B && C && workspace, "null pointer");
  SEGNN_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "workspace must be 16-byte aligned");
  const int chunks = (K + g3::kKC - 1) / g3::kKC;
  const int nblocks = (N + g3::kNMax - 1) / g3::kNMax;
  const int NP = g3::padded_n(N);
  cudaStream_t s = (cudaStream_t)stream;
  const long long prep_total = (long long)nblocks * chunks * NP * g3::kKC;
  g3::gemm_prep_b_kernel<<<(unsigned)((prep_total + 255) / 256 < 1184 ? (prep_total + 255) / 256 : 1184), 256, 0, s>>>(
      B, (int)ldb, K, N, NP, chunks, workspace);
  SEGNN_CHECK_LAUNCH();
  const size_t smem = 1024 + (size_t)g3::kStages * (2 * g3::kBM * 128 + 2 * NP * 128) + 256 +
                      (size_t)g3::kEpiWarps * 32 * g3::kStgLd * sizeof(float);
  cudaError_t err = cudaFuncSetAttribute(g3::gemm_tf32x3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) {
    set_error("segnn_gemm_tf32x3: cudaFuncSetAttribute(%zu bytes): %s", smem, cudaGetErrorString(err));
    return SEGNN_E_CUDA;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long tiles = (M + g3::kBM - 1) / g3::kBM;
  long long gx = sms / nblocks;
  if (gx < 1) gx = 1;
  if (gx > tiles) gx = tiles;
  dim3 grid((unsigned)gx, (unsigned)nblocks);
  g3::gemm_tf32x3_kernel<false><<<grid, g3::kThreads, smem, s>>>(A, (long long)lda, (long long)M, K, N, NP, chunks,
                                                                workspace, C, (long long)ldc, g3::NodeMap{});
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int64_t segnn_node_gemm_tf32x3_workspace(int K, int n_out) {
  const int64_t one = segnn_gemm_tf32x3_workspace(K, n_out);
  return one < 0 ? one : 2 * one;
}

int segnn_node_gemm_tf32x3(const float* x0, const float* x1, int nodes, int n_in, const float* w_s, const float* w_v,
                           const float* bias, int n_bias, int n_out, float* y0, float* y1, int split, float* workspace,
                           segnn_stream_t stream) {
  SEGNN_CHECK_ARG(nodes >= 0 && n_in >= 1 && n_out >= 1, "bad sizes");
  if (nodes == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(x0 && w_s && w_v && y0 && workspace, "null pointer");
  SEGNN_CHECK_ARG(bias == nullptr || (n_bias >= 0 && n_bias <= n_out), "n_bias out of range");
  SEGNN_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "workspace must be 16-byte aligned");
  if (y1 == nullptr) split = n_out;
  SEGNN_CHECK_ARG(split > 0 && split <= n_out, "split out of range");
  const int K = x1 ? 2 * n_in : n_in;
  const int chunks = (K + g3::kKC - 1) / g3::kKC;
  const int nblocks = (n_out + g3::kNMax - 1) / g3::kNMax;
  const int NP = g3::padded_n(n_out);
  const int64_t img = segnn_gemm_tf32x3_workspace(K, n_out) / (int64_t)sizeof(float);
  cudaStream_t s = (cudaStream_t)stream;
  const long long prep_total = (long long)nblocks * chunks * NP * g3::kKC;
  const unsigned prep_grid = (unsigned)((prep_total + 255) / 256 < 1184 ? (prep_total + 255) / 256 : 1184);
  g3::gemm_prep_b_kernel<<<prep_grid, 256, 0, s>>>(w_s, n_out, K, n_out, NP, chunks, workspace);
  g3::gemm_prep_b_kernel<<<prep_grid, 256, 0, s>>>(w_v, n_out, K, n_out, NP, chunks, workspace + img);
  SEGNN_CHECK_LAUNCH();
  const size_t smem = 1024 + (size_t)g3::kStages * (2 * g3::kBM * 128 + 2 * NP * 128) + 256 +
                      (size_t)g3::kEpiWarps * 32 * g3::kStgLd * sizeof(float);
  cudaError_t err = cudaFuncSetAttribute(g3::gemm_tf32x3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
  if (err != cudaSuccess) {
    set_error("segnn_node_gemm_tf32x3: cudaFuncSetAttribute(%zu bytes): %s", smem, cudaGetErrorString(err));
    return SEGNN_E_CUDA;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long tiles = ((long long)nodes + g3::kBM - 1) / g3::kBM + (3LL * nodes + g3::kBM - 1) / g3::kBM;
  long long gx = sms / nblocks;
  if (gx < 1) gx = 1;
  if (gx > tiles) gx = tiles;
  g3::NodeMap nm{x0, x1, n_in, (long long)nodes, bias, n_bias, split, y0, y1, (long long)img};
  dim3 grid((unsigned)gx, (unsigned)nblocks);
  g3::gemm_tf32x3_kernel<true><<<grid, g3::kThreads, smem, s>>>(nullptr, 0, 0, K, n_out, NP, chunks, workspace, nullptr, 0,
                                                               nm);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

}  // extern "C"
