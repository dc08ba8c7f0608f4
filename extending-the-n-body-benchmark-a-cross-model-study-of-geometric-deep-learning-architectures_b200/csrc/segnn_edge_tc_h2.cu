// K3, packed-half variant (SEGNN_MODE_FP16_PACKED): the fused SEGNN edge layer of segnn_edge_tc.cu with the
// message_layer_1 combine done in packed fp16 (HFMA2 over PAIRS OF SENDERS) on fp16 projections, and with the producer
// and epilogue roles on DIFFERENT warps (round 2).
//
//   * P, Q arrive as fp16 with the nodes interleaved in pairs (segnn_node_gemm_tc_pair16): one 32-bit word is the same
//     projection of two consecutive senders, i.e. a ready HFMA2 operand; one 64-bit load fetches (gate, vector) parts;
//   * the tile geometry for the producers is fp16 too, laid out so that one 128-bit load gives a thread the four
//     (receiver, sender-pair) units of 8 columns for one quantity;
//   * the producer math is HFMA2 on sender pairs (receiver-side values are broadcast pairs kept in registers), the gate
//     is one tanh.approx.f16x2 per two edges, and the result IS the fp16 B operand: no conversions, no packing.
//
// Warp roles (16 warps, warp % 4 = SM sub-partition = TMEM lane quadrant; at n = 96 quadrants 0..2 hold channels):
//   per channel quadrant: 2 PRODUCER warps (receiver pair rp = 0 / 1, all 8 senders of the tile: 16 columns each) and
//   2 EPILOGUE warps (same split); quadrant 3: the MMA / copy warp and 3 scalar-channel producers.
// Why split (profiles/r2_k3_split_*): with both roles on every compute warp (round 1, kept as
// experiments/segnn_edge_tc_h2_v15.cu.txt) the four warps of a sub-partition ran the producer phase together (FMA pipe
// saturated by half-rate HFMA2, 960 clk per tile) and then the epilogue phase together (MUFU saturated by 16 tanh per
// thread, 510 clk), one after the other, because all of them meet at the single-buffered accumulators (dfull).
// Producers never touch TMEM, so on their own warps they are gated only by the operand stages (qfull / empty) and run
// up to two tiles ahead, overlapping their FMA-pipe work with the epilogue warps' MUFU work.  A warp owns a receiver
// pair over ALL senders, so the per-item exchange between sender quads (shared memory + named barrier) is gone too.
// The tile geometry (unit vectors, |r|, m_i m_j, validity) is produced by one epilogue warp five tiles ahead
// (ordering: written before that warp's dempty(t + 1) arrival, which the MMA warp observes before it arms qfull(t + 5)).
//
// Tile column order: c = 8 gi + 4 rl + sl with gi = 2 rp + sq (group), receiver = 2 rp + rl, sender = 4 sq + sl, so that
// two adjacent columns are two consecutive senders of one receiver (one half2) and a group's 8 columns are one 16-byte
// chunk of every B row.
//
// Precision: every producer operation rounds to fp16 (11-bit mantissa); the factor 1/2 of the gate pre-activations is
// folded into the node-GEMM weights.  Measured per-layer error vs the float64 oracle: see tests/test_gpu_parity.py.
#include "segnn_edge_tc_common.cuh"

#undef K3_TRACE
#define K3_TRACE(ev, t) \
  do {                  \
  } while (0)

namespace segnn {
namespace tc {

__device__ __forceinline__ uint32_t h2add(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t h2mul(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t h2fma(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t h2tanh(uint32_t a) {
  uint32_t d;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
__device__ __forceinline__ uint32_t h2dup_lo(uint32_t x) { return __byte_perm(x, x, 0x1010); }
__device__ __forceinline__ uint32_t h2dup_hi(uint32_t x) { return __byte_perm(x, x, 0x3232); }
__device__ __forceinline__ uint32_t h2bc(float x) {  // (x, x) as fp16 pair
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(x), "f"(x));
  return d;
}

template <int NMUL, bool MOMENTS>
__global__ void __launch_bounds__(kWarps * 32, 1)
    edge_layer_h2_kernel(const float* __restrict__ pos, const float* __restrict__ mass, int B, int N,
                         const uint32_t* __restrict__ pp, const uint32_t* __restrict__ qq,
                         const float* __restrict__ w_edge1, const float* __restrict__ b2,
                         const uint32_t* __restrict__ w2_tc, const float* __restrict__ bn_mul,
                         const float* __restrict__ bn_add, float* __restrict__ agg, float* __restrict__ moments,
                         __half* __restrict__ agg16) {
  constexpr int n = NMUL;
  constexpr int NW = n / 32;
  static_assert(NW >= 1 && NW <= 3, "warp % 4 == 3 hosts the MMA / scalar-producer warps");
  constexpr int kScalarWarps = NW;
  constexpr int kMmaWarp = 3;
  constexpr int kGeoWarp = 8;                      // epilogue warp (quadrant 0, receiver pair 0) that also owns the geometry
  constexpr int kGeoAhead = 5;                     // geometry of tile t + 5 is written during epilogue(t)
  constexpr int n3 = 3 * n;
  constexpr int kPairRow = 4 * n3;                 // words of one node pair: [plane][3n] half2
  constexpr int kQStageWords = (kSend / 2) * kPairRow;
  constexpr int kWeightCols = 3 * n;
  constexpr int kDBase = kWeightCols;
  constexpr int kGeo32 = 4 * kCols;                // fp32 words per slot: ax, ay, az, valid (epilogue)
  constexpr int kGeo16 = 5 * kCols / 2;            // words per slot: ax, ay, az, len, mm as fp16 (producers)

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sB = smem;                                                  // 5n rows x 128 bytes (two 64-byte stages)
  uint32_t* sQ = reinterpret_cast<uint32_t*>(smem + 5 * n * 128);      // [2][4 sender pairs][4 planes][3n] half2
  uint32_t* sP = sQ + 2 * kQStageWords;                                // [2 receiver pairs][4 planes][3n] half2
  float* geo32 = reinterpret_cast<float*>(sP + (kRecv / 2) * kPairRow);  // [slots][4][32]
  uint32_t* geo16 = reinterpret_cast<uint32_t*>(geo32 + kGeoSlots * kGeo32);  // [slots][5][16] half2
  uint64_t* bars = reinterpret_cast<uint64_t*>(geo16 + kGeoSlots * kGeo16);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  uint64_t* full = bars;
  uint64_t* empty = bars + 2;
  uint64_t* dfull = bars + 4;
  uint64_t* dempty = bars + 5;
  uint64_t* qfull = bars + 6;
  uint64_t* pfull = bars + 16;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int cb = warp & 3, grp = warp >> 2;
  const bool is_compute = cb < NW;
  const bool is_producer = is_compute && grp < 2;
  const bool is_epilogue = is_compute && grp >= 2;
  const bool is_scalar = cb == 3 && grp >= 1 && grp - 1 < NW;

  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 2 * NW + kScalarWarps);
      mbar_init(&empty[i], 1);
      mbar_init(&qfull[i], 1);
    }
    mbar_init(dfull, 1);
    mbar_init(pfull, 1);
    mbar_init(dempty, 2 * NW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tmem != 0u) __trap();

  const int recv_blocks = (N + kRecv - 1) / kRecv;
  const int send_blocks = (N + kSend - 1) / kSend;
  const int items = B * recv_blocks;
  auto advance = [&](TileCursor& c) {
    ++c.t;
    if (++c.sb == send_blocks) {
      c.sb = 0;
      c.item += gridDim.x;
    }
  };
  auto locate = [&](TileCursor& c) {
    c.g = c.item / recv_blocks;
    c.i0 = (c.item - c.g * recv_blocks) * kRecv;
  };

  // ---- geometry producer state (used by warp kGeoWarp only; lane = column c = 8 gi + 4 rl + sl) ----------------
  const int ggi = lane >> 3, grl = (lane >> 2) & 1, gsl_ = lane & 3;
  const int gr = 2 * (ggi >> 1) + grl, gsl = 4 * (ggi & 1) + gsl_;
  float gsx = 0.f, gsy = 0.f, gsz = 0.f, gsm = 0.f, grx = 0.f, gry = 0.f, grz = 0.f, grm = 0.f;
  bool gvalid = false, rvalid = false;
  const float* pos_g = pos;
  const float* mass_g = mass;
  TileCursor gp{(int)blockIdx.x, 0, 0u, 0, 0};
  auto enter_geo = [&](const TileCursor& c) {
    pos_g = pos + (long long)c.g * N * 3;
    mass_g = mass + (long long)c.g * N;
    if (c.item < items) {
      const int ii = c.i0 + gr, is = min(ii, N - 1);
      grx = pos_g[is * 3 + 0];
      gry = pos_g[is * 3 + 1];
      grz = pos_g[is * 3 + 2];
      grm = mass_g[is];
      rvalid = ii < N;
    }
  };
  auto tile_load = [&](const TileCursor& c) {
    if (c.item < items) {
      const int jj = c.sb * kSend + gsl;
      const int js = min(jj, N - 1);
      gsx = pos_g[js * 3 + 0];
      gsy = pos_g[js * 3 + 1];
      gsz = pos_g[js * 3 + 2];
      gsm = mass_g[js];
      gvalid = rvalid && (jj < N) && (jj != c.i0 + gr);
    }
  };
  auto tile_geometry = [&](const TileCursor& c) {
    if (c.item < items) {
      const int slot = c.t & (kGeoSlots - 1);
      const float dx = gsx - grx, dy = gsy - gry, dz = gsz - grz;
      const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
      const float inv = rsqrtf(fmaxf(d2, 1e-24f));
      const float sc = kY1 * inv;
      float* g32 = geo32 + slot * kGeo32;
      g32[0 * kCols + lane] = sc * dx;
      g32[1 * kCols + lane] = sc * dy;
      g32[2 * kCols + lane] = sc * dz;
      g32[3 * kCols + lane] = gvalid ? 1.0f : 0.0f;
      __half* g16 = reinterpret_cast<__half*>(geo16 + slot * kGeo16);
      g16[0 * kCols + lane] = __float2half_rn(sc * dx);
      g16[1 * kCols + lane] = __float2half_rn(sc * dy);
      g16[2 * kCols + lane] = __float2half_rn(sc * dz);
      g16[3 * kCols + lane] = __float2half_rn(d2 * inv);
      g16[4 * kCols + lane] = __float2half_rn(gsm * grm);
    }
    __syncwarp();
  };
  auto advance_geo = [&](TileCursor& c) {
    ++c.t;
    if (++c.sb == send_blocks) {
      c.sb = 0;
      c.item += gridDim.x;
      locate(c);
      enter_geo(c);
    }
  };

  // ---- message_layer_2 weights -> TMEM (identical to segnn_edge_tc.cu) ----------------------------------------
  if (is_compute && grp == 0) {
    const int row = cb * 32 + lane;
    const uint32_t lane_base = (uint32_t)(cb * 32) << 16;
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
  for (int i = tid; i < 2 * kQStageWords + (kRecv / 2) * kPairRow; i += kWarps * 32) sQ[i] = 0u;  // finite padding
  if (warp == kGeoWarp) {  // geometry of the first kGeoAhead tiles, positions of the next one in registers
    locate(gp);
    enter_geo(gp);
    for (int i = 0; i < kGeoAhead; ++i) {
      tile_load(gp);
      tile_geometry(gp);
      advance_geo(gp);
    }
    tile_load(gp);
  }
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  uint32_t bar0 = smem_u32(bars);
  asm volatile("" : "+r"(bar0));  // opaque: otherwise the address is rebuilt from S2R at every wait (round-1 SASS)

  if (is_producer) {
    // ============================ producer warps: gate / vector channels of 16 columns ==========================
    const int w = cb * 32 + lane;
    const int rp = grp & 1;
    // per-edge scalar terms of message_layer_1 (|r| and m_i m_j) as fp16 pairs; gate parts carry the factor 1/2
    const uint32_t wd0g = h2bc(0.5f * w_edge1[n + w]), wm0g = h2bc(0.5f * w_edge1[3 * n + w]);
    const uint32_t wd1 = h2bc(w_edge1[4 * n + w]), wm1 = h2bc(w_edge1[5 * n + w]);
    // receiver-side projections of the warp's two receivers as broadcast pairs: [rl][plane]
    uint32_t Pg[2][4], Pv[2][4];
    uint32_t p_items = 0;
#pragma unroll 1
    for (TileCursor cur{(int)blockIdx.x, 0, 0u, 0, 0}; cur.item < items; advance(cur)) {
      if (cur.sb == 0) {
        mbar_wait_a(bar0 + 8 * 16, p_items & 1);  // pfull
        ++p_items;
        const uint32_t* pr = sP + rp * kPairRow;  // the pair (2 rp, 2 rp + 1); rows past the graph end: finite, masked
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint2 gv = *reinterpret_cast<const uint2*>(pr + c * n3 + n + 2 * w);
          Pg[0][c] = h2dup_lo(gv.x);
          Pg[1][c] = h2dup_hi(gv.x);
          Pv[0][c] = h2dup_lo(gv.y);
          Pv[1][c] = h2dup_hi(gv.y);
        }
      }
      const uint32_t t = cur.t;
      const int st = t & 1, slot = t & (kGeoSlots - 1);
      mbar_wait_a(bar0 + 8 * (6 + st), (t >> 1) & 1);        // qfull[st]: sender rows + geometry
      mbar_wait_a(bar0 + 8 * (2 + st), ((t >> 1) & 1) ^ 1);  // empty[st]: MMAs of tile t - 2 have read the stage
      const uint32_t* qs = sQ + st * kQStageWords;
#pragma unroll
      for (int sq = 0; sq < 2; ++sq) {
        const int gi = rp * 2 + sq;
        const uint32_t* gw = geo16 + slot * kGeo16 + 4 * gi;
        // G[a].{x,y,z,w} = units (rl 0, u 0), (rl 0, u 1), (rl 1, u 0), (rl 1, u 1) of quantity a
        const uint4 GX = *reinterpret_cast<const uint4*>(gw + 0 * (kCols / 2));
        const uint4 GY = *reinterpret_cast<const uint4*>(gw + 1 * (kCols / 2));
        const uint4 GZ = *reinterpret_cast<const uint4*>(gw + 2 * (kCols / 2));
        const uint4 GL = *reinterpret_cast<const uint4*>(gw + 3 * (kCols / 2));
        const uint4 GM = *reinterpret_cast<const uint4*>(gw + 4 * (kCols / 2));
        uint32_t packed[4][5];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const uint32_t* qr = qs + (2 * sq + u) * kPairRow;  // the sender pair (4 sq + 2 u, 4 sq + 2 u + 1)
          uint32_t Qg[4], Qv[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint2 gv = *reinterpret_cast<const uint2*>(qr + c * n3 + n + 2 * w);
            Qg[c] = gv.x;
            Qv[c] = gv.y;
          }
#pragma unroll
          for (int rl = 0; rl < 2; ++rl) {
            const int un = rl * 2 + u;
            const uint32_t ax = un == 0 ? GX.x : un == 1 ? GX.y : un == 2 ? GX.z : GX.w;
            const uint32_t ay = un == 0 ? GY.x : un == 1 ? GY.y : un == 2 ? GY.z : GY.w;
            const uint32_t az = un == 0 ? GZ.x : un == 1 ? GZ.y : un == 2 ? GZ.z : GZ.w;
            const uint32_t le = un == 0 ? GL.x : un == 1 ? GL.y : un == 2 ? GL.z : GL.w;
            const uint32_t mm = un == 0 ? GM.x : un == 1 ? GM.y : un == 2 ? GM.z : GM.w;
            // half gate pre-activation: (P0 + Q0) + a.(P0k + Q0k) + |r| wd + m_i m_j wm  (all with the factor 1/2)
            uint32_t hg = h2add(Pg[rl][0], Qg[0]);
            hg = h2fma(ax, h2add(Pg[rl][1], Qg[1]), hg);
            hg = h2fma(ay, h2add(Pg[rl][2], Qg[2]), hg);
            hg = h2fma(az, h2add(Pg[rl][3], Qg[3]), hg);
            hg = h2fma(le, wd0g, hg);
            hg = h2fma(mm, wm0g, hg);
            uint32_t tt = h2add(Pv[rl][0], Qv[0]);
            tt = h2fma(le, wd1, tt);
            tt = h2fma(mm, wm1, tt);
            const uint32_t zx = h2fma(ax, tt, h2add(Pv[rl][1], Qv[1]));
            const uint32_t zy = h2fma(ay, tt, h2add(Pv[rl][2], Qv[2]));
            const uint32_t zz = h2fma(az, tt, h2add(Pv[rl][3], Qv[3]));
            const uint32_t tg = h2tanh(hg);
            const uint32_t vx = h2fma(tg, zx, zx);  // 2 sigmoid(z_g) z_v
            const uint32_t vy = h2fma(tg, zy, zy);
            const uint32_t vz = h2fma(tg, zz, zz);
            uint32_t dt = h2mul(ax, vx);
            dt = h2fma(ay, vy, dt);
            dt = h2fma(az, vz, dt);
            packed[un][1] = dt;
            packed[un][2] = vx;
            packed[un][3] = vy;
            packed[un][4] = vz;
          }
        }
        const int chunk = st * 4 + gi;
#pragma unroll
        for (int p = 1; p < 5; ++p) {
          const int row = p * n + w;
          *reinterpret_cast<uint4*>(sB + row * 128 + ((chunk ^ (row & 7)) << 4)) =
              make_uint4(packed[0][p], packed[1][p], packed[2][p], packed[3][p]);
        }
      }
      proxy_fence();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(bar0 + 8 * st);  // full[st]
    }
  } else if (is_epilogue) {
    // ============================ epilogue warps: gate + sum over the senders of 16 columns =====================
    const int w = cb * 32 + lane;
    const int rp = grp & 1;
    const uint32_t lane_base = (uint32_t)(cb * 32) << 16;
    const float2 b2s = bc2(0.5f * b2[w]), b2g = bc2(0.5f * b2[n + w]);
    float sc_s = kCSilu, sc_v = 0.5f * kCSig, add_s = 0.f;
    if (bn_mul != nullptr) {
      sc_s *= bn_mul[w];
      sc_v *= bn_mul[n + w];
      add_s = bn_add[w];
    }
    float2 acc[2][4];  // [rl][component]: two partial sums (even / odd senders)
    // MOMENTS (train-mode BatchNorm statistics, trainer.py:373 evaluates with batch statistics): per receiver the sums
    // over its senders of m_s^2 and |m_v|^2 of the raw messages, [rl][scalar / vector]
    float2 m2[2][2];
#pragma unroll
    for (int rl = 0; rl < 2; ++rl) {
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[rl][c] = make_float2(0.f, 0.f);
      m2[rl][0] = m2[rl][1] = make_float2(0.f, 0.f);
    }
#pragma unroll 1
    for (TileCursor cur{(int)blockIdx.x, 0, 0u, 0, 0}; cur.item < items; advance(cur)) {
      const uint32_t t = cur.t;
      const int slot = t & (kGeoSlots - 1);
      mbar_wait_a(bar0 + 8 * 4, t & 1);  // dfull
      tc_fence_after();
#pragma unroll
      for (int sq = 0; sq < 2; ++sq) {
        const int gi = rp * 2 + sq;
        const float* gs = geo32 + slot * kGeo32 + 8 * gi;
        uint32_t d[6][8];
#pragma unroll
        for (int tile = 0; tile < 6; ++tile) SEGNN_TMEM_LD8(tmem + lane_base + kDBase + tile * kCols + 8 * gi, d[tile]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (sq == 1) {  // both column groups of the warp are in registers: the accumulators may be overwritten
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_a(bar0 + 8 * 5);  // dempty
        }
#pragma unroll
        for (int rl = 0; rl < 2; ++rl) {
          const float4 AX = *reinterpret_cast<const float4*>(gs + 0 * kCols + 4 * rl);
          const float4 AY = *reinterpret_cast<const float4*>(gs + 1 * kCols + 4 * rl);
          const float4 AZ = *reinterpret_cast<const float4*>(gs + 2 * kCols + 4 * rl);
          const float4 VA = *reinterpret_cast<const float4*>(gs + 3 * kCols + 4 * rl);
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int j = 4 * rl + 2 * e;  // columns (rl, sl = 2 e) and (rl, sl = 2 e + 1)
            const float2 ax = e ? hi2(AX) : lo2(AX), ay = e ? hi2(AY) : lo2(AY), az = e ? hi2(AZ) : lo2(AZ);
            const float2 va = e ? hi2(VA) : lo2(VA);
            const float2 ys = __fadd2_rn(u2f2(d[0][j], d[0][j + 1]), b2s);
            const float2 yg = __fadd2_rn(u2f2(d[1][j], d[1][j + 1]), b2g);
            const float2 ts = tanh2(ys), tg = tanh2(yg);
            const float2 ms = __ffma2_rn(ys, ts, ys);
            const float2 g1 = __ffma2_rn(va, tg, va);
            const float2 t1 = u2f2(d[2][j], d[2][j + 1]);
            if (MOMENTS) {
              const float2 msv = __fmul2_rn(va, ms);
              const float2 mx = __fmul2_rn(g1, __ffma2_rn(ax, t1, u2f2(d[3][j], d[3][j + 1])));
              const float2 my = __fmul2_rn(g1, __ffma2_rn(ay, t1, u2f2(d[4][j], d[4][j + 1])));
              const float2 mz = __fmul2_rn(g1, __ffma2_rn(az, t1, u2f2(d[5][j], d[5][j + 1])));
              acc[rl][0] = __fadd2_rn(acc[rl][0], msv);
              acc[rl][1] = __fadd2_rn(acc[rl][1], mx);
              acc[rl][2] = __fadd2_rn(acc[rl][2], my);
              acc[rl][3] = __fadd2_rn(acc[rl][3], mz);
              m2[rl][0] = __ffma2_rn(msv, msv, m2[rl][0]);
              m2[rl][1] = __ffma2_rn(mx, mx, __ffma2_rn(my, my, __ffma2_rn(mz, mz, m2[rl][1])));
            } else {
              acc[rl][0] = __ffma2_rn(va, ms, acc[rl][0]);
              acc[rl][1] = __ffma2_rn(g1, __ffma2_rn(ax, t1, u2f2(d[3][j], d[3][j + 1])), acc[rl][1]);
              acc[rl][2] = __ffma2_rn(g1, __ffma2_rn(ay, t1, u2f2(d[4][j], d[4][j + 1])), acc[rl][2]);
              acc[rl][3] = __ffma2_rn(g1, __ffma2_rn(az, t1, u2f2(d[5][j], d[5][j + 1])), acc[rl][3]);
            }
          }
        }
      }
      if (cur.sb == send_blocks - 1) {
        // item complete: the warp holds the sums over all senders of its two receivers
        const long long g = cur.item / recv_blocks;
        const int i0 = (int)(cur.item - g * recv_blocks) * kRecv;
#pragma unroll
        for (int rl = 0; rl < 2; ++rl) {
          const int r = 2 * rp + rl;
          if (i0 + r < N) {
            const float o0 = fmaf(acc[rl][0].x + acc[rl][0].y, sc_s, add_s);
            const float o1 = (acc[rl][1].x + acc[rl][1].y) * sc_v, o2 = (acc[rl][2].x + acc[rl][2].y) * sc_v,
                        o3 = (acc[rl][3].x + acc[rl][3].y) * sc_v;
            if (agg16 != nullptr) {  // fp16 rows for segnn_node_gemm_tc_x16 (the rounding its fp32 loader applies)
              __half* o = agg16 + (g * N + i0 + r) * 4 * n;
              o[w] = __float2half_rn(o0);
              o[n + w] = __float2half_rn(o1);
              o[2 * n + w] = __float2half_rn(o2);
              o[3 * n + w] = __float2half_rn(o3);
            } else {
              float* o = agg + (g * N + i0 + r) * 4 * n;
              o[w] = o0;
              o[n + w] = o1;
              o[2 * n + w] = o2;
              o[3 * n + w] = o3;
            }
            if (MOMENTS) {  // of the raw messages (before any BatchNorm affine), like segnn_edge_fp32.cu
              float* mo = moments + (g * N + i0 + r) * 2 * n;
              mo[w] = (m2[rl][0].x + m2[rl][0].y) * (kCSilu * kCSilu);
              mo[n + w] = (m2[rl][1].x + m2[rl][1].y) * (0.25f * kCSig * kCSig);
            }
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[rl][c] = make_float2(0.f, 0.f);
          m2[rl][0] = m2[rl][1] = make_float2(0.f, 0.f);
        }
      }
      if (warp == kGeoWarp) {  // geometry of tile t + kGeoAhead, positions of tile t + kGeoAhead + 1
        tile_geometry(gp);
        advance_geo(gp);
        tile_load(gp);
      }
    }
  } else if (is_scalar) {
    // ============================ scalar-channel producers (4th SM sub-partition) ===============================
    const int w = (grp - 1) * 32 + lane;
    const uint32_t wd0s = h2bc(0.5f * w_edge1[w]), wm0s = h2bc(0.5f * w_edge1[2 * n + w]);
    uint32_t Ps[4][4];  // [receiver][plane] broadcast pairs (factor 1/2 folded into the weights)
    uint32_t p_items = 0;
#pragma unroll 1
    for (TileCursor c{(int)blockIdx.x, 0, 0u, 0, 0}; c.item < items; advance(c)) {
      const uint32_t t = c.t;
      const int st = t & 1, slot = t & (kGeoSlots - 1);
      if (c.sb == 0) {
        mbar_wait_a(bar0 + 8 * 16, p_items & 1);  // pfull
        ++p_items;
#pragma unroll
        for (int pr = 0; pr < 2; ++pr)
#pragma unroll
          for (int pl = 0; pl < 4; ++pl) {
            const uint32_t v = sP[pr * kPairRow + pl * n3 + w];
            Ps[2 * pr][pl] = h2dup_lo(v);
            Ps[2 * pr + 1][pl] = h2dup_hi(v);
          }
      }
      mbar_wait_a(bar0 + 8 * (6 + st), (t >> 1) & 1);        // qfull[st]
      mbar_wait_a(bar0 + 8 * (2 + st), ((t >> 1) & 1) ^ 1);  // empty[st]
      const uint32_t* qs = sQ + st * kQStageWords;
      const uint32_t* gw = geo16 + slot * kGeo16;
#pragma unroll
      for (int sq = 0; sq < 2; ++sq) {
        uint4 G[2][5];  // [rp][quantity]
#pragma unroll
        for (int rp = 0; rp < 2; ++rp)
#pragma unroll
          for (int a = 0; a < 5; ++a) G[rp][a] = *reinterpret_cast<const uint4*>(gw + a * (kCols / 2) + 4 * (rp * 2 + sq));
        uint32_t packed[2][4];  // [rp][unit]
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const uint32_t* qr = qs + (2 * sq + u) * kPairRow;
          const uint32_t q0 = qr[w], qx = qr[n3 + w], qy = qr[2 * n3 + w], qz = qr[3 * n3 + w];
#pragma unroll
          for (int rp = 0; rp < 2; ++rp)
#pragma unroll
            for (int rl = 0; rl < 2; ++rl) {
              const int un = rl * 2 + u, r = 2 * rp + rl;
              const uint32_t ax = un == 0 ? G[rp][0].x : un == 1 ? G[rp][0].y : un == 2 ? G[rp][0].z : G[rp][0].w;
              const uint32_t ay = un == 0 ? G[rp][1].x : un == 1 ? G[rp][1].y : un == 2 ? G[rp][1].z : G[rp][1].w;
              const uint32_t az = un == 0 ? G[rp][2].x : un == 1 ? G[rp][2].y : un == 2 ? G[rp][2].z : G[rp][2].w;
              const uint32_t le = un == 0 ? G[rp][3].x : un == 1 ? G[rp][3].y : un == 2 ? G[rp][3].z : G[rp][3].w;
              const uint32_t mm = un == 0 ? G[rp][4].x : un == 1 ? G[rp][4].y : un == 2 ? G[rp][4].z : G[rp][4].w;
              uint32_t hs = h2add(Ps[r][0], q0);
              hs = h2fma(ax, h2add(Ps[r][1], qx), hs);
              hs = h2fma(ay, h2add(Ps[r][2], qy), hs);
              hs = h2fma(az, h2add(Ps[r][3], qz), hs);
              hs = h2fma(le, wd0s, hs);
              hs = h2fma(mm, wm0s, hs);
              packed[rp][un] = h2fma(hs, h2tanh(hs), hs);  // silu(z) / c = z/2 (1 + tanh(z/2))
            }
        }
#pragma unroll
        for (int rp = 0; rp < 2; ++rp) {
          const int chunk = st * 4 + rp * 2 + sq;
          *reinterpret_cast<uint4*>(sB + w * 128 + ((chunk ^ (w & 7)) << 4)) =
              make_uint4(packed[rp][0], packed[rp][1], packed[rp][2], packed[rp][3]);
        }
      }
      proxy_fence();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(bar0 + 8 * st);  // full[st]
    }
  } else if (warp == kMmaWarp) {
    // ============================ MMA issuer + bulk copies =====================================================
    const uint32_t idesc = make_idesc(true);
    const uint32_t tm = 0u;
    const uint64_t bdesc0 = make_b_desc(smem_u32(sB));
    const uint32_t d0 = tm + kDBase;
    const uint32_t pair_bytes = (uint32_t)kPairRow * 4u;
    auto load_prow = [&](int item) {
      if (item < items && elect_one()) {
        const long long g = item / recv_blocks;
        const int i0 = (int)(item - g * recv_blocks) * kRecv;
        const uint32_t bytes = (uint32_t)((min(kRecv, N - i0) + 1) / 2) * pair_bytes;
        const uint32_t* src = pp + ((g * N + i0) / 2) * kPairRow;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(pfull)), "r"(bytes)
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(sP)),
                     "l"(src), "r"(bytes), "r"(smem_u32(pfull))
                     : "memory");
      }
      __syncwarp();
    };
    const uint32_t* q_g = qq;
    const uint32_t last_block_bytes = (uint32_t)((N - (send_blocks - 1) * kSend + 1) / 2) * pair_bytes;
    auto enter_copy = [&](const TileCursor& c) { q_g = qq + ((long long)c.g * N / 2) * kPairRow; };
    auto tile_copy = [&](const TileCursor& c) {
      if (c.item < items && elect_one()) {
        const int st = c.t & 1;
        const uint32_t bytes = c.sb == send_blocks - 1 ? last_block_bytes : (uint32_t)(kSend / 2) * pair_bytes;
        const uint32_t* src = q_g + c.sb * ((kSend / 2) * kPairRow);
        const uint32_t bar = smem_u32(&qfull[st]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(sQ + st * kQStageWords)),
                     "l"(src), "r"(bytes), "r"(bar)
                     : "memory");
      }
      __syncwarp();
    };
    auto advance_copy = [&](TileCursor& c) {
      ++c.t;
      if (++c.sb == send_blocks) {
        c.sb = 0;
        c.item += gridDim.x;
        locate(c);
        enter_copy(c);
      }
    };
    TileCursor cp{(int)blockIdx.x, 0, 0u, 0, 0};
    locate(cp);
    enter_copy(cp);
    load_prow(blockIdx.x);
    for (int i = 0; i < 2; ++i) {
      tile_copy(cp);
      advance_copy(cp);
    }
#pragma unroll 1
    for (TileCursor c{(int)blockIdx.x, 0, 0u, 0, 0}; c.item < items; advance(c)) {
      const uint32_t t = c.t;
      const int st = t & 1;
      mbar_wait_a(bar0 + 8 * st, (t >> 1) & 1);  // full[st]
      tile_copy(cp);
      advance_copy(cp);
      if (c.sb == 0) load_prow(c.item + (int)gridDim.x);
      mbar_wait_a(bar0 + 8 * 5, (t & 1) ^ 1);  // dempty
      tc_fence_after();
      const uint64_t bst = bdesc0 + (uint64_t)(st * (64 >> 4));
      if (elect_one()) {
#pragma unroll
        for (int s = 0; s < 2 * n / 16; ++s) {
          const uint64_t b_sd = bst + (uint64_t)(s * (2048 >> 4));
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
        tc_commit(&empty[st]);
        tc_commit(dfull);
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

template <int NMUL, bool MOMENTS>
static int launch_h2(const float* pos, const float* mass, int B, int N, const void* pp, const void* qq,
                     const float* w_edge1, const float* b2, const void* w2_tc, const float* bn_mul,
                     const float* bn_add, float* agg, float* moments, void* agg16, cudaStream_t stream) {
  constexpr int threads = kWarps * 32;
  const size_t smem = 1024 + (size_t)5 * NMUL * 128 + (size_t)(2 * (kSend / 2) + kRecv / 2) * 4 * 3 * NMUL * 4 +
                      (size_t)kGeoSlots * (4 * kCols * 4 + 5 * kCols * 2) + 24 * sizeof(uint64_t) + 16;
  auto kern = edge_layer_h2_kernel<NMUL, MOMENTS>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) {
    set_error("edge_layer_h2: cudaFuncSetAttribute: %s", cudaGetErrorString(err));
    return SEGNN_E_CUDA;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long items = (long long)B * ((N + kRecv - 1) / kRecv);
  if (items > 0x7fffffffLL) {
    set_error("edge_layer_h2: too many work items");
    return SEGNN_E_UNSUPPORTED;
  }
  const unsigned grid = (unsigned)(items < sms ? items : sms);
  kern<<<grid, threads, smem, stream>>>(pos, mass, B, N, (const uint32_t*)pp, (const uint32_t*)qq, w_edge1, b2,
                                        (const uint32_t*)w2_tc, bn_mul, bn_add, agg, moments, (__half*)agg16);
  err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error("edge_layer_h2: launch: %s", cudaGetErrorString(err));
    return SEGNN_E_CUDA;
  }
  return SEGNN_OK;
}

}  // namespace tc

int edge_layer_h2(const float* pos, const float* mass, int B, int N, int n, const void* pp, const void* qq,
                  const float* w_edge1, const float* b2, const void* w2_tc, const float* bn_mul, const float* bn_add,
                  float* agg, float* moments, void* agg16, cudaStream_t stream) {
  if (N % 2 != 0) {
    set_error("edge_layer_h2: the packed-half mode needs an even graph size (sender pairs), got N=%d", N);
    return SEGNN_E_UNSUPPORTED;
  }
#define SEGNN_H2_CASE(NM)                                                                                            \
  if (n == NM)                                                                                                       \
    return moments ? tc::launch_h2<NM, true>(pos, mass, B, N, pp, qq, w_edge1, b2, w2_tc, bn_mul, bn_add, agg,       \
                                             moments, agg16, stream)                                                     \
                   : tc::launch_h2<NM, false>(pos, mass, B, N, pp, qq, w_edge1, b2, w2_tc, bn_mul, bn_add, agg,      \
                                              nullptr, agg16, stream);
  SEGNN_H2_CASE(32)
  SEGNN_H2_CASE(64)
  SEGNN_H2_CASE(96)
#undef SEGNN_H2_CASE
  set_error("edge_layer_h2: built for hidden multiplicity n in {32, 64, 96}, got n=%d", n);
  return SEGNN_E_UNSUPPORTED;
}

}  // namespace segnn
