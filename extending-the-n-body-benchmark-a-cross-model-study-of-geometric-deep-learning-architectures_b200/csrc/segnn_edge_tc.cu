// K3 (tensor-core modes with bf16 / fp16 operands and fp32 producers; the packed-half variant is segnn_edge_tc_h2.cu):
// fused SEGNN edge layer on tcgen05.
//
// Orientation: OUTPUT CHANNELS are the MMA M dimension (TMEM lanes), EDGES are the MMA N dimension (TMEM columns).
//   D_tile[128 lanes = channel w][32 cols = edges] += A_tile[128 x K] (message_layer_2 weights, resident in TMEM,
//   tcgen05.mma "TS" form) x B[K x 32] (per-edge features: gate(message_layer_1) output, bf16 in shared memory,
//   MN-major, 128B swizzle).
// Six accumulator tiles per 32-edge tile: s (K = 2n: [s' | v'.a1]), g (same K), T1 = W_sv^T s' (K = n),
// D_k = W_vv^T v'_k (K = n, k = x,y,z); message = (silu(s), sigmoid(g) * (a1_k * T1 + D_k)).
//
// Work item = (graph, 4 receivers); tile = 4 receivers x 8 senders = 32 edge columns, column c = (r / 2) * 16 +
// sender_local * 2 + (r % 2).  One persistent CTA of 16 warps per SM:
//   compute warps (warp % 4 < n / 32; warp % 4 = channel block = TMEM lane quadrant, warp / 4 = group g):
//     every thread is one channel.  Group g owns 8 consecutive columns = 2 receivers x 4 senders of every tile
//     (12 receiver-side projection pairs live in registers), in BOTH roles, software-pipelined:
//     produce(t + 1)  ->  epilogue(t)  ->  produce(t + 2)  -> ...
//       produce:  message_layer_1 combine (hoisted projections P_i + Q_j, geometry) + gate -> bf16 B tile (one
//                 conflict-free 16-byte store per plane);
//       epilogue: tcgen05.ld of the six accumulators, gate, in-register accumulation over senders (no atomics, no
//                 shuffles: channels sit on lanes, so the sum over senders is a per-thread sum).
//     Merging the two roles removes the spinning consumer warps of the previous design (12% of the issue slots)
//     and the tensor pipe works on tile t + 1 while the CUDA cores run epilogue(t) / produce(t + 2).
//     All fp32 math is packed (FFMA2 / FADD2 / FMUL2 over receiver pairs; scalar operands ride as broadcast
//     operands), which halves the issue slots of the fp32 work; the gates use one MUFU.TANH each and their
//     constants are folded: pre-activations are carried as z/2 (the 1/2 is folded into the weights), the gate
//     outputs as z/2 * (1 + tanh(z/2)) and (1 + tanh(z/2)), and c_silu, c_sig/2 are folded into the
//     message_layer_2 weight image / the final per-receiver scale.
//   The fourth SM sub-partition (warp % 4 == 3) owns no TMEM lane quadrant at n <= 96; it hosts
//   warp 3:  per tile, in this order: wait full[st]; cp.async.bulk of the sender rows of Q of tile t + 2 (36 KB at
//            n = 96) into the stage the producers just released (issued after the MMAs it used to land only just in
//            time) and, once per item, of the 4 receiver rows of P; wait dempty; the tile's 48 MMAs; then the geometry
//            of tile t + 3 (lane = column: unit vectors, distances, mass products, validity; 8-slot ring) and the
//            position loads of tile t + 4.  This warp's serial loop has to fit the tile period: its index arithmetic
//            runs on per-item base pointers.
// Projection rows P, Q [node][plane][3n] arrive as (scalar part [n] | (gate, vector) pairs [n][2]): the node GEMM
// writes that column order (its weight image is permuted at pack time), so a thread fetches both parts of a plane
// with one 64-bit shared-memory load.
//   warps 7, 11, 15: scalar-channel producers (channel blocks 0, 1, 2): the s' plane of the B tile needs only the
//            (0s) parts of P and Q, so it is produced here for all 32 columns, which moves 22% of the producer
//            instructions off the three compute sub-partitions.
// Synchronisation is mbarrier-only between roles; every wait is bounded and traps instead of hanging.
#include "segnn_edge_tc_common.cuh"

namespace segnn {
namespace tc {

#ifdef SEGNN_K3_TRACE
__device__ long long* g_k3_trace = nullptr;  // [32 warps][64 tiles][8 events]
#endif

template <int NMUL, bool HALF>
__global__ void __launch_bounds__(kWarps * 32, 1)
    edge_layer_tc_kernel(const float* __restrict__ pos, const float* __restrict__ mass, int B, int N,
                         const float* __restrict__ pp, const float* __restrict__ qq,
                         const float* __restrict__ w_edge1,
                         const float* __restrict__ b2, const uint32_t* __restrict__ w2_tc,
                         const float* __restrict__ bn_mul, const float* __restrict__ bn_add,
                         float* __restrict__ agg, int* __restrict__ err_flag) {
  constexpr int n = NMUL;
  constexpr int NW = n / 32;            // channel blocks (compute warps per group)
  static_assert(NW >= 1 && NW <= 3, "warp % 4 == 3 hosts the MMA / geometry warps");
  constexpr int kComputeThreads = 4 * NW * 32;
  constexpr int kScalarWarps = NW;      // warps 7, 11, 15: scalar-channel producers of channel blocks 0, 1, 2
  constexpr int kMmaWarp = 3;           // MMA issuer + bulk copies + tile geometry (lane = column)
  constexpr int n3 = 3 * n;
  constexpr int kQStageFloats = kSend * 4 * n3;
  constexpr int kWeightCols = 3 * n;    // TMEM columns of the weight image
  constexpr int kDBase = kWeightCols;   // accumulator tiles start here (6 x 32 columns)

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
  uint8_t* sB = smem;                                                  // 5n rows x 128 bytes (two 64-byte stages)
  float* sQ = reinterpret_cast<float*>(smem + 5 * n * 128);            // [2][8 senders][4 planes][3n]
  float* sP = sQ + 2 * kQStageFloats;                                  // [4 receivers][4 planes][3n]
  float* geo = sP + kRecv * 4 * n3;                                // [slots][6][32]: ax, ay, az, len, mm, valid
  float* xch = geo + kGeoSlots * 6 * kCols;                            // [2][4 groups][4 comp][n]
  uint64_t* bars = reinterpret_cast<uint64_t*>(xch + 2 * 4 * 4 * n);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  uint64_t* full = bars;         // [2] B stage written (all compute + scalar-producer threads)
  uint64_t* empty = bars + 2;    // [2] B stage consumed (tcgen05.commit; only the scalar producers need it)
  uint64_t* dfull = bars + 4;    // accumulators complete (tcgen05.commit)
  uint64_t* dempty = bars + 5;   // accumulators read out (all compute threads)
  uint64_t* qfull = bars + 6;    // [2] sender rows landed (expect_tx) and tile geometry written
  uint64_t* pfull = bars + 16;   // receiver rows of P landed (expect_tx)

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int cb = warp & 3, grp = warp >> 2;
  const bool is_compute = cb < NW;
  const bool is_scalar = cb == 3 && grp >= 1 && grp - 1 < NW;

  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 4 * NW + kScalarWarps);  // one arrival per producer warp
      mbar_init(&empty[i], 1);
      mbar_init(&qfull[i], 1);
    }
    mbar_init(dfull, 1);
    mbar_init(pfull, 1);
    mbar_init(dempty, 4 * NW);  // one arrival per compute warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // The CTA owns all 512 TMEM columns, so the allocation starts at lane 0 / column 0.  The MMA issuer relies on it
  // (compile-time TMEM addresses keep every tcgen05.mma operand in uniform registers); fail loudly otherwise.
  if (tmem != 0u) {
    if (err_flag) atomicExch(err_flag, 2);
    __trap();
  }

  // ---- message_layer_2 weights -> TMEM (lane = output channel, 2 bf16 of K per column) ------------------------
  if (is_compute && grp == 0) {  // lanes >= n are never read back
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
  // staging buffers start as zeros so that rows never written by a (partial) bulk copy are finite
  for (int i = tid; i < 2 * kQStageFloats + kRecv * 4 * n3; i += kWarps * 32) sQ[i] = 0.f;
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int recv_blocks = (N + kRecv - 1) / kRecv;
  const int send_blocks = (N + kSend - 1) / kSend;
  const int items = B * recv_blocks;  // launch_tc checks that this fits an int
  auto advance = [&](TileCursor& c) {
    ++c.t;
    if (++c.sb == send_blocks) {
      c.sb = 0;
      c.item += gridDim.x;
    }
  };
  auto locate = [&](TileCursor& c) {  // (g, i0) of c.item
    c.g = c.item / recv_blocks;
    c.i0 = (c.item - c.g * recv_blocks) * kRecv;
  };
  auto advance_located = [&](TileCursor& c) {
    ++c.t;
    if (++c.sb == send_blocks) {
      c.sb = 0;
      c.item += gridDim.x;
      locate(c);
    }
  };

  if (is_compute) {
    // ============================ compute warps: produce(t + 1) / epilogue(t) ==================================
    const int w = cb * 32 + lane;  // channel = B-tile row (producer role) = TMEM lane (epilogue role)
    const uint32_t lane_base = (uint32_t)(cb * 32) << 16;
    const uint32_t bar0 = smem_u32(bars);  // barrier addresses: bar0 + 8 * index
    const float2 wd0g = bc2(0.5f * w_edge1[n + w]), wm0g = bc2(0.5f * w_edge1[3 * n + w]), wd1 = bc2(w_edge1[4 * n + w]),
                 wm1 = bc2(w_edge1[5 * n + w]);
    const float2 b2s = bc2(0.5f * b2[w]), b2g = bc2(0.5f * b2[n + w]);
    const float2 half2v = bc2(0.5f);
    float sc_s = kCSilu, sc_v = 0.5f * kCSig, add_s = 0.f;
    if (bn_mul != nullptr) {
      sc_s *= bn_mul[w];
      sc_v *= bn_mul[n + w];
      add_s = bn_add[w];
    }
    // group shape: 2 receivers (pair rp) x 4 senders (quad sq); tile column c = rp * 16 + sender_local * 2 + (r & 1),
    // so the group's 8 columns [8 gi, 8 gi + 8), gi = rp * 2 + sq, are one 16-byte chunk of every B row and one
    // contiguous TMEM column range.  Every packed operation handles the receiver pair of one sender.
    const int rp = grp & 1, sq = grp >> 1, gi = rp * 2 + sq;
    // receiver-side projections of the pair: P[plane * 3 + part]; parts (0s, 0g) carry the factor 1/2
    float2 P[12];
    float2 acc[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[c] = make_float2(0.f, 0.f);

    uint32_t p_items = 0;  // items whose P rows this thread has loaded
    auto load_p = [&]() {
      mbar_wait_a(bar0 + 8 * 16, p_items & 1);  // pfull: this item's rows have landed
      ++p_items;
      const float* r0 = sP + (2 * rp) * 4 * n3;  // rows past the graph end: stale but finite, masked by valid = 0
      const float* r1 = sP + (2 * rp + 1) * 4 * n3;
#pragma unroll
      for (int c = 0; c < 4; ++c) {  // (gate, vector) parts of a channel are adjacent: one 64-bit load per plane
        const float2 a = *reinterpret_cast<const float2*>(r0 + c * n3 + n + 2 * w);
        const float2 b = *reinterpret_cast<const float2*>(r1 + c * n3 + n + 2 * w);
        P[c * 3 + 1] = make_float2(0.5f * a.x, 0.5f * b.x);
        P[c * 3 + 2] = make_float2(a.y, b.y);
      }
    };

    auto produce = [&](const TileCursor& cur) {
      const uint32_t t = cur.t;
      const int st = t & 1, slot = t & (kGeoSlots - 1);
      K3_TRACE(0, t);
      // One wait: qfull[st] = the tile's sender rows have landed AND its geometry is in the ring (the MMA warp writes
      // the geometry before it arms the barrier for the bulk copy).  B stage st was last read by the MMAs of tile
      // t - 2, whose completion (dfull) this thread observed in epilogue(t - 2), which precedes produce(t) in program
      // order: no "stage empty" barrier.  Geometry slot t % 4 was last read in epilogue(t - 4): no "slot empty" one.
      mbar_wait_a(bar0 + 8 * (6 + st), (t >> 1) & 1);               // qfull[st]
      K3_TRACE(1, t);
      const float* qs = sQ + st * kQStageFloats;
      const float* gs = geo + slot * 6 * kCols + 8 * gi;
      uint32_t packed[4][5];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 AX = *reinterpret_cast<const float4*>(gs + 0 * kCols + 4 * h);
        const float4 AY = *reinterpret_cast<const float4*>(gs + 1 * kCols + 4 * h);
        const float4 AZ = *reinterpret_cast<const float4*>(gs + 2 * kCols + 4 * h);
        const float4 LE = *reinterpret_cast<const float4*>(gs + 3 * kCols + 4 * h);
        const float4 MM = *reinterpret_cast<const float4*>(gs + 4 * kCols + 4 * h);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int s4 = 2 * h + e;
          const int sl = 4 * sq + s4;
          // rows past the graph end hold stale (finite: the buffers are zero-initialised) data, masked by valid = 0
          const float* qr = qs + sl * 4 * n3;
          float q[12];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float2 gv = *reinterpret_cast<const float2*>(qr + c * n3 + n + 2 * w);
            q[c * 3 + 1] = gv.x;
            q[c * 3 + 2] = gv.y;
          }
          const float2 ax = e ? hi2(AX) : lo2(AX), ay = e ? hi2(AY) : lo2(AY), az = e ? hi2(AZ) : lo2(AZ);
          const float2 le = e ? hi2(LE) : lo2(LE), mm = e ? hi2(MM) : lo2(MM);
          // half pre-activations of the scalar / gate channels: (P0 + Q0)/2 + a.(P0k + Q0k)/2 + |r| wd/2 + m_i m_j wm/2
          // (the scalar channel s' is produced by the scalar-producer warps on the fourth SM sub-partition)
          float2 hg = __ffma2_rn(bc2(q[1]), half2v, P[1]);
          hg = __ffma2_rn(ax, __ffma2_rn(bc2(q[4]), half2v, P[4]), hg);
          hg = __ffma2_rn(ay, __ffma2_rn(bc2(q[7]), half2v, P[7]), hg);
          hg = __ffma2_rn(az, __ffma2_rn(bc2(q[10]), half2v, P[10]), hg);
          hg = __ffma2_rn(le, wd0g, hg);
          hg = __ffma2_rn(mm, wm0g, hg);
          float2 tt = __fadd2_rn(P[2], bc2(q[2]));
          tt = __ffma2_rn(le, wd1, tt);
          tt = __ffma2_rn(mm, wm1, tt);
          const float2 zx = __ffma2_rn(ax, tt, __fadd2_rn(P[5], bc2(q[5])));
          const float2 zy = __ffma2_rn(ay, tt, __fadd2_rn(P[8], bc2(q[8])));
          const float2 zz = __ffma2_rn(az, tt, __fadd2_rn(P[11], bc2(q[11])));
          const float2 tg = tanh2(hg);
          const float2 vx = __ffma2_rn(tg, zx, zx);   // 2 sigmoid(z_g) z_v
          const float2 vy = __ffma2_rn(tg, zy, zy);
          const float2 vz = __ffma2_rn(tg, zz, zz);
          float2 dt = __fmul2_rn(ax, vx);
          dt = __ffma2_rn(ay, vy, dt);
          dt = __ffma2_rn(az, vz, dt);
          packed[s4][1] = pack_pair<HALF>(dt.x, dt.y);
          packed[s4][2] = pack_pair<HALF>(vx.x, vx.y);
          packed[s4][3] = pack_pair<HALF>(vy.x, vy.y);
          packed[s4][4] = pack_pair<HALF>(vz.x, vz.y);
        }
      }
      // 4 senders x 2 receivers = 8 consecutive columns = one 16-byte chunk per plane row (conflict-free with the
      // 128B swizzle: 8 consecutive rows hit 8 distinct chunks)
      const int chunk = st * 4 + gi;
#pragma unroll
      for (int p = 1; p < 5; ++p) {
        const int row = p * n + w;
        *reinterpret_cast<uint4*>(sB + row * 128 + ((chunk ^ (row & 7)) << 4)) =
            make_uint4(packed[0][p], packed[1][p], packed[2][p], packed[3][p]);
      }
      // one arrival per warp (480 per-thread arrivals on one mbarrier serialise: the phase completed several hundred
      // clk after the last store): every lane fences its own stores, the warp syncs, lane 0 arrives
      proxy_fence();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(bar0 + 8 * st);  // full[st]
      K3_TRACE(2, t);
    };

    uint32_t items_done = 0;
    auto epilogue = [&](const TileCursor& cur) {
      const uint32_t t = cur.t;
      const int slot = t & (kGeoSlots - 1);
      const float* gs = geo + slot * 6 * kCols + 8 * gi;
      mbar_wait_a(bar0 + 8 * 4, t & 1);  // dfull
      K3_TRACE(3, t);
      tc_fence_after();
      uint32_t d[6][8];
#pragma unroll
      for (int tile = 0; tile < 6; ++tile) SEGNN_TMEM_LD8(tmem + lane_base + kDBase + tile * kCols + 8 * gi, d[tile]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      // accumulators are in registers: hand the TMEM tiles back to the MMA warps
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(bar0 + 8 * 5);  // dempty
      K3_TRACE(4, t);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 AX = *reinterpret_cast<const float4*>(gs + 0 * kCols + 4 * h);
        const float4 AY = *reinterpret_cast<const float4*>(gs + 1 * kCols + 4 * h);
        const float4 AZ = *reinterpret_cast<const float4*>(gs + 2 * kCols + 4 * h);
        const float4 VA = *reinterpret_cast<const float4*>(gs + 5 * kCols + 4 * h);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = 4 * h + 2 * e;
          const float2 ax = e ? hi2(AX) : lo2(AX), ay = e ? hi2(AY) : lo2(AY), az = e ? hi2(AZ) : lo2(AZ);
          const float2 va = e ? hi2(VA) : lo2(VA);
          const float2 ys = __fadd2_rn(u2f2(d[0][j], d[0][j + 1]), b2s);
          const float2 yg = __fadd2_rn(u2f2(d[1][j], d[1][j + 1]), b2g);
          const float2 ts = tanh2(ys), tg = tanh2(yg);
          const float2 ms = __ffma2_rn(ys, ts, ys);
          const float2 g1 = __ffma2_rn(va, tg, va);  // valid * (1 + tanh): masks self edges and padding
          const float2 t1 = u2f2(d[2][j], d[2][j + 1]);
          acc[0] = __ffma2_rn(va, ms, acc[0]);
          acc[1] = __ffma2_rn(g1, __ffma2_rn(ax, t1, u2f2(d[3][j], d[3][j + 1])), acc[1]);
          acc[2] = __ffma2_rn(g1, __ffma2_rn(ay, t1, u2f2(d[4][j], d[4][j + 1])), acc[2]);
          acc[3] = __ffma2_rn(g1, __ffma2_rn(az, t1, u2f2(d[5][j], d[5][j + 1])), acc[3]);
        }
      }
      K3_TRACE(5, t);
      if (cur.sb == send_blocks - 1) {
        // item complete: the two groups of a receiver pair (sender quads 0 / 1) each hold half of the senders for both
        // receivers; quad sq keeps receiver 2 rp + sq and hands the other one to its partner.  xch is double
        // buffered by item parity, one named barrier per item.
        const long long g = cur.item / recv_blocks;
        const int i0 = (int)(cur.item - g * recv_blocks) * kRecv;
        float* xb = xch + (items_done & 1) * (4 * 4 * n);
        float own[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          own[c] = sq ? acc[c].y : acc[c].x;
          xb[((rp * 2 + (sq ^ 1)) * 4 + c) * n + w] = sq ? acc[c].x : acc[c].y;  // partner's slot
        }
        named_barrier(1, kComputeThreads);
#pragma unroll
        for (int c = 0; c < 4; ++c) own[c] += xb[((rp * 2 + sq) * 4 + c) * n + w];
        const int r = 2 * rp + sq;
        if (i0 + r < N) {
          float* o = agg + (g * N + i0 + r) * 4 * n;
          o[w] = fmaf(own[0], sc_s, add_s);
          o[n + w] = own[1] * sc_v;
          o[2 * n + w] = own[2] * sc_v;
          o[3 * n + w] = own[3] * sc_v;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[c] = make_float2(0.f, 0.f);
        ++items_done;
      }
    };

    // iteration i: produce(tile i), then epilogue(tile i - 1)
    TileCursor pc{(int)blockIdx.x, 0, 0u, 0, 0}, ec{(int)blockIdx.x, 0, 0u, 0, 0};
    bool primed = false;
#pragma unroll 1
    while (ec.item < items) {
      if (pc.item < items) {
        if (pc.sb == 0) load_p();
        produce(pc);
        advance(pc);
      }
      if (primed) {
        epilogue(ec);
        advance(ec);
      }
      primed = true;
    }
  } else if (is_scalar) {
    // ============================ scalar-channel producers (4th SM sub-partition) ===============================
    // s' = silu-gate of the l=0 "scalar" pre-activation needs only the (0s) parts of P and Q; it is independent of
    // the gate / vector pipeline, so three warps on the sub-partition that owns no TMEM lane quadrant produce plane 0
    // of the B tile for all 32 columns: 22% of the producer instructions leave the three compute sub-partitions.
    const int w = (grp - 1) * 32 + lane;
    const uint32_t bar0 = smem_u32(bars);
    const float2 wd0s = bc2(0.5f * w_edge1[w]), wm0s = bc2(0.5f * w_edge1[2 * n + w]);
    const float2 half2v = bc2(0.5f);
    float2 Ps[4][2];  // [plane][receiver pair], factor 1/2 folded
    uint32_t p_items = 0;
    for (TileCursor c{(int)blockIdx.x, 0, 0u, 0, 0}; c.item < items; advance(c)) {
      const uint32_t t = c.t;
      const int st = t & 1, slot = t & (kGeoSlots - 1);
      if (c.sb == 0) {
        mbar_wait_a(bar0 + 8 * 16, p_items & 1);  // pfull
        ++p_items;
#pragma unroll
        for (int pl = 0; pl < 4; ++pl)
#pragma unroll
          for (int rp = 0; rp < 2; ++rp)
            Ps[pl][rp] = make_float2(0.5f * sP[(2 * rp) * 4 * n3 + pl * n3 + w], 0.5f * sP[(2 * rp + 1) * 4 * n3 + pl * n3 + w]);
      }
      K3_TRACE(0, t);
      mbar_wait_a(bar0 + 8 * (6 + st), (t >> 1) & 1);            // qfull[st]: sender rows + geometry
      mbar_wait_a(bar0 + 8 * (2 + st), ((t >> 1) & 1) ^ 1);      // empty[st]: MMAs of tile t - 2 have read the stage
      K3_TRACE(1, t);
      const float* qs = sQ + st * kQStageFloats;
      const float* gs = geo + slot * 6 * kCols;
#pragma unroll
      for (int sq = 0; sq < 2; ++sq) {
        uint32_t packed[2][4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float4 G[2][5];
#pragma unroll
          for (int rp = 0; rp < 2; ++rp)
#pragma unroll
            for (int a = 0; a < 5; ++a)
              G[rp][a] = *reinterpret_cast<const float4*>(gs + a * kCols + 8 * (rp * 2 + sq) + 4 * h);
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int s4 = 2 * h + e;
            const float* qr = qs + (4 * sq + s4) * 4 * n3;
            const float q0 = qr[w], qx = qr[n3 + w], qy = qr[2 * n3 + w], qz = qr[3 * n3 + w];
#pragma unroll
            for (int rp = 0; rp < 2; ++rp) {
              const float2 ax = e ? hi2(G[rp][0]) : lo2(G[rp][0]), ay = e ? hi2(G[rp][1]) : lo2(G[rp][1]);
              const float2 az = e ? hi2(G[rp][2]) : lo2(G[rp][2]), le = e ? hi2(G[rp][3]) : lo2(G[rp][3]);
              const float2 mm = e ? hi2(G[rp][4]) : lo2(G[rp][4]);
              float2 hs = __ffma2_rn(bc2(q0), half2v, Ps[0][rp]);
              hs = __ffma2_rn(ax, __ffma2_rn(bc2(qx), half2v, Ps[1][rp]), hs);
              hs = __ffma2_rn(ay, __ffma2_rn(bc2(qy), half2v, Ps[2][rp]), hs);
              hs = __ffma2_rn(az, __ffma2_rn(bc2(qz), half2v, Ps[3][rp]), hs);
              hs = __ffma2_rn(le, wd0s, hs);
              hs = __ffma2_rn(mm, wm0s, hs);
              const float2 so = __ffma2_rn(hs, tanh2(hs), hs);  // silu(z) / c = z/2 (1 + tanh(z/2))
              packed[rp][s4] = pack_pair<HALF>(so.x, so.y);
            }
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
      K3_TRACE(2, t);
    }
  } else if (warp == kMmaWarp) {
    // ============================ MMA issuer ==================================================================
    // Warp-uniform loop; only the tcgen05 instructions sit under elect.sync so descriptors stay in uniform registers.
    const uint32_t idesc = make_idesc(HALF);
    const uint32_t tm = 0u;  // checked above
    const uint64_t bdesc0 = make_b_desc(smem_u32(sB));
    const uint32_t d0 = tm + kDBase;
    // bulk copy of the (up to) 4 receiver rows of P of an item (one elected lane)
    auto load_prow = [&](int item) {
      if (item < items && elect_one()) {
        const long long g = item / recv_blocks;
        const int i0 = (int)(item - g * recv_blocks) * kRecv;
        const uint32_t bytes = (uint32_t)min(kRecv, N - i0) * 4 * n3 * (uint32_t)sizeof(float);
        const float* src = pp + (g * N + i0) * 4 * n3;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(pfull)), "r"(bytes)
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(sP)),
                     "l"(src), "r"(bytes), "r"(smem_u32(pfull))
                     : "memory");
      }
      __syncwarp();
    };
    // Inputs of a tile = its geometry (lane = column = rp * 16 + sender * 2 + (r & 1)) + the 8 sender rows of Q.
    // This warp's loop is a serial instruction chain that has to fit the tile period, so its index arithmetic is kept
    // to per-item base pointers (set when a cursor enters an item) plus 32-bit offsets.
    const int gsl = (lane >> 1) & 7, gr = 2 * (lane >> 4) + (lane & 1);
    float gsx = 0.f, gsy = 0.f, gsz = 0.f, gsm = 0.f, grx = 0.f, gry = 0.f, grz = 0.f, grm = 0.f;  // raw loads
    bool gvalid = false;
    const float* pos_g = pos;    // positions / masses of the geometry cursor's graph
    const float* mass_g = mass;
    const float* q_g = qq;       // sender rows of the copy cursor's graph
    const uint32_t row_bytes = 4u * n3 * (uint32_t)sizeof(float);
    const uint32_t last_block_bytes = (uint32_t)(N - (send_blocks - 1) * kSend) * row_bytes;
    bool rvalid = false;
    auto enter_geo = [&](const TileCursor& c) {  // per item: graph base pointers + this lane's receiver
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
    auto enter_copy = [&](const TileCursor& c) { q_g = qq + (long long)c.g * N * 4 * n3; };
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
    // geometry of tile c into ring slot t % 8 (last read in epilogue(t - 8), long complete)
    auto tile_geometry = [&](const TileCursor& c) {
      if (c.item < items) {
        const int slot = c.t & (kGeoSlots - 1);
        const float dx = gsx - grx, dy = gsy - gry, dz = gsz - grz;
        const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
        const float inv = rsqrtf(fmaxf(d2, 1e-24f));  // r / max(|r|, 1e-12) with the approximate reciprocal root
        const float sc = kY1 * inv;
        float* gs = geo + slot * 6 * kCols;
        gs[0 * kCols + lane] = sc * dx;
        gs[1 * kCols + lane] = sc * dy;
        gs[2 * kCols + lane] = sc * dz;
        gs[3 * kCols + lane] = d2 * inv;
        gs[4 * kCols + lane] = gsm * grm;
        gs[5 * kCols + lane] = gvalid ? 1.0f : 0.0f;
      }
      __syncwarp();
    };
    // sender rows of tile c into Q stage t & 1: arms qfull[st], which also publishes the tile's geometry (written by this
    // warp at least one loop iteration earlier)
    auto tile_copy = [&](const TileCursor& c) {
      if (c.item < items && elect_one()) {
        const int st = c.t & 1;
        const uint32_t bytes = c.sb == send_blocks - 1 ? last_block_bytes : (uint32_t)kSend * row_bytes;
        const float* src = q_g + c.sb * (kSend * 4 * n3);
        const uint32_t bar = smem_u32(&qfull[st]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(sQ + st * kQStageFloats)),
                     "l"(src), "r"(bytes), "r"(bar)
                     : "memory");
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
    auto advance_copy = [&](TileCursor& c) {
      ++c.t;
      if (++c.sb == send_blocks) {
        c.sb = 0;
        c.item += gridDim.x;
        locate(c);
        enter_copy(c);
      }
    };
    // cp: next tile whose rows are copied (t + 2 in the loop); gp: next tile whose geometry is written (t + 3)
    TileCursor cp{(int)blockIdx.x, 0, 0u, 0, 0}, gp{(int)blockIdx.x, 0, 0u, 0, 0};
    locate(cp);
    locate(gp);
    enter_copy(cp);
    enter_geo(gp);
    load_prow(blockIdx.x);
    for (int i = 0; i < 3; ++i) {  // geometry of tiles 0, 1, 2; rows of tiles 0, 1
      tile_load(gp);
      tile_geometry(gp);
      advance_geo(gp);
      if (i < 2) {
        tile_copy(cp);
        advance_copy(cp);
      }
    }
    tile_load(gp);  // positions of tile 3 in flight
    for (TileCursor c{(int)blockIdx.x, 0, 0u, 0, 0}; c.item < items; advance(c)) {
      const uint32_t t = c.t;
      const int st = t & 1;
      K3_TRACE(0, t);
      mbar_wait_a(smem_u32(&full[st]), (t >> 1) & 1);
      K3_TRACE(1, t);
      // first thing after full[st]: every producer is done with Q stage st, so the rows of tile t + 2 start their trip
      // from L2 now (issued after the MMAs they landed only just in time: the copy latency was the tile period's
      // critical path); every producer thread holds this item's P rows in registers: fetch the next item's
      tile_copy(cp);
      advance_copy(cp);
      if (c.sb == 0) load_prow(c.item + (int)gridDim.x);
      K3_TRACE(5, t);
      mbar_wait_a(smem_u32(dempty), (t & 1) ^ 1);
      K3_TRACE(2, t);
      tc_fence_after();
      const uint64_t bst = bdesc0 + (uint64_t)(st * (64 >> 4));  // column half of the swizzled rows
      if (elect_one()) {
#pragma unroll
        for (int s = 0; s < 2 * n / 16; ++s) {
          // rows [16 s, 16 s + 16) of the (s', dot) planes; 16 rows = 2048 bytes
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
      K3_TRACE(3, t);
      // off the critical path: geometry of tile t + 3 into the ring, positions of tile t + 4 in flight
      tile_geometry(gp);
      K3_TRACE(6, t);
      advance_geo(gp);
      tile_load(gp);
      K3_TRACE(4, t);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

// ---- weight image: [128 lanes][3n columns] of bf16 pairs ------------------------------------------------------
__global__ void pack_w2_kernel(const float* __restrict__ ss, const float* __restrict__ vs,
                               const float* __restrict__ sv, const float* __restrict__ vv, int n, int half,
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
          // output carried as z/2; inputs arrive as silu/c_silu and 2 sigmoid/c_sig * v (see the kernel header)
          v[h] = k < n ? 0.5f * kCSilu * ss[(size_t)k * 2 * n + oc] : 0.25f * kCSig * vs[(size_t)(k - n) * 2 * n + oc];
        } else if (c < 2 * n + n / 2) {  // T1 tile: K = s'
          const int k = 2 * (c - 2 * n) + h;
          v[h] = kCSilu * sv[(size_t)k * n + w];
        } else {  // D_k tiles: K = v'_k
          const int k = 2 * (c - 2 * n - n / 2) + h;
          v[h] = 0.5f * kCSig * vv[(size_t)k * n + w];
        }
      }
    }
    if (half) {
      out[idx] = (uint32_t)__half_as_ushort(__float2half_rn(v[0])) | ((uint32_t)__half_as_ushort(__float2half_rn(v[1])) << 16);
    } else {
      const __nv_bfloat16 lo = __float2bfloat16(v[0]), hi = __float2bfloat16(v[1]);
      out[idx] = (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
    }
  }
}

template <int NMUL, bool HALF>
static int launch_tc(const float* pos, const float* mass, int B, int N, const float* pp, const float* qq,
                     const float* w_edge1,
                     const float* b2, const void* w2_tc, const float* bn_mul, const float* bn_add, float* agg,
                     int* err_flag, cudaStream_t stream) {
  constexpr int threads = kWarps * 32;
  const size_t smem = 1024 + (size_t)5 * NMUL * 128 + (size_t)2 * kSend * 4 * 3 * NMUL * sizeof(float) +
                      (size_t)kRecv * 4 * 3 * NMUL * sizeof(float) + (size_t)kGeoSlots * 6 * kCols * sizeof(float) +
                      (size_t)2 * 4 * 4 * NMUL * sizeof(float) +
                      24 * sizeof(uint64_t) + 16;
  auto kern = edge_layer_tc_kernel<NMUL, HALF>;
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
  if (items > 0x7fffffffLL) {
    set_error("edge_layer_tc: too many work items");
    return SEGNN_E_UNSUPPORTED;
  }
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
                  const float* b2, const void* w2_tc, const float* bn_mul, const float* bn_add, float* agg, int half,
                  cudaStream_t stream) {
  if (n == 32)
    return half ? tc::launch_tc<32, true>(pos, mass, B, N, pp, qq, w_edge1, b2, w2_tc, bn_mul, bn_add, agg, nullptr, stream)
                : tc::launch_tc<32, false>(pos, mass, B, N, pp, qq, w_edge1, b2, w2_tc, bn_mul, bn_add, agg, nullptr, stream);
  if (n == 64)
    return half ? tc::launch_tc<64, true>(pos, mass, B, N, pp, qq, w_edge1, b2, w2_tc, bn_mul, bn_add, agg, nullptr, stream)
                : tc::launch_tc<64, false>(pos, mass, B, N, pp, qq, w_edge1, b2, w2_tc, bn_mul, bn_add, agg, nullptr, stream);
  if (n == 96)
    return half ? tc::launch_tc<96, true>(pos, mass, B, N, pp, qq, w_edge1, b2, w2_tc, bn_mul, bn_add, agg, nullptr, stream)
                : tc::launch_tc<96, false>(pos, mass, B, N, pp, qq, w_edge1, b2, w2_tc, bn_mul, bn_add, agg, nullptr, stream);
  set_error("edge_layer_tc: tensor-core mode is built for hidden multiplicity n in {32, 64, 96} (hidden_features "
            "64/128/192), got n=%d", n);
  return SEGNN_E_UNSUPPORTED;
}

#ifdef SEGNN_K3_TRACE
}  // namespace segnn
extern "C" int segnn_debug_set_k3_trace(long long* buf) {
  return (int)cudaMemcpyToSymbol(segnn::tc::g_k3_trace, &buf, sizeof(buf));
}
namespace segnn {
#endif

int64_t pack_w2_tc(const float* ss, const float* vs, const float* sv, const float* vv, int n, int half, void* out,
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
  tc::pack_w2_kernel<<<(128 * 3 * n + 255) / 256, 256, 0, stream>>>(ss, vs, sv, vv, n, half, (uint32_t*)out);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error("segnn_pack_w2_tc: launch: %s", cudaGetErrorString(err));
    return SEGNN_E_CUDA;
  }
  return bytes;
}

}  // namespace segnn
