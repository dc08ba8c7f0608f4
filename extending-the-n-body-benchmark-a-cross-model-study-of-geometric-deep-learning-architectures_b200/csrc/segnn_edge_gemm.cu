// Large-graph form of the SEGNN edge layer (models/segnn/segnn.py:264-284 + the scatter-add of :205), forward and
// backward, fp32-accurate on the tensor cores.  It exists for BASELINE configuration 4 (one N = 1000 fully connected
// graph, ~1 M edges, forward + backward): the fused fp32 kernels (segnn_edge_fp32.cu, segnn_edge_bwd.cu) keep one thread
// group per stationary node, which leaves a 1000-node graph at ~1 % of the tensor roofline.
//
// Formulation.  message_layer_1 stays hoisted to node level (P, Q: DESIGN section 3), so per edge only elementwise work
// and the message_layer_2 contraction remain.  The edges of a chunk of graphs are enumerated as rows
// row = (graph * N + receiver) * N + sender (the diagonal is kept and masked), and the contraction runs as plain GEMMs
// over those rows with the 3xTF32 split (hi * hi + hi * lo + lo * hi in fp32 TMEM accumulators):
//   E1   msg1_rows_kernel       P_i + Q_j, geometry, gate           -> XS [rows][2n] = (s', v'.a), XV [rows][3][n] = v'
//   G1   segnn_gemm_tf32x3      Y  [rows][3n] = XS * Wcat [2n][3n]   (ys | yg | t1),  DV [3 rows][n] = XV * W_vv
//   E2f  gate2_fwd_kernel       gate, sum over senders, moments     -> agg (forward ends here)
//   E2b  gate2_bwd_kernel       dm = A dagg_i + B m + C, gate bwd   -> dY, dDV in place, bias-gradient rows
//   G2   gemm_tn_tf32x3         dWcat = XS^T dY, dW_vv = XV^T dDV    (K = rows, split over the CTAs, partial sums in
//                               TMEM, fixed-order reduction: bit-identical from run to run, no atomics)
//   G3   segnn_gemm_tf32x3      dXS = dY * Wcat^T, dXV = dDV * W_vv^T
//   E3   msg1_bwd_tile_kernel   gate / combine backward on 32 x 32 (receiver, sender) tiles: per-tile partial sums over
//                               senders (dP, d w_edge1) and receivers (dQ), then a fixed-order reduction over the tiles
// Per-edge tensors live in a caller-provided workspace for one chunk of graphs at a time (16 n floats per row), never
// for the whole batch; every kernel streams them once at HBM speed, the GEMMs run on tcgen05.
#include "segnn_common.cuh"

namespace segnn {
namespace eg {

// Gates of this file: sigmoid(x) = rcp.rn(1 + ex2.approx(-x log2 e)), ~1e-7 relative for the pre-activations that occur
// (|x| of a few units; the error grows like |x| * 6e-8), a fifth of the instructions of 1 / (1 + expf(-x)): the
// elementwise kernels below are bound by instruction latency, not by HBM.
__device__ __forceinline__ float sigmoid_f(float x) { return __frcp_rn(1.0f + __expf(-x)); }
__device__ __forceinline__ float silu_gate_f(float x) { return kCSilu * x * sigmoid_f(x); }
__device__ __forceinline__ float sig_gate_f(float x) { return kCSig * sigmoid_f(x); }
// value and derivative from one sigmoid
__device__ __forceinline__ void silu_gate_vg(float x, float& val, float& grad) {
  const float s = sigmoid_f(x);
  val = kCSilu * x * s;
  grad = kCSilu * s * (1.0f + x * (1.0f - s));
}
__device__ __forceinline__ void sig_gate_vg(float x, float& val, float& grad) {
  const float s = sigmoid_f(x);
  val = kCSig * s;
  grad = kCSig * s * (1.0f - s);
}
__device__ __forceinline__ float silu_gate_grad(float x) {
  float v, g;
  silu_gate_vg(x, v, g);
  return g;
}

constexpr int kRecv = 4;    // receivers per block of E1
constexpr int kTileJ = 32;  // senders per geometry tile of E1
constexpr int kSY = 4;      // row lanes (threadIdx.y) of the per-node kernels
constexpr int kTileO = 256; // "other" nodes per geometry tile of the per-node kernels

struct RowArgs {
  const float *pos, *mass, *pp, *qq, *w_edge1;
  int N, n;
  int64_t node0;  // first node of the chunk
};

// ---------------------------------------------------------------------------------------------------------------
// E1: rows of the message_layer_2 input.  block = (NT channels, 4 sender lanes), 4 receivers of one graph.
// ---------------------------------------------------------------------------------------------------------------
// NCH = hidden multiplicity n as a compile-time constant (32, 64, 96: every offset v * n becomes an immediate; the
// address arithmetic of the runtime-n build was two thirds of the instructions of these kernels), 0 = runtime n.
template <int NCH>
__global__ void __launch_bounds__(96 * 4) msg1_rows_kernel(const RowArgs a, float* __restrict__ xs,
                                                         float* __restrict__ xv) {
  __shared__ float gs[kRecv][kTileJ][8];
  const int w = threadIdx.x, y = threadIdx.y;
  const int n = NCH ? NCH : a.n, N = a.N, n3 = 3 * n;
  const int bpg = (N + kRecv - 1) / kRecv;
  const int64_t gl = blockIdx.x / bpg;
  const int i0 = (int)(blockIdx.x - gl * bpg) * kRecv;
  const int64_t base = a.node0 + gl * N;
  const bool act = w < n;
  const int tid = y * blockDim.x + w;

  float st[kRecv][12];
  float wd0s = 0.f, wd0g = 0.f, wm0s = 0.f, wm0g = 0.f, wd1 = 0.f, wm1 = 0.f;
#pragma unroll
  for (int r = 0; r < kRecv; ++r)
#pragma unroll
    for (int k = 0; k < 12; ++k) st[r][k] = 0.f;
  if (act) {
#pragma unroll
    for (int r = 0; r < kRecv; ++r) {
      if (i0 + r < N) {
        const float* sr = a.pp + (base + i0 + r) * 4 * n3;
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int k = 0; k < 3; ++k) st[r][c * 3 + k] = sr[c * n3 + k * n + w];
      }
    }
    wd0s = a.w_edge1[w];
    wd0g = a.w_edge1[n + w];
    wm0s = a.w_edge1[2 * n + w];
    wm0g = a.w_edge1[3 * n + w];
    wd1 = a.w_edge1[4 * n + w];
    wm1 = a.w_edge1[5 * n + w];
  }
  for (int j0 = 0; j0 < N; j0 += kTileJ) {
    __syncthreads();
    if (tid < kRecv * kTileJ) {
      const int r = tid / kTileJ, jj = tid % kTileJ;
      const int i = min(i0 + r, N - 1), j = min(j0 + jj, N - 1);
      const int64_t ni = base + i, nj = base + j;
      float ux, uy, uz, len;
      unit_vec(a.pos[nj * 3 + 0] - a.pos[ni * 3 + 0], a.pos[nj * 3 + 1] - a.pos[ni * 3 + 1],
               a.pos[nj * 3 + 2] - a.pos[ni * 3 + 2], ux, uy, uz, len);
      gs[r][jj][0] = kY1 * ux;
      gs[r][jj][1] = kY1 * uy;
      gs[r][jj][2] = kY1 * uz;
      gs[r][jj][3] = len;
      gs[r][jj][4] = a.mass[nj] * a.mass[ni];
    }
    __syncthreads();
    if (!act) continue;
    for (int jj = y; jj < kTileJ && j0 + jj < N; jj += blockDim.y) {
      const int j = j0 + jj;
      const float* qr = a.qq + (base + j) * 4 * n3;
      float qv[12];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int k = 0; k < 3; ++k) qv[c * 3 + k] = qr[c * n3 + k * n + w];
#pragma unroll
      for (int r = 0; r < kRecv; ++r) {
        if (i0 + r >= N) break;
        const float ax = gs[r][jj][0], ay = gs[r][jj][1], az = gs[r][jj][2], len = gs[r][jj][3], mm = gs[r][jj][4];
        float S[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) S[k] = st[r][k] + qv[k];
        const float zs = S[0] + ax * S[3] + ay * S[6] + az * S[9] + len * wd0s + mm * wm0s;
        const float zg = S[1] + ax * S[4] + ay * S[7] + az * S[10] + len * wd0g + mm * wm0g;
        const float t = S[2] + len * wd1 + mm * wm1;
        const float gg = sig_gate_f(zg);
        const float vx = gg * (ax * t + S[5]), vy = gg * (ay * t + S[8]), vz = gg * (az * t + S[11]);
        const int64_t row = (gl * N + i0 + r) * N + j;
        float* xr = xs + row * 2 * n;
        xr[w] = silu_gate_f(zs);
        xr[n + w] = ax * vx + ay * vy + az * vz;
        float* vr = xv + row * n3;
        vr[w] = vx;
        vr[n + w] = vy;
        vr[2 * n + w] = vz;
      }
    }
  }
}

// geometry of (stationary node, tile of other nodes) for the per-node kernels: (ax, ay, az, valid, len, mm)
__device__ __forceinline__ void node_tile_geometry(const float* __restrict__ pos, const float* __restrict__ mass,
                                                   int64_t base, int N, int ir, int o0, float sgn, float (*gs)[8],
                                                   int tid, int nthreads) {
  const float prx = pos[(base + ir) * 3 + 0], pry = pos[(base + ir) * 3 + 1], prz = pos[(base + ir) * 3 + 2];
  const float mr = mass[base + ir];
  for (int t = tid; t < kTileO; t += nthreads) {
    const int oo = o0 + t;
    const int64_t on = base + (oo < N ? oo : N - 1);
    float ux, uy, uz, len;
    unit_vec(sgn * (pos[on * 3 + 0] - prx), sgn * (pos[on * 3 + 1] - pry), sgn * (pos[on * 3 + 2] - prz), ux, uy, uz,
             len);
    gs[t][0] = kY1 * ux;
    gs[t][1] = kY1 * uy;
    gs[t][2] = kY1 * uz;
    gs[t][3] = (oo < N && oo != ir) ? 1.0f : 0.0f;
    gs[t][4] = len;
    gs[t][5] = mass[on] * mr;
  }
}

// fixed-order sum over the kSY row lanes of V per-thread values: red[y][v][x]; lane y == 0 returns the totals in vals
template <int V>
__device__ __forceinline__ void lane_reduce(float* red, float (&vals)[V], int w, int y, int NT) {
#pragma unroll
  for (int v = 0; v < V; ++v) red[(y * V + v) * NT + w] = vals[v];
  __syncthreads();
  if (y == 0) {
#pragma unroll
    for (int v = 0; v < V; ++v) {
      float s = red[v * NT + w];
      for (int yy = 1; yy < kSY; ++yy) s += red[(yy * V + v) * NT + w];
      vals[v] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// E2f: gate of message_layer_2 + sum over senders (+ moments, folded eval BatchNorm).  block = one receiver.
// ---------------------------------------------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(96 * kSY)
    gate2_fwd_kernel(const RowArgs a, const float* __restrict__ yy, const float* __restrict__ dv,
                     const float* __restrict__ b2, const float* __restrict__ bn_mul, const float* __restrict__ bn_add,
                     float* __restrict__ agg, float* __restrict__ moments) {
  extern __shared__ __align__(16) float smem[];
  float(*gs)[8] = reinterpret_cast<float(*)[8]>(smem);
  float* red = smem + kTileO * 8;
  const int w = threadIdx.x, y = threadIdx.y, NT = blockDim.x;
  const int n = NCH ? NCH : a.n, N = a.N, n3 = 3 * n;
  const int64_t rl = blockIdx.x;  // receiver, local to the chunk
  const int64_t gl = rl / N;
  const int ir = (int)(rl - gl * N);
  const int64_t base = a.node0 + gl * N;
  const bool act = w < n;
  const float b2s = act ? b2[w] : 0.f, b2g = act ? b2[n + w] : 0.f;
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // sum m_s, m_x, m_y, m_z, sum m_s^2, sum |m_v|^2
  for (int o0 = 0; o0 < N; o0 += kTileO) {
    __syncthreads();
    node_tile_geometry(a.pos, a.mass, base, N, ir, o0, 1.0f, gs, y * NT + w, NT * kSY);
    __syncthreads();
    if (!act) continue;
    const int lim = min(kTileO, N - o0);
#pragma unroll 4
    for (int t = y; t < lim; t += kSY) {
      const float valid = gs[t][3];  // the diagonal row is finite and masked (no branch: the loads of four rows overlap)
      const int64_t row = rl * N + o0 + t;
      const float* yr = yy + row * n3;
      const float* dr = dv + row * n3;
      const float ys = yr[w] + b2s, yg = yr[n + w] + b2g, t1 = yr[2 * n + w];
      const float dx = dr[w], dy = dr[n + w], dz = dr[2 * n + w];
      const float ms = valid * silu_gate_f(ys), gt = valid * sig_gate_f(yg);
      const float mx = gt * fmaf(gs[t][0], t1, dx), my = gt * fmaf(gs[t][1], t1, dy), mz = gt * fmaf(gs[t][2], t1, dz);
      acc[0] += ms;
      acc[1] += mx;
      acc[2] += my;
      acc[3] += mz;
      acc[4] = fmaf(ms, ms, acc[4]);
      acc[5] += mx * mx + my * my + mz * mz;
    }
  }
  __syncthreads();
  lane_reduce<6>(red, acc, w, y, NT);
  if (y == 0 && act) {
    const int64_t r = a.node0 + rl;
    if (moments != nullptr) {
      moments[r * 2 * n + w] = acc[4];
      moments[r * 2 * n + n + w] = acc[5];
    }
    if (bn_mul != nullptr) {
      const float ms = bn_mul[w], mv = bn_mul[n + w];
      acc[0] = fmaf(acc[0], ms, bn_add[w]);
      acc[1] *= mv;
      acc[2] *= mv;
      acc[3] *= mv;
    }
    float* o = agg + r * 4 * n;
    o[w] = acc[0];
    o[n + w] = acc[1];
    o[2 * n + w] = acc[2];
    o[3 * n + w] = acc[3];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// E2b: backward of the gate of message_layer_2: (Y, DV) -> (dY, dDV), in place or from the rows the forward kept;
// bias-gradient row per receiver.
// ---------------------------------------------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(96 * kSY)
    gate2_bwd_kernel(const RowArgs a, const float* y_in, const float* dv_in, float* yy, float* dv,
                     const float* __restrict__ b2, const float* __restrict__ bnA, const float* __restrict__ bnB,
                     const float* __restrict__ bnC, const float* __restrict__ dagg, float* __restrict__ db2_rows) {
  extern __shared__ __align__(16) float smem[];
  float(*gs)[8] = reinterpret_cast<float(*)[8]>(smem);
  float* red = smem + kTileO * 8;
  const int w = threadIdx.x, y = threadIdx.y, NT = blockDim.x;
  const int n = NCH ? NCH : a.n, N = a.N, n3 = 3 * n;
  const int64_t rl = blockIdx.x;
  const int64_t gl = rl / N;
  const int ir = (int)(rl - gl * N);
  const int64_t base = a.node0 + gl * N;
  const bool act = w < n;
  float b2s = 0.f, b2g = 0.f, As = 0.f, Av = 0.f, Bs = 0.f, Bv = 0.f, Cs = 0.f, G[4] = {0.f, 0.f, 0.f, 0.f};
  if (act) {
    b2s = b2[w];
    b2g = b2[n + w];
    As = bnA[w];
    Av = bnA[n + w];
    Bs = bnB[w];
    Bv = bnB[n + w];
    Cs = bnC[w];
    const float* gr = dagg + (a.node0 + rl) * 4 * n;
#pragma unroll
    for (int c = 0; c < 4; ++c) G[c] = gr[c * n + w];
  }
  float acc[2] = {0.f, 0.f};
  for (int o0 = 0; o0 < N; o0 += kTileO) {
    __syncthreads();
    node_tile_geometry(a.pos, a.mass, base, N, ir, o0, 1.0f, gs, y * NT + w, NT * kSY);
    __syncthreads();
    if (!act) continue;
    const int lim = min(kTileO, N - o0);
    // four rows per step, all loads before any store: the rows may be updated in place (y_in == yy), so the compiler
    // cannot move a load above a store by itself and one row per step left a single row of loads in flight
    constexpr int kU = 4;
    for (int t0 = y; t0 < lim; t0 += kSY * kU) {
      float yv[kU][3], dvv[kU][3];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int t = min(t0 + u * kSY, lim - 1);
        const int64_t row = rl * N + o0 + t;
        const float* yi = y_in + row * n3;
        const float* di = dv_in + row * n3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          yv[u][c] = yi[c * n + w];
          dvv[u][c] = di[c * n + w];
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int t = t0 + u * kSY;
        if (t < lim) {
          const int64_t row = rl * N + o0 + t;
          float* yr = yy + row * n3;
          float* dr = dv + row * n3;
          const float ax = gs[t][0], ay = gs[t][1], az = gs[t][2], valid = gs[t][3];
          const float ys = yv[u][0] + b2s, yg = yv[u][1] + b2g, t1 = yv[u][2];
          float ms, dsilu, gt, dsig;
          silu_gate_vg(ys, ms, dsilu);
          sig_gate_vg(yg, gt, dsig);
          const float ux = fmaf(ax, t1, dvv[u][0]), uy = fmaf(ay, t1, dvv[u][1]), uz = fmaf(az, t1, dvv[u][2]);
          const float dms = valid * (As * G[0] + Bs * ms + Cs);
          const float dmx = valid * (Av * G[1] + Bv * gt * ux);
          const float dmy = valid * (Av * G[2] + Bv * gt * uy);
          const float dmz = valid * (Av * G[3] + Bv * gt * uz);
          const float dys = dms * dsilu;
          const float dyg = dsig * (dmx * ux + dmy * uy + dmz * uz);
          const float dux = gt * dmx, duy = gt * dmy, duz = gt * dmz;
          yr[w] = dys;
          yr[n + w] = dyg;
          yr[2 * n + w] = ax * dux + ay * duy + az * duz;
          dr[w] = dux;
          dr[n + w] = duy;
          dr[2 * n + w] = duz;
          acc[0] += dys;
          acc[1] += dyg;
        }
      }
    }
  }
  __syncthreads();
  lane_reduce<2>(red, acc, w, y, NT);
  if (y == 0 && act) {
    db2_rows[rl * 2 * n + w] = acc[0];
    db2_rows[rl * 2 * n + n + w] = acc[1];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// E3 (one pass): the pre-activation of message_layer_1 depends on P_i + Q_j only, so every edge contributes the SAME
// twelve numbers c_ij to dP_i and to dQ_j.  block = (32 receivers) x (16 senders) of one graph.  A warp covers 8 channels
// x 4 sender lanes (a request reads 32-byte pieces of four rows: whole sectors), so the sum over the sender lanes is
// two xor-shuffles and the main loop has no block-level barrier:
//   dP side: c_ij summed over the thread's senders in registers, over the sender lanes by shuffles, one partial row per
//            (receiver, sender tile);
//   dQ side: a shared-memory tile [16 senders][12][channels]; every (sender, channel) cell is owned by ONE thread, which
//            adds its receivers' terms in program order (plain read-modify-write, no atomics).
// tile_reduce_kernel adds the partial rows in tile order: bit-identical results from run to run.  dxs / dxv are read
// once and the gates recomputed once (the two-pass form did both twice).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTI = 32, kTJ = 16, kSL = 4;

// STAGE_Q: the sender-side projections of the tile are staged in shared memory (12 of the 17 global loads per edge go
// away, the shared-memory footprint doubles).  It wins when the grid is small (training-size graphs: latency), it
// loses on large graphs (one block less per SM: configuration 4 went 31.0 -> 32.4 ms), so the host picks.
template <int NCH, bool STAGE_Q>
__global__ void __launch_bounds__(384)
    msg1_bwd_tile_kernel(const RowArgs a, const float* __restrict__ dxs, const float* __restrict__ dxv, int tiles_i,
                         int tiles_j, int64_t chunk_nodes, float* __restrict__ dp_part, float* __restrict__ dq_part,
                         float* __restrict__ dwe_part) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sl = lane >> 3, w = warp * 8 + (lane & 7);
  const int NC = NCH ? ((NCH + 7) & ~7) : (blockDim.x >> 5) * 8;   // channels covered by the block (>= n)
  float(*gs)[6] = reinterpret_cast<float(*)[6]>(smem);          // [kTI * kTJ][6]
  float* dq = smem + kTI * kTJ * 6;                              // [kTJ][12 * NC + 8]: the four sender lanes of a warp
  const int dqs = 12 * NC + 8;                                   // own consecutive senders: + 8 floats keeps them on distinct banks
  float* qs = dq + kTJ * dqs;                                    // [kTJ][12 * NC + 8]: the Q rows of the tile's senders
  const int n = NCH ? NCH : a.n, N = a.N, n3 = 3 * n;
  const int jt = blockIdx.x % tiles_j, it_ = (blockIdx.x / tiles_j) % tiles_i;
  const int64_t gl = blockIdx.x / (tiles_i * tiles_j);
  const int i0 = it_ * kTI, j0 = jt * kTJ;
  const int64_t base = a.node0 + gl * N;        // first node of the graph
  const int64_t lbase = gl * N;                 // the same, local to the chunk
  const bool act = w < n;
  for (int t = threadIdx.x; t < kTI * kTJ; t += blockDim.x) {
    const int ii = t / kTJ, jj = t % kTJ;
    const int i = min(i0 + ii, N - 1), j = min(j0 + jj, N - 1);
    const int64_t ni = base + i, nj = base + j;
    float ux, uy, uz, len;
    unit_vec(a.pos[nj * 3 + 0] - a.pos[ni * 3 + 0], a.pos[nj * 3 + 1] - a.pos[ni * 3 + 1],
             a.pos[nj * 3 + 2] - a.pos[ni * 3 + 2], ux, uy, uz, len);
    gs[t][0] = kY1 * ux;
    gs[t][1] = kY1 * uy;
    gs[t][2] = kY1 * uz;
    gs[t][3] = (i0 + ii < N && j0 + jj < N && i != j) ? 1.0f : 0.0f;
    gs[t][4] = len;
    gs[t][5] = a.mass[nj] * a.mass[ni];
  }
#pragma unroll
  for (int q = 0; q < kTJ / kSL; ++q) {
    const int jj = sl + kSL * q;
    // the sender-side projections of the tile are read by all 32 receivers: staged once (they were 12 of the 17 global
    // loads per edge, each a four-line request -- the L1 tag stage was the busiest unit of the kernel)
    const float* qrow = a.qq + (base + min(j0 + jj, N - 1)) * 4 * n3;
#pragma unroll
    for (int v = 0; v < 12; ++v) {
      dq[jj * dqs + v * NC + w] = 0.f;
      if (STAGE_Q) qs[jj * dqs + v * NC + w] = act ? qrow[v * n + w] : 0.f;
    }
  }
  float wd0s = 0.f, wd0g = 0.f, wm0s = 0.f, wm0g = 0.f, wd1 = 0.f, wm1 = 0.f;
  if (act) {
    wd0s = a.w_edge1[w];
    wd0g = a.w_edge1[n + w];
    wm0s = a.w_edge1[2 * n + w];
    wm0g = a.w_edge1[3 * n + w];
    wd1 = a.w_edge1[4 * n + w];
    wm1 = a.w_edge1[5 * n + w];
  }
  __syncthreads();
  const int ilim = min(kTI, N - i0);
  const int wc = act ? w : 0;  // inactive lanes compute on channel 0 (keeps the warp converged for the shuffles)
  for (int ii = 0; ii < ilim; ++ii) {
    const int i = i0 + ii;
    float acc[18];
#pragma unroll
    for (int v = 0; v < 18; ++v) acc[v] = 0.f;
    float st[12];
    const float* sr = a.pp + (base + i) * 4 * n3;
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int k = 0; k < 3; ++k) st[c * 3 + k] = sr[c * n3 + k * n + wc];
#pragma unroll
    for (int q = 0; q < kTJ / kSL; ++q) {
      const int jj = sl + kSL * q;
      const float* g = gs[ii * kTJ + jj];
      if (g[3] != 0.f) {
        const int j = j0 + jj;
        const int64_t row = (lbase + i) * N + j;
        const float* xr = dxs + row * 2 * n;
        const float* vr = dxv + row * n3;
        const float d0 = xr[wc], d1 = xr[n + wc], d2 = vr[wc], d3 = vr[n + wc], d4 = vr[2 * n + wc];
        float S[12];
        if (STAGE_Q) {
          const float* orow = qs + jj * dqs + w;  // this thread staged exactly these cells (same jj, same w): no barrier
#pragma unroll
          for (int v = 0; v < 12; ++v) S[v] = st[v] + orow[v * NC];
        } else {
          const float* orow = a.qq + (base + j) * 4 * n3 + wc;
#pragma unroll
          for (int v = 0; v < 12; ++v) S[v] = st[v] + orow[v * n];
        }
        const float ax = g[0], ay = g[1], az = g[2], len = g[4], mm = g[5];
        const float zs = S[0] + ax * S[3] + ay * S[6] + az * S[9] + len * wd0s + mm * wm0s;
        const float zg = S[1] + ax * S[4] + ay * S[7] + az * S[10] + len * wd0g + mm * wm0g;
        const float tt = S[2] + len * wd1 + mm * wm1;
        const float zx = ax * tt + S[5], zy = ay * tt + S[8], zz = az * tt + S[11];
        float gg, dsig;
        sig_gate_vg(zg, gg, dsig);
        const float tvx = d2 + ax * d1, tvy = d3 + ay * d1, tvz = d4 + az * d1;
        const float dzs = silu_gate_grad(zs) * d0;
        const float dzg = dsig * (zx * tvx + zy * tvy + zz * tvz);
        const float dzx = gg * tvx, dzy = gg * tvy, dzz = gg * tvz;
        const float dt = ax * dzx + ay * dzy + az * dzz;
        const float c12[12] = {dzs, dzg, dt, ax * dzs, ax * dzg, dzx, ay * dzs, ay * dzg, dzy, az * dzs, az * dzg, dzz};
        if (act) {
          float* dqj = dq + jj * dqs + w;
#pragma unroll
          for (int v = 0; v < 12; ++v) dqj[v * NC] += c12[v];
        }
#pragma unroll
        for (int v = 0; v < 12; ++v) acc[v] += c12[v];
        acc[12] += len * dzs;
        acc[13] += len * dzg;
        acc[14] += mm * dzs;
        acc[15] += mm * dzg;
        acc[16] += len * dt;
        acc[17] += mm * dt;
      }
    }
    // sum over the four sender lanes (lanes l, l ^ 8, l ^ 16, l ^ 24 hold the same channel): fixed association
#pragma unroll
    for (int v = 0; v < 18; ++v) {
      acc[v] += __shfl_xor_sync(0xffffffffu, acc[v], 8);
      acc[v] += __shfl_xor_sync(0xffffffffu, acc[v], 16);
    }
    if (sl == 0 && act) {
      const int64_t pr = (int64_t)jt * chunk_nodes + lbase + i;  // [sender tile][node of the chunk]
      float* o = dp_part + pr * 12 * n;
      float* ow = dwe_part + pr * 6 * n;
#pragma unroll
      for (int v = 0; v < 12; ++v) o[v * n + w] = acc[v];
#pragma unroll
      for (int v = 0; v < 6; ++v) ow[v * n + w] = acc[12 + v];
    }
  }
  if (act) {
#pragma unroll
    for (int q = 0; q < kTJ / kSL; ++q) {
      const int jj = sl + kSL * q;
      if (j0 + jj < N) {
        const int64_t pr = (int64_t)it_ * chunk_nodes + lbase + j0 + jj;
        float* o = dq_part + pr * 12 * n;
#pragma unroll
        for (int v = 0; v < 12; ++v) o[v * n + w] = dq[jj * dqs + v * NC + w];
      }
    }
  }
}

// out[node0 + r][e] = sum over the tiles t (in order) of part[t][r][e]
__global__ void tile_reduce_kernel(const float* __restrict__ part, int tiles, int64_t rows, int cols,
                                   float* __restrict__ out) {
  const int64_t total = rows * cols;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    int t = 0;
    for (; t + 4 <= tiles; t += 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) s[k] += part[(int64_t)(t + k) * total + idx];
    }
    for (int k = 0; t < tiles; ++t, ++k) s[k] += part[(int64_t)t * total + idx];
    out[idx] = (s[0] + s[1]) + (s[2] + s[3]);
  }
}

// Wcat [2n][3n] = [[W_ss | W_sv], [W_vs | 0]],  WcatT [3n][2n] = its transpose (from the transposed blocks)
__global__ void build_wcat_kernel(int n, const float* __restrict__ ss, const float* __restrict__ vs,
                                  const float* __restrict__ sv, const float* __restrict__ tss,
                                  const float* __restrict__ tvs, const float* __restrict__ tsv,
                                  float* __restrict__ wcat, float* __restrict__ wcat_t) {
  const int total = 6 * n * n;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    {
      const int u = idx / (3 * n), c = idx % (3 * n);
      float v;
      if (u < n) v = c < 2 * n ? ss[u * 2 * n + c] : sv[u * n + (c - 2 * n)];
      else v = c < 2 * n ? vs[(u - n) * 2 * n + c] : 0.f;
      wcat[idx] = v;
    }
    if (wcat_t != nullptr) {
      const int c = idx / (2 * n), u = idx % (2 * n);  // WcatT[c][u] = Wcat[u][c]
      float v;
      if (c < 2 * n) v = u < n ? tss[c * n + u] : tvs[c * n + (u - n)];
      else v = u < n ? tsv[(c - 2 * n) * n + u] : 0.f;
      wcat_t[idx] = v;
    }
  }
}

// dWcat [2n][3n], dWvv [n][n], db2 [2n] -> the caller's gradient blocks
__global__ void scatter_w2_grads_kernel(int n, const float* __restrict__ dwcat, const float* __restrict__ dwvv,
                                        const float* __restrict__ db2_chunks, int chunks, float* __restrict__ dss,
                                        float* __restrict__ dvs, float* __restrict__ dsv, float* __restrict__ dvv,
                                        float* __restrict__ db2) {
  const int total = 6 * n * n;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int u = idx / (3 * n), c = idx % (3 * n);
    const float v = dwcat[idx];
    if (u < n) {
      if (c < 2 * n) dss[u * 2 * n + c] = v;
      else dsv[u * n + (c - 2 * n)] = v;
    } else if (c < 2 * n) {
      dvs[(u - n) * 2 * n + c] = v;
    }
    if (idx < n * n) dvv[idx] = dwvv[idx];
    if (idx < 2 * n) {
      float s = db2_chunks[idx];
      for (int k = 1; k < chunks; ++k) s += db2_chunks[(int64_t)k * 2 * n + idx];
      db2[idx] = s;
    }
  }
}

// ===============================================================================================================
// G2: C[M][N] (+)= sum_k A[k][m] B[k][n]  (both operands row-major over K = rows: "TN"), 3xTF32 on tcgen05.
// Both operands are MN-major for the tensor core (consecutive m / n are contiguous in HBM), so a row of A or B goes
// into shared memory as it is read: 128-byte pieces of 32 floats form the swizzle atoms, 4 consecutive k per atom.
//   smem stage: A hi|lo [kc / 4 k-groups][4 m-atoms][4][128 B], B hi|lo [kc / 4][NP / 32 n-atoms][4][128 B]
// grid = (splits, m-blocks of 128); every CTA accumulates its K range in TMEM and writes one partial tile; the partial
// tiles are added in a fixed order by tn_reduce_kernel.
// ===============================================================================================================
namespace tn {

constexpr int kStages = 2;
// 7 loader warps + the MMA warp = 256 threads: registers are allocated for a warp count rounded up to a multiple of four,
// so a ninth warp would cap every thread at 168 registers and spill the loaders' two prefetch sets
constexpr int kLoadWarps = 7;
constexpr int kMmaWarp = kLoadWarps;
constexpr int kThreads = (kLoadWarps + 1) * 32;

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
// MN-major tf32 operands have ONE legal shared-memory layout, "128-byte swizzle with 32-byte atomicity"
// (SWIZZLE_128B_BASE32B, layout type 1): atoms of 32 floats (128 B along M / N) x 4 k, rows 128 B apart, the 32-byte
// piece c of row r stored at piece c ^ r.  LBO = distance of consecutive atoms along M / N (512 B), SBO = distance of
// consecutive 4-k groups.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(512 >> 4) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
// kind::tf32, D = f32, A and B MN-major ("transposed"), M = 128
__device__ __forceinline__ uint32_t make_idesc(int N) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 2u << 7;
  d |= 2u << 10;
  d |= 1u << 15;
  d |= 1u << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(128 >> 4) << 24;
  return d;
}
// v = hi + lo with hi a tf32 number (round to nearest, ties away from zero, on the integer pipe: cvt.rna.tf32 runs on
// the 16-lane XU pipe and was the busiest unit of the loaders) and lo = v - hi exact in fp32; the tensor core ignores
// the low 13 mantissa bits of lo (2^-22 of v)
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
  lo = v - hi;
}

#define SEGNN_TN_LD16(taddr, r)                                                                                    \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"((r)[0]), "=r"((r)[1]), "=r"((r)[2]), "=r"((r)[3]), "=r"((r)[4]), "=r"((r)[5]), "=r"((r)[6]),   \
                 "=r"((r)[7]), "=r"((r)[8]), "=r"((r)[9]), "=r"((r)[10]), "=r"((r)[11]), "=r"((r)[12]),              \
                 "=r"((r)[13]), "=r"((r)[14]), "=r"((r)[15])                                                         \
               : "r"(taddr))

struct Args {
  const float* A;
  int64_t lda;
  const float* B;
  int64_t ldb;
  int64_t K;
  int M, N, NP;      // NP = N rounded up to 32
  int kc;            // rows of K per stage: 16, 32 or 64 (as many as the loaders' registers and shared memory hold)
  int groups;        // G > 1: C = sum_g A[:, g M : (g + 1) M]^T B[:, g N : (g + 1) N] (rows that hold G blocks side by side)
  int a_atoms;       // 32-float atoms of A per 4-k group in shared memory (4; G M / 32 + padding for G > 1)
  int tmem_cols;     // power of two >= NP (>= 32)
  float* part;       // [splits][mblocks * 128][NP]
};

constexpr int kMaxPieces = 15;  // 16-byte pieces per loader thread and stage

__global__ void __launch_bounds__(kThreads, 1) gemm_tn_tf32x3_kernel(const Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int NP = a.NP, KC = a.kc;
  const int G = a.groups;
  const int a_group = a.a_atoms * 512;         // one 4-k group of A
  const int a_bytes = (KC / 4) * a_group;      // one of (hi, lo): [KC / 4 k-groups][a_atoms m-atoms][4][128 B]
  const int b_bytes = KC * G * NP * 4;         //                 [KC / 4][G NP / 32 n-atoms][4][128 B]
  const int stage_bytes = 2 * a_bytes + 2 * b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * stage_bytes);
  uint64_t* full = bars;              // [kStages] loaders -> MMA
  uint64_t* empty = bars + kStages;   // [kStages] MMA -> loaders
  uint64_t* dfull = bars + 2 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dfull + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x, splits = gridDim.x, mb = blockIdx.y;
  const int m0 = mb * 128;
  // 16-byte pieces per row of A that carry data (G > 1: the G blocks of M columns side by side, M a multiple of 32)
  const int ma4 = G > 1 ? G * a.M / 4 : ((min(a.M - m0, 128) + 31) & ~31) >> 2;
  const int64_t chunks_total = (a.K + KC - 1) / KC;
  const int64_t c_begin = chunks_total * split / splits, c_end = chunks_total * (split + 1) / splits;

  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(a.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], kLoadWarps);
      mbar_init(&empty[i], 1);
    }
    mbar_init(dfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (ma4 < 32 || G > 1) {  // m-atoms past M are never written by the loaders: they must read as zeros
    for (int i = tid; i < kStages * stage_bytes / 16; i += kThreads)
      reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    proxy_fence();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < kLoadWarps) {
    // ===================== loaders: fp32 rows -> (hi, lo) tf32, MN-major swizzled atoms =============================
    // The loads of chunk c + 1 are issued before chunk c is converted (two register sets): two stages of bytes in
    // flight per SM.
    const int pa = KC * ma4;            // 16-byte pieces of the A chunk
    const int npr = G * NP / 4;         // pieces per row of B
    const int total = pa + KC * npr;
    const int b_atoms_bytes = (G * NP / 32) * 512;  // one 4-k group of B
    // piece -> (row k of the chunk, shared-memory offset, offset from the chunk's first row in HBM): the same for every
    // chunk, so the divisions are done once.  meta = smem offset | k << 20 | is_a << 30, -1 for no piece; goff < 0 for
    // the zero padding past M / N.
    int meta[kMaxPieces], goff[kMaxPieces];
#pragma unroll
    for (int b = 0; b < kMaxPieces; ++b) {
      const int p = tid + b * (kLoadWarps * 32);
      meta[b] = -1;
      goff[b] = -1;
      if (p < total) {
        const bool is_a = p < pa;
        const int q = is_a ? p : p - pa;
        const int per = is_a ? ma4 : npr;
        const int k = q / per, c4 = q - k * per;
        const int rr = k & 3, cc = c4 & 7;
        const int off = (is_a ? 0 : 2 * a_bytes) + (k >> 2) * (is_a ? a_group : b_atoms_bytes) + (c4 >> 3) * 512 +
                        rr * 128 + ((((cc >> 1) ^ rr) << 5) | ((cc & 1) << 4));
        meta[b] = off | (k << 20) | ((is_a ? 1 : 0) << 30);
        const int col = (is_a ? m0 : 0) + 4 * c4;
        if (col < (is_a ? G * a.M : G * a.N)) goff[b] = (int)(k * (is_a ? a.lda : a.ldb)) + col;
      }
    }
    auto issue = [&](int64_t c, float4 (&r)[kMaxPieces]) {
      const int64_t k0 = c * KC;
      const float* abase = a.A + k0 * a.lda;
      const float* bbase = a.B + k0 * a.ldb;
      const int rows_left = (int)min((int64_t)KC, a.K - k0);
#pragma unroll
      for (int b = 0; b < kMaxPieces; ++b) {
        r[b] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (goff[b] >= 0 && ((meta[b] >> 20) & 63) < rows_left)
          r[b] = *reinterpret_cast<const float4*>(((meta[b] >> 30) ? abase : bbase) + goff[b]);
      }
    };
    auto consume = [&](uint32_t it, const float4 (&r)[kMaxPieces]) {
      const int s = it % kStages;
      mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);
      uint8_t* st = smem + s * stage_bytes;
#pragma unroll
      for (int b = 0; b < kMaxPieces; ++b) {
        if (meta[b] >= 0) {
          const int off = meta[b] & 0xFFFFF;
          const int lo_off = (meta[b] >> 30) ? a_bytes : b_bytes;
          float4 hi, lo;
          split_tf32(r[b].x, hi.x, lo.x);
          split_tf32(r[b].y, hi.y, lo.y);
          split_tf32(r[b].z, hi.z, lo.z);
          split_tf32(r[b].w, hi.w, lo.w);
          *reinterpret_cast<float4*>(st + off) = hi;
          *reinterpret_cast<float4*>(st + off + lo_off) = lo;
        }
      }
      proxy_fence();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
    };
    float4 r0[kMaxPieces], r1[kMaxPieces];
    if (c_begin < c_end) issue(c_begin, r0);
    uint32_t it = 0;
    for (int64_t c = c_begin; c < c_end; c += 2, it += 2) {
      if (c + 1 < c_end) issue(c + 1, r1);
      consume(it, r0);
      if (c + 1 < c_end) {
        if (c + 2 < c_end) issue(c + 2, r0);
        consume(it + 1, r1);
      }
    }
  } else {
    // ===================== MMA issuer ==============================================================================
    uint32_t it = 0;
    const int nsub = (NP + 255) / 256;
    for (int64_t c = c_begin; c < c_end; ++c, ++it) {
      const int s = it % kStages;
      mbar_wait(&full[s], (it / kStages) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sa = smem_u32(smem + s * stage_bytes);
        const uint32_t sa_lo = sa + a_bytes, sb = sa + 2 * a_bytes, sb_lo = sb + b_bytes;
        const uint32_t b_group = (uint32_t)(G * NP / 32) * 512u;  // one 4-k group of B; an MMA (K = 8) reads two
        for (int kg = 0; kg < KC / 8; ++kg) {
          if (G > 1) {
            // group g: rows [g M, g M + 128) of A (those past M belong to the next group or to the zero padding: they
            // only feed accumulator rows nobody reads) against columns [g N, (g + 1) N) of B, all into one accumulator
            const uint32_t idesc = make_idesc(NP);
            for (int g = 0; g < G; ++g) {
              const uint32_t ao = kg * 2 * a_group + g * (a.M / 32) * 512, bo = kg * 2 * b_group + g * (NP / 32) * 512;
              const uint64_t ah = make_desc_mn(sa + ao, a_group), al = make_desc_mn(sa_lo + ao, a_group);
              const uint64_t bh = make_desc_mn(sb + bo, b_group), bl = make_desc_mn(sb_lo + bo, b_group);
              mma_tf32(tmem, ah, bl, idesc, (it > 0 || kg > 0 || g > 0) ? 1u : 0u);
              mma_tf32(tmem, al, bh, idesc, 1u);
              mma_tf32(tmem, ah, bh, idesc, 1u);
            }
            continue;
          }
          for (int sub = 0; sub < nsub; ++sub) {
            const int width = min(256, NP - sub * 256);
            const uint32_t idesc = make_idesc(width);
            const uint32_t d = tmem + sub * 256;
            const uint32_t ao = kg * 2 * a_group, bo = kg * 2 * b_group + sub * 8 * 512;
            const uint64_t ah = make_desc_mn(sa + ao, a_group), al = make_desc_mn(sa_lo + ao, a_group);
            const uint64_t bh = make_desc_mn(sb + bo, b_group), bl = make_desc_mn(sb_lo + bo, b_group);
            mma_tf32(d, ah, bl, idesc, (it > 0 || kg > 0) ? 1u : 0u);
            mma_tf32(d, al, bh, idesc, 1u);
            mma_tf32(d, ah, bh, idesc, 1u);
          }
        }
        tc_commit(&empty[s]);
        if (c == c_end - 1) tc_commit(dfull);
      }
      __syncwarp();
    }
  }
  // ===================== epilogue (warps 0..3): TMEM -> partial tile ===============================================
  if (warp < 4 && c_end > c_begin) {
    mbar_wait(dfull, 0);
    tc_fence_after();
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const int mblocks = gridDim.y;
    float* dst = a.part + (((int64_t)split * mblocks + mb) * 128 + warp * 32 + lane) * NP;
    for (int col = 0; col < NP; col += 16) {
      uint32_t u[16];
      SEGNN_TN_LD16(tmem + lane_base + col, u);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(dst + col + j) = make_float4(__uint_as_float(u[j]), __uint_as_float(u[j + 1]),
                                                                __uint_as_float(u[j + 2]), __uint_as_float(u[j + 3]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(a.tmem_cols));
}

// C[m][n] = (accumulate ? C : 0) + sum over the splits in order
__global__ void tn_reduce_kernel(const float* __restrict__ part, int splits, int mblocks, int M, int N, int NP,
                                 float* __restrict__ C, int64_t ldc, int accumulate) {
  const int64_t total = (int64_t)M * N;
  const int64_t tile = (int64_t)mblocks * 128 * NP;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(idx / N), c = (int)(idx - (int64_t)m * N);
    const float* p = part + (int64_t)m * NP + c;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    int k = 0;
    for (; k + 4 <= splits; k += 4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) s[j] += p[(int64_t)(k + j) * tile];
    }
    for (int j = 0; k < splits; ++k, ++j) s[j] += p[(int64_t)k * tile];
    const float v = (s[0] + s[1]) + (s[2] + s[3]);
    float* o = C + (int64_t)m * ldc + c;
    *o = accumulate ? *o + v : v;
  }
}

// rows per stage: as many of 64 / 32 / 16 as the loaders' registers (kMaxPieces pieces per thread) and two stages of
// shared memory hold; 0 = does not fit
static inline int a_atoms_for(int M, int groups) { return groups > 1 ? groups * (M / 32) + 4 - M / 32 : 4; }
static inline size_t smem_for(int kc, int M, int NP, int groups) {
  return 1024 + (size_t)kStages * 2 * ((size_t)(kc / 4) * a_atoms_for(M, groups) * 512 + (size_t)kc * groups * NP * 4) + 64;
}
static inline int kc_for(int M, int NP, int groups) {
  const int ma = groups > 1 ? groups * M : ((M < 128 ? M : 128) + 31) & ~31;
  for (int kc = 64; kc >= 16; kc >>= 1) {
    if (kc * (ma / 4 + groups * NP / 4) <= kMaxPieces * kLoadWarps * 32 &&
        smem_for(kc, M, NP, groups) <= (kc == 64 ? 200 : 227) * 1024)
      return kc;
  }
  return 0;
}
static inline int splits_for(int64_t K, int mblocks, int sms) {
  const int64_t chunks = (K + 63) / 64;  // rows per stage <= 64: never more splits than chunks
  int64_t s = sms / mblocks;
  if (s < 1) s = 1;
  if (s > chunks) s = chunks;
  return (int)s;
}

}  // namespace tn
}  // namespace eg

static int device_sms() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

}  // namespace segnn

using namespace segnn;

extern "C" {

int64_t segnn_gemm_tn_tf32x3_workspace(int64_t K, int M, int N) {
  if (K < 1 || M < 1 || N < 1 || N > 512) return -1;
  const int mblocks = (M + 127) / 128;
  const int NP = (N + 31) & ~31;
  const int splits = eg::tn::splits_for(K, mblocks, 148 * 2);  // upper bound independent of the device
  return (int64_t)splits * mblocks * 128 * NP * (int64_t)sizeof(float);
}

int segnn_gemm_tn_grouped_tf32x3(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t K, int M, int N,
                                 int groups, float* C, int64_t ldc, int accumulate, float* workspace,
                                 segnn_stream_t stream) {
  SEGNN_CHECK_ARG(K >= 1 && M >= 1 && N >= 1 && N <= 512 && groups >= 1 && lda >= (int64_t)groups * M &&
                      ldb >= (int64_t)groups * N && ldc >= N,
                  "bad sizes");
  SEGNN_CHECK_ARG(A && B && C && workspace, "null pointer");
  SEGNN_CHECK_ARG((M & 3) == 0 && (N & 3) == 0 && (lda & 3) == 0 && (ldb & 3) == 0 &&
                      (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
                  "M, N, lda, ldb must be multiples of 4 floats and the pointers 16-byte aligned");
  SEGNN_CHECK_ARG(groups == 1 || ((M & 31) == 0 && (N & 31) == 0 && M <= 128 && N <= 256),
                  "groups > 1 needs M, N multiples of 32, M <= 128, N <= 256");
  const int mblocks = (M + 127) / 128;
  const int NP = (N + 31) & ~31;
  int tmem_cols = 32;
  while (tmem_cols < NP) tmem_cols <<= 1;
  const int splits = eg::tn::splits_for(K, mblocks, device_sms());
  const int kc = eg::tn::kc_for(M, NP, groups);
  if (kc == 0) {
    set_error("segnn_gemm_tn_tf32x3: M=%d N=%d groups=%d does not fit shared memory / the loaders", M, N, groups);
    return SEGNN_E_UNSUPPORTED;
  }
  eg::tn::Args a{A, lda, B, ldb, K, M, N, NP, kc, groups, eg::tn::a_atoms_for(M, groups), tmem_cols, workspace};
  const size_t smem = eg::tn::smem_for(kc, M, NP, groups);
  cudaError_t err = cudaFuncSetAttribute(eg::tn::gemm_tn_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
  if (err != cudaSuccess) {
    set_error("segnn_gemm_tn_tf32x3: cudaFuncSetAttribute(%zu bytes): %s", smem, cudaGetErrorString(err));
    return SEGNN_E_CUDA;
  }
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid((unsigned)splits, (unsigned)mblocks);
  eg::tn::gemm_tn_tf32x3_kernel<<<grid, eg::tn::kThreads, smem, s>>>(a);
  SEGNN_CHECK_LAUNCH();
  const int64_t total = (int64_t)M * N;
  eg::tn::tn_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(workspace, splits, mblocks, M, N, NP, C, ldc,
                                                                         accumulate);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_gemm_tn_tf32x3(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t K, int M, int N, float* C,
                         int64_t ldc, int accumulate, float* workspace, segnn_stream_t stream) {
  return segnn_gemm_tn_grouped_tf32x3(A, lda, B, ldb, K, M, N, 1, C, ldc, accumulate, workspace, stream);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// host orchestration of the edge layer in GEMM form
// ---------------------------------------------------------------------------------------------------------------
namespace segnn {
namespace eg {

static inline int64_t align64(int64_t floats) { return (floats + 63) & ~(int64_t)63; }  // 256-byte granules

struct Plan {
  int64_t fixed_floats, per_graph_floats;
  int64_t o_wcat, o_wcat_t, o_ws_g, o_dwcat, o_dwvv, o_db2c, o_tn, o_colsum, o_chunk;  // offsets in floats
  int64_t ws_g_floats, tn_floats, colsum_floats;
};

// backward == false: only (XS, XV, Y, DV) per row
static Plan make_plan(int N, int n, bool backward, int64_t max_chunks) {
  Plan p{};
  int64_t o = 0;
  p.o_wcat = o; o += align64((int64_t)6 * n * n);
  p.o_wcat_t = o; o += align64((int64_t)6 * n * n);
  const int64_t g1 = segnn_gemm_tf32x3_workspace(2 * n, 3 * n), g2 = segnn_gemm_tf32x3_workspace(3 * n, 2 * n),
                g3 = segnn_gemm_tf32x3_workspace(n, n);
  int64_t gmax = g1 > g2 ? g1 : g2;
  if (g3 > gmax) gmax = g3;
  p.ws_g_floats = align64(gmax / 4);
  p.o_ws_g = o; o += p.ws_g_floats;
  p.o_dwcat = o; o += align64((int64_t)6 * n * n);
  p.o_dwvv = o; o += align64((int64_t)n * n);
  p.o_db2c = o; o += align64(max_chunks * 2 * n);
  const int64_t t1 = segnn_gemm_tn_tf32x3_workspace((int64_t)1 << 40, 2 * n, 3 * n),
                t2 = segnn_gemm_tn_tf32x3_workspace((int64_t)1 << 40, n, n);
  p.tn_floats = backward ? align64((t1 > t2 ? t1 : t2) / 4) : 0;
  p.o_tn = o; o += p.tn_floats;
  p.o_colsum = o;
  p.colsum_floats = 0;  // filled by the caller once the chunk size is known
  p.fixed_floats = o;
  const int64_t tiles_i = (N + kTI - 1) / kTI, tiles_j = (N + kTJ - 1) / kTJ;
  p.per_graph_floats = (int64_t)N * N * (backward ? 16 : 11) * n + (backward ? (int64_t)N * 2 * n : 0) +
                       (backward ? (tiles_j * 18 + tiles_i * 12) * N * n : 0);  // partial dP, d w_edge1 / dQ rows
  return p;
}

struct Kernels {
  decltype(&msg1_rows_kernel<0>) rows;
  decltype(&gate2_fwd_kernel<0>) g2f;
  decltype(&gate2_bwd_kernel<0>) g2b;
  decltype(&msg1_bwd_tile_kernel<0, false>) tile, tile_q;
};
template <int NCH>
static Kernels make_kernels() {
  return {msg1_rows_kernel<NCH>, gate2_fwd_kernel<NCH>, gate2_bwd_kernel<NCH>, msg1_bwd_tile_kernel<NCH, false>,
          msg1_bwd_tile_kernel<NCH, true>};
}
// the multiplicities of the BASELINE configurations (hidden 64 / 128 / 192) get compile-time offsets
static Kernels kernels_for(int n) {
  switch (n) {
    case 32: return make_kernels<32>();
    case 64: return make_kernels<64>();
    case 96: return make_kernels<96>();
    default: return make_kernels<0>();
  }
}

static int check_common(int B, int N, int n) {
  if (n < 4 || n > 96 || (n & 3) != 0) {
    set_error("segnn_edge_layer_gemm: hidden multiplicity n=%d must be a multiple of 4 in [4, 96]", n);
    return SEGNN_E_UNSUPPORTED;
  }
  if (B < 0 || N < 2) {
    set_error("segnn_edge_layer_gemm: bad sizes B=%d N=%d", B, N);
    return SEGNN_E_INVALID;
  }
  return SEGNN_OK;
}

}  // namespace eg
}  // namespace segnn

extern "C" {

int64_t segnn_edge_layer_gemm_workspace(int B, int N, int n, int backward, int64_t budget_bytes) {
  if (eg::check_common(B, N, n) != SEGNN_OK || B < 1) return -1;
  // chunks of whole graphs; at least one graph per chunk whatever the budget
  eg::Plan p = eg::make_plan(N, n, backward != 0, B);
  const int64_t colsum = backward ? segnn_colsum_workspace((int64_t)B * N, 2 * n) / 4 + 64 : 0;
  const int64_t fixed = p.fixed_floats + eg::align64(colsum);
  int64_t graphs = B;
  if (budget_bytes > 0) {
    const int64_t room = budget_bytes / 4 - fixed;
    graphs = room / (p.per_graph_floats + 64 * 12);
    if (graphs < 1) graphs = 1;
    if (graphs > B) graphs = B;
  }
  return (fixed + graphs * (p.per_graph_floats + 64 * 12)) * (int64_t)sizeof(float);
}

static int64_t graphs_per_chunk(const eg::Plan& p, int64_t colsum_floats, int64_t ws_bytes, int B) {
  const int64_t room = ws_bytes / 4 - p.fixed_floats - eg::align64(colsum_floats);
  int64_t graphs = room / (p.per_graph_floats + 64 * 12);
  if (graphs > B) graphs = B;
  return graphs;
}

int segnn_edge_layer_gemm_fwd(const float* pos, const float* mass, int B, int N, int n, const float* p, const float* q,
                              const float* w_edge1, const float* w2_ss, const float* w2_vs, const float* w2_sv,
                              const float* w2_vv, const float* b2, const float* bn_mul, const float* bn_add,
                              float* agg_out, float* moments, float* workspace, int64_t workspace_bytes,
                              segnn_stream_t stream) {
  int rc = eg::check_common(B, N, n);
  if (rc != SEGNN_OK) return rc;
  if (B == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(pos && mass && p && q && w_edge1 && w2_ss && w2_vs && w2_sv && w2_vv && b2 && agg_out && workspace,
                  "null pointer");
  SEGNN_CHECK_ARG((bn_mul == nullptr) == (bn_add == nullptr), "bn_mul and bn_add go together");
  SEGNN_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  eg::Plan pl = eg::make_plan(N, n, false, B);
  const int64_t gpc = graphs_per_chunk(pl, 0, workspace_bytes, B);
  SEGNN_CHECK_ARG(gpc >= 1, "workspace too small for one graph (segnn_edge_layer_gemm_workspace)");
  cudaStream_t s = (cudaStream_t)stream;
  float* wcat = workspace + pl.o_wcat;
  float* ws_g = workspace + pl.o_ws_g;
  float* chunk = workspace + pl.fixed_floats;
  eg::build_wcat_kernel<<<(6 * n * n + 255) / 256, 256, 0, s>>>(n, w2_ss, w2_vs, w2_sv, nullptr, nullptr, nullptr, wcat,
                                                              nullptr);
  SEGNN_CHECK_LAUNCH();
  const int NT = (n + 31) & ~31;
  const size_t smem_node = sizeof(float) * ((size_t)eg::kTileO * 8 + (size_t)eg::kSY * 18 * NT);
  const eg::Kernels kn = eg::kernels_for(n);
  cudaFuncSetAttribute(kn.g2f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_node);
  for (int64_t g0 = 0; g0 < B; g0 += gpc) {
    const int64_t gc = (B - g0 < gpc) ? B - g0 : gpc;
    const int64_t rows = gc * N * N;
    float* xs = chunk;
    float* xv = xs + eg::align64(rows * 2 * n);
    float* yy = xv + eg::align64(rows * 3 * n);
    float* dv = yy + eg::align64(rows * 3 * n);
    eg::RowArgs ra{pos, mass, p, q, w_edge1, N, n, g0 * N};
    const int64_t blocks1 = gc * ((N + eg::kRecv - 1) / eg::kRecv);
    kn.rows<<<(unsigned)blocks1, dim3(NT, 4), 0, s>>>(ra, xs, xv);
    SEGNN_CHECK_LAUNCH();
    rc = segnn_gemm_tf32x3(xs, 2 * n, wcat, 3 * n, rows, 2 * n, 3 * n, yy, 3 * n, ws_g, stream);
    if (rc != SEGNN_OK) return rc;
    rc = segnn_gemm_tf32x3(xv, n, w2_vv, n, 3 * rows, n, n, dv, n, ws_g, stream);
    if (rc != SEGNN_OK) return rc;
    kn.g2f<<<(unsigned)(gc * N), dim3(NT, eg::kSY), smem_node, s>>>(ra, yy, dv, b2, bn_mul, bn_add, agg_out,
                                                                                 moments);
    SEGNN_CHECK_LAUNCH();
  }
  return SEGNN_OK;
}

int segnn_edge_layer_gemm_bwd_phases(const float* pos, const float* mass, int B, int N, int n, const float* p,
                                     const float* q, const float* w_edge1, const float* w2_ss, const float* w2_vs,
                                     const float* w2_sv, const float* w2_vv, const float* b2, const float* w2t_ss,
                                     const float* w2t_vs, const float* w2t_sv, const float* w2t_vv, const float* bn_a,
                                     const float* bn_b, const float* bn_c, const float* dagg, float* dP, float* dQ,
                                     float* dw2_ss, float* dw2_vs, float* dw2_sv, float* dw2_vv, float* db2,
                                     float* dwe_partial, float* workspace, int64_t workspace_bytes,
                                     const float* fwd_workspace, int64_t fwd_workspace_bytes, int phases,
                                     segnn_stream_t stream) {
  SEGNN_CHECK_ARG(phases >= 1 && phases <= 7, "phases: bit 0 rows + gate backward, bit 1 weight gradients, bit 2 data gradients");
  const bool ph_rows = (phases & 1) != 0, ph_w = (phases & 2) != 0, ph_d = (phases & 4) != 0;
  int rc = eg::check_common(B, N, n);
  if (rc != SEGNN_OK) return rc;
  if (B == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(pos && mass && p && q && w_edge1 && w2_ss && w2_vs && w2_sv && w2_vv && b2 && w2t_ss && w2t_vs &&
                      w2t_sv && w2t_vv && bn_a && bn_b && bn_c && dagg && dP && dQ && dw2_ss && dw2_vs && dw2_sv &&
                      dw2_vv && db2 && dwe_partial && workspace,
                  "null pointer");
  SEGNN_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  eg::Plan pl = eg::make_plan(N, n, true, B);
  const int64_t colsum_floats = segnn_colsum_workspace((int64_t)B * N, 2 * n) / 4 + 64;
  const int64_t gpc = graphs_per_chunk(pl, colsum_floats, workspace_bytes, B);
  SEGNN_CHECK_ARG(gpc >= 1, "workspace too small for one graph (segnn_edge_layer_gemm_workspace)");
  SEGNN_CHECK_ARG(phases == 7 || gpc >= B, "separate phases need a workspace that holds every graph in one chunk");
  cudaStream_t s = (cudaStream_t)stream;
  float* wcat = workspace + pl.o_wcat;
  float* wcat_t = workspace + pl.o_wcat_t;
  float* ws_g = workspace + pl.o_ws_g;
  float* dwcat = workspace + pl.o_dwcat;
  float* dwvv = workspace + pl.o_dwvv;
  float* db2c = workspace + pl.o_db2c;
  float* ws_tn = workspace + pl.o_tn;
  float* ws_col = workspace + pl.o_colsum;
  float* chunk = ws_col + eg::align64(colsum_floats);
  if (ph_rows) {
    eg::build_wcat_kernel<<<(6 * n * n + 255) / 256, 256, 0, s>>>(n, w2_ss, w2_vs, w2_sv, w2t_ss, w2t_vs, w2t_sv, wcat,
                                                                wcat_t);
    SEGNN_CHECK_LAUNCH();
  }
  const int NT = (n + 31) & ~31;
  const size_t smem_node = sizeof(float) * ((size_t)eg::kTileO * 8 + (size_t)eg::kSY * 18 * NT);
  const eg::Kernels kn = eg::kernels_for(n);
  cudaFuncSetAttribute(kn.g2b, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_node);
  const int tiles_i = (N + eg::kTI - 1) / eg::kTI, tiles_j = (N + eg::kTJ - 1) / eg::kTJ;
  const int tile_warps = (n + 7) / 8;  // a warp of the tile kernel covers 8 channels x 4 sender lanes
  // small grids (training-size graphs) stage the sender projections of a tile in shared memory
  const bool stage_q = (int64_t)(gpc < B ? gpc : B) * tiles_i * tiles_j < 2 * 148;
  const auto k_tile = stage_q ? kn.tile_q : kn.tile;
  const size_t smem_tile = sizeof(float) * ((size_t)eg::kTI * eg::kTJ * 6 +
                                            (stage_q ? 2 : 1) * (size_t)eg::kTJ * (12 * tile_warps * 8 + 8));
  cudaFuncSetAttribute(k_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tile);
  // rows kept by the forward call (XS, XV, Y, DV of ALL graphs, i.e. the forward ran as one chunk): no recompute
  const float *kxs = nullptr, *kxv = nullptr, *kyy = nullptr, *kdv = nullptr;
  if (fwd_workspace != nullptr) {
    eg::Plan fp = eg::make_plan(N, n, false, B);
    SEGNN_CHECK_ARG(graphs_per_chunk(fp, 0, fwd_workspace_bytes, B) >= B,
                    "fwd_workspace does not hold the rows of all graphs (the forward call was chunked)");
    const int64_t all = (int64_t)B * N * N;
    kxs = fwd_workspace + fp.fixed_floats;
    kxv = kxs + eg::align64(all * 2 * n);
    kyy = kxv + eg::align64(all * 3 * n);
    kdv = kyy + eg::align64(all * 3 * n);
  }
  int chunk_idx = 0;
  for (int64_t g0 = 0; g0 < B; g0 += gpc, ++chunk_idx) {
    const int64_t gc = (B - g0 < gpc) ? B - g0 : gpc;
    const int64_t rows = gc * N * N;
    float* xs = chunk;
    float* xv = xs + eg::align64(rows * 2 * n);
    float* yy = xv + eg::align64(rows * 3 * n);
    float* dv = yy + eg::align64(rows * 3 * n);
    float* dxs = dv + eg::align64(rows * 3 * n);
    float* dxv = dxs + eg::align64(rows * 2 * n);
    float* db2_rows = dxv + eg::align64(rows * 3 * n);
    eg::RowArgs ra{pos, mass, p, q, w_edge1, N, n, g0 * N};
    const int64_t blocks1 = gc * ((N + eg::kRecv - 1) / eg::kRecv);
    const float *cxs = xs, *cxv = xv, *cyy = yy, *cdv = dv;
    if (kxs != nullptr) {
      const int64_t r0 = g0 * N * N;
      cxs = kxs + r0 * 2 * n;
      cxv = kxv + r0 * 3 * n;
      cyy = kyy + r0 * 3 * n;
      cdv = kdv + r0 * 3 * n;
    } else if (ph_rows) {
      // recompute: message_layer_2 input rows and pre-activations
      kn.rows<<<(unsigned)blocks1, dim3(NT, 4), 0, s>>>(ra, xs, xv);
      SEGNN_CHECK_LAUNCH();
      rc = segnn_gemm_tf32x3(xs, 2 * n, wcat, 3 * n, rows, 2 * n, 3 * n, yy, 3 * n, ws_g, stream);
      if (rc != SEGNN_OK) return rc;
      rc = segnn_gemm_tf32x3(xv, n, w2_vv, n, 3 * rows, n, n, dv, n, ws_g, stream);
      if (rc != SEGNN_OK) return rc;
    }
    if (ph_rows) {
      // gate backward: (Y, DV) -> (dY, dDV), in place when recomputed
      kn.g2b<<<(unsigned)(gc * N), dim3(NT, eg::kSY), smem_node, s>>>(ra, cyy, cdv, yy, dv, b2, bn_a, bn_b,
                                                                                   bn_c, dagg, db2_rows);
      SEGNN_CHECK_LAUNCH();
      rc = segnn_colsum(db2_rows, nullptr, gc * N, 2 * n, 0, ws_col, db2c + (int64_t)chunk_idx * 2 * n, stream);
      if (rc != SEGNN_OK) return rc;
    }
    if (ph_w) {
      // weight gradients: K = rows
      rc = segnn_gemm_tn_tf32x3(cxs, 2 * n, yy, 3 * n, rows, 2 * n, 3 * n, dwcat, 3 * n, chunk_idx > 0, ws_tn, stream);
      if (rc != SEGNN_OK) return rc;
      // dW_vv = sum over the three components of XV_k^T dDV_k: the components sit side by side in a row
      if ((n & 31) == 0)
        rc = segnn_gemm_tn_grouped_tf32x3(cxv, 3 * n, dv, 3 * n, rows, n, n, 3, dwvv, n, chunk_idx > 0, ws_tn, stream);
      else
        rc = segnn_gemm_tn_tf32x3(cxv, n, dv, n, 3 * rows, n, n, dwvv, n, chunk_idx > 0, ws_tn, stream);
      if (rc != SEGNN_OK) return rc;
    }
    if (!ph_d) continue;
    // data gradients
    rc = segnn_gemm_tf32x3(yy, 3 * n, wcat_t, 2 * n, rows, 3 * n, 2 * n, dxs, 2 * n, ws_g, stream);
    if (rc != SEGNN_OK) return rc;
    rc = segnn_gemm_tf32x3(dv, n, w2t_vv, n, 3 * rows, n, n, dxv, n, ws_g, stream);
    if (rc != SEGNN_OK) return rc;
    // gate / combine backward of message_layer_1 in one pass over the rows: per-tile partial sums, fixed-order reduction
    const int64_t cn = gc * N;
    float* dp_part = db2_rows + eg::align64(cn * 2 * n);
    float* dq_part = dp_part + eg::align64(tiles_j * cn * 12 * n);
    float* dwe_part = dq_part + eg::align64(tiles_i * cn * 12 * n);
    k_tile<<<(unsigned)(gc * tiles_i * tiles_j), tile_warps * 32, smem_tile, s>>>(
        ra, dxs, dxv, tiles_i, tiles_j, cn, dp_part, dq_part, dwe_part);
    SEGNN_CHECK_LAUNCH();
    const int64_t node0 = g0 * N;
    eg::tile_reduce_kernel<<<(unsigned)((cn * 12 * n + 255) / 256), 256, 0, s>>>(dp_part, tiles_j, cn, 12 * n,
                                                                               dP + node0 * 12 * n);
    SEGNN_CHECK_LAUNCH();
    eg::tile_reduce_kernel<<<(unsigned)((cn * 12 * n + 255) / 256), 256, 0, s>>>(dq_part, tiles_i, cn, 12 * n,
                                                                               dQ + node0 * 12 * n);
    SEGNN_CHECK_LAUNCH();
    eg::tile_reduce_kernel<<<(unsigned)((cn * 6 * n + 255) / 256), 256, 0, s>>>(dwe_part, tiles_j, cn, 6 * n,
                                                                              dwe_partial + node0 * 6 * n);
    SEGNN_CHECK_LAUNCH();
  }
  if (!ph_w) return SEGNN_OK;
  eg::scatter_w2_grads_kernel<<<(6 * n * n + 255) / 256, 256, 0, s>>>(n, dwcat, dwvv, db2c, chunk_idx, dw2_ss, dw2_vs,
                                                                    dw2_sv, dw2_vv, db2);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_edge_layer_gemm_bwd(const float* pos, const float* mass, int B, int N, int n, const float* p, const float* q,
                              const float* w_edge1, const float* w2_ss, const float* w2_vs, const float* w2_sv,
                              const float* w2_vv, const float* b2, const float* w2t_ss, const float* w2t_vs,
                              const float* w2t_sv, const float* w2t_vv, const float* bn_a, const float* bn_b,
                              const float* bn_c, const float* dagg, float* dP, float* dQ, float* dw2_ss, float* dw2_vs,
                              float* dw2_sv, float* dw2_vv, float* db2, float* dwe_partial, float* workspace,
                              int64_t workspace_bytes, const float* fwd_workspace, int64_t fwd_workspace_bytes,
                              segnn_stream_t stream) {
  return segnn_edge_layer_gemm_bwd_phases(pos, mass, B, N, n, p, q, w_edge1, w2_ss, w2_vs, w2_sv, w2_vv, b2, w2t_ss,
                                          w2t_vs, w2t_sv, w2t_vv, bn_a, bn_b, bn_c, dagg, dP, dQ, dw2_ss, dw2_vs,
                                          dw2_sv, dw2_vv, db2, dwe_partial, workspace, workspace_bytes, fwd_workspace,
                                          fwd_workspace_bytes, 7, stream);
}

}  // extern "C"
