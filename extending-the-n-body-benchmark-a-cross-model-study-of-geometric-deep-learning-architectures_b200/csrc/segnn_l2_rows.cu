// lmax_h = 2 (hidden irreps n x 0e + n x 1o + n x 2e, BASELINE configuration 3) edge layer in GEMM form, inference.
//
// The generic-irreps path (segnn_generic.cu) follows the reference literally: message_layer_1 output [E, 5n + ...] ->
// gate -> per output irrep an expansion kernel -> GEMM -> scatter -> gate -> BatchNorm -> aggregate, every step a
// table-driven kernel over an [E, .] tensor in HBM.  Here the per-edge work around the weight contraction of
// message_layer_2 is two kernels specialised for this irreps structure (the 7 instruction types (l1, l2, lo) of a
// 0e+1o+2e feature with a 0e+1o attribute), models/segnn/segnn.py:264-284, o3_building_blocks.py:150-203:
//   l2_msg_rows_kernel        hoisted message_layer_1 (P_i + Q_j per instruction, coupling with the edge attribute,
//                             additional scalars, bias), e3nn Gate, and the coupling of the gated message with the edge
//                             attribute for message_layer_2 -> the three GEMM operands A0 [E][K0], A1 [3E][K1], A2 [5E][K2]
//   (segnn_gemm_tf32x3 x 3)   Y_b = A_b * stacked path weights, fp32-accurate on tcgen05
//   l2_gate_aggregate_kernel  bias, e3nn Gate, sum over senders, folded eval BatchNorm -> agg [nodes][9n] (e3nn layout)
// Edge rows are (graph, receiver, sender) with the diagonal kept and masked.  The real Wigner 3j tables come from the
// host (cg.py) so that signs and normalisation are the oracle's; M_t[i][k] = sum_j C_t[i][j][k] attr[j] is built once
// per edge in shared memory and shared by all channels.
#include "segnn_common.cuh"

namespace segnn {
namespace l2 {

// instruction types t = 0..6: (l1, l2, lo) = (0,0,0) (0,1,1) (1,0,1) (1,1,0) (1,1,2) (2,0,2) (2,1,1)
// dimensions (2 l1 + 1, 2 l2 + 1, 2 lo + 1) = (1,1,1) (1,3,3) (3,1,3) (3,3,1) (3,3,5) (5,1,5) (5,3,3)
constexpr int kTypes = 7;
constexpr int kWarps = 4;
constexpr int kMaxCpt = 2;  // channels a lane processes together (32 lanes x 2 per pass, passes until n is covered)

struct RowsArgs {
  const float *pos, *mass;
  const float* Y[3];     // node-level products by input degree l1: Y[l1][(node * (2 l1 + 1) + i)][ldy[l1]]
  long long ldy[3];
  int B, N, n;
  int yoff[kTypes][2];   // column of (type, role) in a row of Y[l1(type)]: role 0 = receiver x_i, 1 = sender x_j
  const float* cg;       // [7][5][3][5] net couplings
  const float* w_add0;   // [2][3n] weights of the additional scalars -> 0e outputs
  const float* w_add1;   // [2][n]                                   -> 1o outputs
  const float* bias1;    // [3n] bias of the 0e outputs of message_layer_1
  int koff[kTypes];      // column of (type) inside its message_layer_2 GEMM operand
  long long lda0, lda1, lda2;
  float *A0, *A1, *A2;
  long long node0;       // first node of the chunk
  int graphs;            // graphs in the chunk
};

// z[c][k] += sum_i M[i][k] S[c][i] for the kMaxCpt channels of a lane at once: every coupling coefficient is read from
// shared memory once per edge and lane, not once per channel
template <int CW, int D1, int DO>
__device__ __forceinline__ void couple(const float* __restrict__ M, const float (&S)[CW][D1], float (&z)[CW][DO]) {
#pragma unroll
  for (int i = 0; i < D1; ++i)
#pragma unroll
    for (int k = 0; k < DO; ++k) {
      const float m = M[i * 5 + k];
#pragma unroll
      for (int c = 0; c < CW; ++c) z[c][k] = fmaf(m, S[c][i], z[c][k]);
    }
}
// S[c][i] = P_i + Q_j of one instruction type for output channel w0 + 32 c: component i of node r sits in row
// r * D1 + i of the degree's Y, consecutive channels are consecutive floats (coalesced over the lanes)
template <int CW, int D1>
__device__ __forceinline__ void load_sum(const float* __restrict__ yi, const float* __restrict__ yj, long long ld,
                                         const int (&w)[CW], float (&S)[CW][D1]) {
#pragma unroll
  for (int c = 0; c < CW; ++c)
#pragma unroll
    for (int i = 0; i < D1; ++i) S[c][i] = yi[i * ld + w[c]] + yj[i * ld + w[c]];
}

struct EdgeCtx {
  const float *yi0, *yi1, *yi2, *yj0, *yj1, *yj2, *M;
  float len, mm;
  long long row;
  bool live;  // false: the lane has no edge in this pass (paired remainder pass of an odd edge count)
};

// CW channels of a lane at once (u = lane + 32 (c0 + c)): message_layer_1 coupling, gate, message_layer_2 coupling
template <int CW>
__device__ __forceinline__ void edge_channels(const RowsArgs& a, const EdgeCtx& e, int c0, int lane) {
  const int n = a.n;
  const float *yi0 = e.yi0, *yi1 = e.yi1, *yi2 = e.yi2, *yj0 = e.yj0, *yj1 = e.yj1, *yj2 = e.yj2, *M = e.M;
  const float len = e.len, mm = e.mm;
  const long long row = e.row;
    // the lane's channels of this pass, u = lane + 32 (c0 + c); channels past n compute on channel n - 1 and are not stored
    int uc[CW];
    bool on[CW];
#pragma unroll
    for (int c = 0; c < CW; ++c) {
      on[c] = e.live && lane + 32 * (c0 + c) < n;
      uc[c] = on[c] ? lane + 32 * (c0 + c) : n - 1;
    }
    // ---- message_layer_1: 0e outputs (scalar u, gate of 1o_u, gate of 2e_u) -------------------------------------
    float z0[3][CW][1];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      int w[CW];
      float S0[CW][1], S3[CW][3];
#pragma unroll
      for (int c = 0; c < CW; ++c) {
        w[c] = q * n + uc[c];
        // bias + additional scalars (type (0,0,0))
        z0[q][c][0] = a.bias1[w[c]] + M[0] * (a.w_add0[w[c]] * len + a.w_add0[3 * n + w[c]] * mm);
      }
      load_sum<CW, 1>(yi0 + a.yoff[0][0], yj0 + a.yoff[0][1], a.ldy[0], w, S0);
      couple<CW, 1, 1>(M + 0 * 25, S0, z0[q]);
      load_sum<CW, 3>(yi1 + a.yoff[3][0], yj1 + a.yoff[3][1], a.ldy[1], w, S3);
      couple<CW, 3, 1>(M + 3 * 25, S3, z0[q]);
    }
    // ---- 1o and 2e outputs ---------------------------------------------------------------------------------------
    float z1[CW][3], z2[CW][5];
#pragma unroll
    for (int c = 0; c < CW; ++c) {
#pragma unroll
      for (int k = 0; k < 3; ++k) z1[c][k] = 0.f;
#pragma unroll
      for (int k = 0; k < 5; ++k) z2[c][k] = 0.f;
    }
    {
      float S1[CW][1], S3[CW][3], S5[CW][5];
      load_sum<CW, 1>(yi0 + a.yoff[1][0], yj0 + a.yoff[1][1], a.ldy[0], uc, S1);
#pragma unroll
      for (int c = 0; c < CW; ++c)
        S1[c][0] += a.w_add1[uc[c]] * len + a.w_add1[n + uc[c]] * mm;  // additional scalars: type (0,1,1)
      couple<CW, 1, 3>(M + 1 * 25, S1, z1);
      load_sum<CW, 3>(yi1 + a.yoff[2][0], yj1 + a.yoff[2][1], a.ldy[1], uc, S3);
      couple<CW, 3, 3>(M + 2 * 25, S3, z1);
      load_sum<CW, 5>(yi2 + a.yoff[6][0], yj2 + a.yoff[6][1], a.ldy[2], uc, S5);
      couple<CW, 5, 3>(M + 6 * 25, S5, z1);
      load_sum<CW, 3>(yi1 + a.yoff[4][0], yj1 + a.yoff[4][1], a.ldy[1], uc, S3);
      couple<CW, 3, 5>(M + 4 * 25, S3, z2);
      load_sum<CW, 5>(yi2 + a.yoff[5][0], yj2 + a.yoff[5][1], a.ldy[2], uc, S5);
      couple<CW, 5, 5>(M + 5 * 25, S5, z2);
    }
    // ---- e3nn Gate -------------------------------------------------------------------------------------------------
    float s[CW][1], v[CW][3], qq[CW][5];
#pragma unroll
    for (int c = 0; c < CW; ++c) {
      s[c][0] = silu_gate(z0[0][c][0]);
      const float g1 = sig_gate(z0[1][c][0]), g2 = sig_gate(z0[2][c][0]);
#pragma unroll
      for (int k = 0; k < 3; ++k) v[c][k] = g1 * z1[c][k];
#pragma unroll
      for (int k = 0; k < 5; ++k) qq[c][k] = g2 * z2[c][k];
    }
    // ---- message_layer_2: coupling of the gated message with the edge attribute -> GEMM operand rows ---------------
    {
      float o0[CW][1], o1[CW][1];
#pragma unroll
      for (int c = 0; c < CW; ++c) o0[c][0] = o1[c][0] = 0.f;
      couple<CW, 1, 1>(M + 0 * 25, s, o0);
      couple<CW, 3, 1>(M + 3 * 25, v, o1);
      float* r0 = a.A0 + row * a.lda0;
#pragma unroll
      for (int c = 0; c < CW; ++c)
        if (on[c]) {
          r0[a.koff[0] + uc[c]] = o0[c][0];
          r0[a.koff[3] + uc[c]] = o1[c][0];
        }
    }
    {
      float p0[CW][3], p1[CW][3], p2[CW][3];
#pragma unroll
      for (int c = 0; c < CW; ++c)
#pragma unroll
        for (int k = 0; k < 3; ++k) p0[c][k] = p1[c][k] = p2[c][k] = 0.f;
      couple<CW, 1, 3>(M + 1 * 25, s, p0);
      couple<CW, 3, 3>(M + 2 * 25, v, p1);
      couple<CW, 5, 3>(M + 6 * 25, qq, p2);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float* r1 = a.A1 + (row * 3 + k) * a.lda1;
#pragma unroll
        for (int c = 0; c < CW; ++c)
          if (on[c]) {
            r1[a.koff[1] + uc[c]] = p0[c][k];
            r1[a.koff[2] + uc[c]] = p1[c][k];
            r1[a.koff[6] + uc[c]] = p2[c][k];
          }
      }
    }
    {
      float p0[CW][5], p1[CW][5];
#pragma unroll
      for (int c = 0; c < CW; ++c)
#pragma unroll
        for (int k = 0; k < 5; ++k) p0[c][k] = p1[c][k] = 0.f;
      couple<CW, 3, 5>(M + 4 * 25, v, p0);
      couple<CW, 5, 5>(M + 5 * 25, qq, p1);
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        float* r2 = a.A2 + (row * 5 + k) * a.lda2;
#pragma unroll
        for (int c = 0; c < CW; ++c)
          if (on[c]) {
            r2[a.koff[4] + uc[c]] = p0[c][k];
            r2[a.koff[5] + uc[c]] = p1[c][k];
          }
      }
    }
}

__global__ void __launch_bounds__(kWarps * 32, 4) l2_msg_rows_kernel(const RowsArgs a) {
  __shared__ float cg_s[kTypes * 75];
  __shared__ __align__(16) float Ms[kWarps][2][kTypes * 25];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = a.n, N = a.N;
  for (int t = threadIdx.x; t < kTypes * 75; t += blockDim.x) cg_s[t] = a.cg[t];
  __syncthreads();
  const long long rl = blockIdx.x;  // receiver, local to the chunk
  const long long gl = rl / N;
  const int ir = (int)(rl - gl * N);
  const long long base = a.node0 + gl * N;
  const float* yi0 = a.Y[0] + (base + ir) * a.ldy[0];
  const float* yi1 = a.Y[1] + (base + ir) * 3 * a.ldy[1];
  const float* yi2 = a.Y[2] + (base + ir) * 5 * a.ldy[2];
  const float pix = a.pos[(base + ir) * 3 + 0], piy = a.pos[(base + ir) * 3 + 1], piz = a.pos[(base + ir) * 3 + 2];
  const float mi = a.mass[base + ir];
  // Two edges (senders j and j + kWarps) per iteration: the passes over full groups of 32 channels run per edge; the
  // remainder pass (n mod 32 channels, 9 of 32 lanes for the n = 73 of BASELINE configuration 3) runs ONCE for both
  // edges when it fits half a warp, lanes 0-15 on the first edge and 16-31 on the second.
  const int full = n / 32, rem = n - 32 * full;
  const bool paired = rem > 0 && rem <= 16;
  for (int j = warp; j < N; j += 2 * kWarps) {
    const int jb = j + kWarps;
    const bool live_b = jb < N;
    float len2[2], mm2[2];
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      // ---- per edge: attribute, additional scalars, the seven coupling matrices ----------------------------------
      const int jj = h == 0 ? j : (live_b ? jb : j);
      float ux, uy, uz, len;
      unit_vec(a.pos[(base + jj) * 3 + 0] - pix, a.pos[(base + jj) * 3 + 1] - piy, a.pos[(base + jj) * 3 + 2] - piz, ux,
               uy, uz, len);
      const float at[4] = {kY0, kY1 * ux, kY1 * uy, kY1 * uz};
      len2[h] = len;
      mm2[h] = a.mass[base + jj] * mi;
      float* M = Ms[warp][h];
      for (int idx = lane; idx < kTypes * 25; idx += 32) {
        const int t = idx / 25, r = idx - t * 25, i = r / 5, k = r - i * 5;
        const int d2 = t == 0 || t == 2 || t == 5 ? 1 : 3;   // kD2
        const int jo = d2 == 1 ? 0 : 1;                      // offset of the attribute irrep (0e at 0, 1o at 1)
        float m = 0.f;
        for (int q = 0; q < d2; ++q) m = fmaf(cg_s[t * 75 + (i * 3 + q) * 5 + k], at[jo + q], m);
        M[idx] = m;
      }
    }
    __syncwarp();
    auto ctx = [&](int h) {
      const int jj = (h != 0 && live_b) ? jb : j;
      return EdgeCtx{yi0, yi1, yi2, a.Y[0] + (base + jj) * a.ldy[0], a.Y[1] + (base + jj) * 3 * a.ldy[1],
                     a.Y[2] + (base + jj) * 5 * a.ldy[2], Ms[warp][h != 0], h != 0 ? len2[1] : len2[0],
                     h != 0 ? mm2[1] : mm2[0], rl * N + jj, h == 0 || live_b};
    };
    // one channel per lane and pass: two at once (edge_channels<2>, every coupling coefficient read once for both)
    // was measured slower -- 168+ registers against 128 cost more occupancy than the shared-memory reads saved
    const int npass = paired ? full : full + (rem > 0 ? 1 : 0);
#pragma unroll 1
    for (int c0 = 0; c0 < npass; ++c0) {
#pragma unroll 1
      for (int h = 0; h < (live_b ? 2 : 1); ++h) edge_channels<1>(a, ctx(h), c0, lane);
    }
    if (paired)  // lanes 0-15: first edge, 16-31: second; lanes whose channel 32 full + (lane & 15) >= n store nothing
      edge_channels<1>(a, ctx(lane >> 4), full, lane & 15);
  }
}

// block = one receiver; thread = (channel u, sender lane); sums the gated message_layer_2 outputs over the senders
constexpr int kGY = 4;

__global__ void __launch_bounds__(96 * kGY)
    l2_gate_aggregate_kernel(int N, int n, const float* __restrict__ Y0, long long ld0, const float* __restrict__ Y1,
                             long long ld1, const float* __restrict__ Y2, long long ld2,
                             const float* __restrict__ bias2, const float* __restrict__ bn_mul,
                             const float* __restrict__ bn_add, long long node0, float* __restrict__ agg) {
  extern __shared__ float red[];  // [kGY][9][NT]
  const int u = threadIdx.x, y = threadIdx.y, NT = blockDim.x;
  const long long rl = blockIdx.x;
  const int ir = (int)(rl % N);
  const bool act = u < n;
  float acc[9];
#pragma unroll
  for (int v = 0; v < 9; ++v) acc[v] = 0.f;
  if (act) {
    const float b0 = bias2[u], b1 = bias2[n + u], b2 = bias2[2 * n + u];
#pragma unroll 4
    for (int j = y; j < N; j += kGY) {
      const float valid = j == ir ? 0.f : 1.f;
      const long long row = rl * N + j;
      const float* r0 = Y0 + row * ld0;
      const float s = r0[u] + b0, g1 = r0[n + u] + b1, g2 = r0[2 * n + u] + b2;
      float v1[3], v2[5];
#pragma unroll
      for (int k = 0; k < 3; ++k) v1[k] = Y1[(row * 3 + k) * ld1 + u];
#pragma unroll
      for (int k = 0; k < 5; ++k) v2[k] = Y2[(row * 5 + k) * ld2 + u];
      const float ms = valid * silu_gate(s), m1 = valid * sig_gate(g1), m2 = valid * sig_gate(g2);
      acc[0] += ms;
#pragma unroll
      for (int k = 0; k < 3; ++k) acc[1 + k] = fmaf(m1, v1[k], acc[1 + k]);
#pragma unroll
      for (int k = 0; k < 5; ++k) acc[4 + k] = fmaf(m2, v2[k], acc[4 + k]);
    }
  }
#pragma unroll
  for (int v = 0; v < 9; ++v) red[(y * 9 + v) * NT + u] = acc[v];
  __syncthreads();
  if (y == 0 && act) {
#pragma unroll
    for (int v = 0; v < 9; ++v) {
      float s = red[v * NT + u];
      for (int yy = 1; yy < kGY; ++yy) s += red[(yy * 9 + v) * NT + u];
      acc[v] = s;
    }
    // e3nn layout of n x 0e + n x 1o + n x 2e: [u] | [n + 3u + k] | [4n + 5u + k]
    float* o = agg + (node0 + rl) * 9 * n;
    const int c0 = u, c1 = n + 3 * u, c2 = 4 * n + 5 * u;
    const float deg = (float)(N - 1);
    auto fold = [&](int col, float val) {
      return bn_mul != nullptr ? fmaf(val, bn_mul[col], deg * bn_add[col]) : val;
    };
    o[c0] = fold(c0, acc[0]);
#pragma unroll
    for (int k = 0; k < 3; ++k) o[c1 + k] = fold(c1 + k, acc[1 + k]);
#pragma unroll
    for (int k = 0; k < 5; ++k) o[c2 + k] = fold(c2 + k, acc[4 + k]);
  }
}

// x [nodes][9n] in e3nn layout (n x 0e | n x 1o | n x 2e, index u (2l + 1) + i) -> planar rows per degree:
// xp_l[(node * (2l + 1) + i)][ld] with the channel u contiguous: the A operands of the node-level GEMMs
__global__ void l2_planarize_kernel(const float* __restrict__ x, long long nodes, int n, long long ld,
                                    float* __restrict__ x0, float* __restrict__ x1, float* __restrict__ x2) {
  const long long total = nodes * 9 * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long node = idx / (9 * n);
    const int c = (int)(idx - node * 9 * n);
    const float v = x[idx];
    if (c < n) {
      x0[node * ld + c] = v;
    } else if (c < 4 * n) {
      const int u = (c - n) / 3, i = (c - n) - 3 * u;
      x1[(node * 3 + i) * ld + u] = v;
    } else {
      const int u = (c - 4 * n) / 5, i = (c - 4 * n) - 5 * u;
      x2[(node * 5 + i) * ld + u] = v;
    }
  }
}

}  // namespace l2
}  // namespace segnn

using namespace segnn;

extern "C" {

int segnn_l2_msg_rows(const float* pos, const float* mass, int graphs, int N, int n, int64_t node0, const float* Y0,
                      int64_t ldy0, const float* Y1, int64_t ldy1, const float* Y2, int64_t ldy2, const int* yoff,
                      const float* cg, const float* w_add0, const float* w_add1,
                      const float* bias1, const int* koff, int64_t lda0, int64_t lda1, int64_t lda2, float* A0,
                      float* A1, float* A2, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(graphs >= 0 && N >= 2 && n >= 1 && n <= 96 && node0 >= 0, "bad sizes");
  if (graphs == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(pos && mass && Y0 && Y1 && Y2 && yoff && cg && w_add0 && w_add1 && bias1 && koff && A0 && A1 && A2,
                  "null pointer");
  SEGNN_CHECK_ARG(lda0 >= 2 * n && lda1 >= 3 * n && lda2 >= 2 * n, "leading dimensions too small");
  l2::RowsArgs a{};
  a.pos = pos; a.mass = mass;
  a.Y[0] = Y0; a.Y[1] = Y1; a.Y[2] = Y2;
  a.ldy[0] = ldy0; a.ldy[1] = ldy1; a.ldy[2] = ldy2;
  a.B = graphs; a.N = N; a.n = n;
  for (int t = 0; t < l2::kTypes; ++t) {
    a.yoff[t][0] = yoff[2 * t];
    a.yoff[t][1] = yoff[2 * t + 1];
    a.koff[t] = koff[t];
  }
  a.cg = cg; a.w_add0 = w_add0; a.w_add1 = w_add1; a.bias1 = bias1;
  a.lda0 = lda0; a.lda1 = lda1; a.lda2 = lda2;
  a.A0 = A0; a.A1 = A1; a.A2 = A2;
  a.node0 = node0; a.graphs = graphs;
  const long long blocks = (long long)graphs * N;
  SEGNN_CHECK_ARG(blocks <= 0x7fffffff, "too many receivers per call");
  l2::l2_msg_rows_kernel<<<(unsigned)blocks, l2::kWarps * 32, 0, (cudaStream_t)stream>>>(a);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_l2_planarize(const float* x, int64_t nodes, int n, int64_t ld, float* x0, float* x1, float* x2,
                       segnn_stream_t stream) {
  SEGNN_CHECK_ARG(nodes >= 0 && n >= 1 && ld >= n, "bad sizes");
  if (nodes == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(x && x0 && x1 && x2, "null pointer");
  const long long total = nodes * 9 * n;
  const long long blocks = (total + 255) / 256;
  l2::l2_planarize_kernel<<<(unsigned)(blocks < 2368 ? blocks : 2368), 256, 0, (cudaStream_t)stream>>>(x, nodes, n, ld, x0,
                                                                                                    x1, x2);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_l2_gate_aggregate(int graphs, int N, int n, int64_t node0, const float* Y0, int64_t ld0, const float* Y1,
                            int64_t ld1, const float* Y2, int64_t ld2, const float* bias2, const float* bn_mul,
                            const float* bn_add, float* agg, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(graphs >= 0 && N >= 2 && n >= 1 && n <= 96 && node0 >= 0, "bad sizes");
  if (graphs == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(Y0 && Y1 && Y2 && bias2 && agg, "null pointer");
  SEGNN_CHECK_ARG((bn_mul == nullptr) == (bn_add == nullptr), "bn_mul and bn_add go together");
  SEGNN_CHECK_ARG(ld0 >= 3 * n && ld1 >= n && ld2 >= n, "leading dimensions too small");
  const int NT = (n + 31) & ~31;
  const size_t smem = sizeof(float) * l2::kGY * 9 * NT;
  const long long blocks = (long long)graphs * N;
  SEGNN_CHECK_ARG(blocks <= 0x7fffffff, "too many receivers per call");
  l2::l2_gate_aggregate_kernel<<<(unsigned)blocks, dim3(NT, l2::kGY), smem, (cudaStream_t)stream>>>(
      N, n, Y0, ld0, Y1, ld1, Y2, ld2, bias2, bn_mul, bn_add, node0, agg);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

}  // extern "C"
