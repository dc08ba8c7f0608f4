// Generic-irreps SEGNN path (fp32, inference): the reference's formulation -- materialised edge list order, gathered
// message inputs, one FullyConnectedTensorProduct per call, e3nn Gate, scatter-sum -- as plain CUDA kernels that work
// for ANY hidden irreps (e.g. lmax_h = 2, BASELINE config 3).  It exists for parity coverage of configurations the
// fused kernels are not specialised for; it is not tuned (per-edge tensors live in HBM like in the reference).
//   models/segnn/o3_building_blocks.py:150-162 (tensor product + rescale + bias), :197-203 (gate),
//   models/segnn/segnn.py:264-304 (message / aggregate / update).
#include "segnn_common.cuh"

namespace segnn {

constexpr int kInstrInts = 9;    // off1, mul1, dim1, off2, dim2, offo, mulo, dimo, woff
constexpr int kCgFloats = 125;   // [5][5][5] coupling (net coefficient folded), row-major (i, j, k)

// out[row][c] = bias[c] + sum_{instructions writing c} sum_u W[u][w] * sum_{i,j} C[i][j][k] x1[row][u, i] x2[row][j]
__global__ void generic_tp_kernel(const float* __restrict__ x1, int d1, const float* __restrict__ x2, int d2,
                                  long long rows, const float* __restrict__ weights, const int* __restrict__ instr,
                                  int n_instr, const float* __restrict__ cg, const float* __restrict__ bias, int dout,
                                  float* __restrict__ out) {
  const long long total = rows * dout;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / dout;
    const int c = (int)(idx - row * dout);
    const float* a = x1 + row * d1;
    const float* b = x2 + row * d2;
    float acc = bias != nullptr ? bias[c] : 0.f;
    for (int q = 0; q < n_instr; ++q) {
      const int* in = instr + q * kInstrInts;
      const int offo = in[5], mulo = in[6], dimo = in[7];
      if (c < offo || c >= offo + mulo * dimo) continue;
      const int w = (c - offo) / dimo, k = (c - offo) - w * dimo;
      const int off1 = in[0], mul1 = in[1], dim1 = in[2], off2 = in[3], dim2 = in[4], woff = in[8];
      const float* C = cg + q * kCgFloats;
      float m[5];
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        float s = 0.f;
        if (i < dim1)
          for (int j = 0; j < dim2; ++j) s = fmaf(C[(i * 5 + j) * 5 + k], b[off2 + j], s);
        m[i] = s;
      }
      const float* W = weights + woff + w;
      const float* xa = a + off1;
      for (int u = 0; u < mul1; ++u) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 5; ++i)
          if (i < dim1) t = fmaf(m[i], xa[u * dim1 + i], t);
        acc = fmaf(W[(long long)u * mulo], t, acc);
      }
    }
    out[idx] = acc;
  }
}

// Row-tiled version of the same contraction: a block stages R rows of x1 / x2 in shared memory and every thread keeps R
// accumulators per output column, so a weight element is fetched once per R rows (the one-thread-per-output kernel
// above re-reads the whole weight column for every row: 1.3% of the FFMA peak on BASELINE config 3) and the inputs are
// warp-broadcast shared-memory reads.  Same summation order per output as the kernel above.
template <int R>
__global__ void __launch_bounds__(256)
    generic_tp_tiled_kernel(const float* __restrict__ x1, int d1, const float* __restrict__ x2, int d2, long long rows,
                            const float* __restrict__ weights, const int* __restrict__ instr, int n_instr,
                            const float* __restrict__ cg, const float* __restrict__ bias, int dout,
                            float* __restrict__ out) {
  extern __shared__ float tile[];
  float* x1s = tile;            // [R][d1]
  float* x2s = tile + R * d1;   // [R][d2]
  for (long long row0 = (long long)blockIdx.x * R; row0 < rows; row0 += (long long)gridDim.x * R) {
    const int live = (int)(rows - row0 < R ? rows - row0 : R);
    __syncthreads();  // previous tile fully consumed
    for (int i = threadIdx.x; i < R * d1; i += blockDim.x) {
      const int r = i / d1;
      x1s[i] = r < live ? x1[row0 * d1 + i] : 0.f;
    }
    for (int i = threadIdx.x; i < R * d2; i += blockDim.x) {
      const int r = i / d2;
      x2s[i] = r < live ? x2[row0 * d2 + i] : 0.f;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < dout; c += blockDim.x) {
      float acc[R];
      const float b0 = bias != nullptr ? bias[c] : 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = b0;
      for (int q = 0; q < n_instr; ++q) {
        const int* in = instr + q * kInstrInts;
        const int offo = in[5], mulo = in[6], dimo = in[7];
        if (c < offo || c >= offo + mulo * dimo) continue;
        const int w = (c - offo) / dimo, k = (c - offo) - w * dimo;
        const int off1 = in[0], mul1 = in[1], dim1 = in[2], off2 = in[3], dim2 = in[4], woff = in[8];
        const float* C = cg + q * kCgFloats;
        float m[R][5];  // coupling contracted with the second operand, per row
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            float sm = 0.f;
            if (i < dim1)
              for (int j = 0; j < dim2; ++j) sm = fmaf(C[(i * 5 + j) * 5 + k], x2s[r * d2 + off2 + j], sm);
            m[r][i] = sm;
          }
        const float* W = weights + woff + w;
        const float* xa = x1s + off1;
        if (dim1 == 1) {
          for (int u = 0; u < mul1; ++u) {
            const float wv = W[(long long)u * mulo];
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = fmaf(wv, m[r][0] * xa[r * d1 + u], acc[r]);
          }
        } else if (dim1 == 3) {
          for (int u = 0; u < mul1; ++u) {
            const float wv = W[(long long)u * mulo];
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const float* xr = xa + r * d1 + u * 3;
              float t = m[r][0] * xr[0];
              t = fmaf(m[r][1], xr[1], t);
              t = fmaf(m[r][2], xr[2], t);
              acc[r] = fmaf(wv, t, acc[r]);
            }
          }
        } else {
          for (int u = 0; u < mul1; ++u) {
            const float wv = W[(long long)u * mulo];
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const float* xr = xa + r * d1 + u * dim1;
              float t = 0.f;
#pragma unroll
              for (int i = 0; i < 5; ++i)
                if (i < dim1) t = fmaf(m[r][i], xr[i], t);
              acc[r] = fmaf(wv, t, acc[r]);
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (r < live) out[(row0 + r) * dout + c] = acc[r];
    }
  }
}

// ---- "expand + GEMM" form for large row counts (per-edge tensor products) -----------------------------------------
// For one output irrep block (mulo x dimo) the contraction is a plain GEMM once the coupling with the second operand
// has been applied:  out[row][w][k] = sum_{(p,u)} Wcat[(p,u)][w] * A[row * dimo + k][(p,u)],
//   A[row * dimo + k][koff_p + u] = sum_{i,j} C_p[i][j][k] x1[row][off1_p + u * dim1_p + i] x2[row][off2_p + j]
// over the paths p that write the block.  generic_tp_expand_kernel builds A, a library SGEMM (plain GEMM, fp32) does
// the weight contraction, generic_tp_scatter_kernel puts the result back into the e3nn column order (+ bias).
constexpr int kPathInts = 6;  // off1, mul1, dim1, off2, dim2, koff

__global__ void generic_tp_expand_kernel(const float* __restrict__ x1, int d1, const float* __restrict__ x2, int d2,
                                         long long rows, const int* __restrict__ paths, int n_paths,
                                         const float* __restrict__ cg, int dimo, int K, long long lda,
                                         float* __restrict__ A) {
  const long long total = rows * K;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / K;
    const int kc = (int)(idx - row * K);
    int p = 0;
    while (p + 1 < n_paths && kc >= paths[(p + 1) * kPathInts + 5]) ++p;
    const int* pa = paths + p * kPathInts;
    const int off1 = pa[0], dim1 = pa[2], off2 = pa[3], dim2 = pa[4], u = kc - pa[5];
    const float* C = cg + p * kCgFloats;
    const float* xa = x1 + row * d1 + off1 + u * dim1;
    const float* b = x2 + row * d2 + off2;
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < dim1; ++i) {
      const float xv = xa[i];
      for (int j = 0; j < dim2; ++j) {
        const float xb = xv * b[j];
#pragma unroll
        for (int k = 0; k < 5; ++k)
          if (k < dimo) acc[k] = fmaf(C[(i * 5 + j) * 5 + k], xb, acc[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 5; ++k)
      if (k < dimo) A[(row * dimo + k) * lda + kc] = acc[k];
  }
}

// out[row][offo + w * dimo + k] = Y[row * dimo + k][w] + bias[offo + w * dimo + k]
__global__ void generic_tp_scatter_kernel(const float* __restrict__ Y, long long ldy, long long rows, int dimo, int mulo,
                                          int offo, int dout, const float* __restrict__ bias, float* __restrict__ out) {
  const int blk = mulo * dimo;
  const long long total = rows * blk;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / blk;
    const int c = (int)(idx - row * blk);
    const int w = c / dimo, k = c - w * dimo;
    float v = Y[(row * dimo + k) * ldy + w];
    if (bias != nullptr) v += bias[offo + c];
    out[row * dout + offo + c] = v;
  }
}

// ---- message_layer_1 hoisted to node level (any irreps) ---------------------------------------------------------------
// The tensor product is linear in its first operand cat(x_i, x_j, add), so the weight contraction of the x_i / x_j
// parts is done per NODE (Y = segnn_generic_tp with an identity coupling: Y[node][yoff + w * dim1 + i] =
// sum_u W[u][w] x[node][u, i]) and per edge only the coupling with the edge attribute remains:
//   out[e][offo + w * dimo + k] = bias + sum_pairs sum_{i,j} C[i][j][k] (Yi[b][yi + w dim1 + i] + Yj[a][yj + w dim1 + i]) attr[e][off2 + j]
//                                + sum_add sum_u W[u][w] add[e][u] sum_j C[0][j][k] attr[e][off2 + j]
// for edge e = (source a -> target b) in the reference order.  ~15 fused multiply-adds per output instead of ~350.
constexpr int kPairInts = 8;  // offo, mulo, dimo, dim1, off2, dim2, yoff_i, yoff_j  (coupling cg[pair])
constexpr int kAddInts = 7;   // offo, mulo, dimo, off2, dim2, woff, mul1              (coupling cg[n_pairs + index])

// One block per edge: the coupling of every pair with the edge attribute (M[q][i][k] = sum_j C[i][j][k] attr[j]) is
// computed once per edge into shared memory; every thread then owns (output block, channel w) items and produces all
// dimo components of them (one thread per output redid the coupling and walked the whole pair table for each).
constexpr int kMaxPairs = 12, kMaxAdds = 4;

__global__ void __launch_bounds__(128)
    generic_hoisted_msg1_kernel(const float* __restrict__ Y, int ydim, const float* __restrict__ attr, int d2,
                                const float* __restrict__ add, int d_add, int B, int N, const int* __restrict__ pairs,
                                int n_pairs, const int* __restrict__ adds, int n_adds, const int* __restrict__ blocks,
                                int n_blocks, int n_items, const float* __restrict__ cg,
                                const float* __restrict__ weights, const float* __restrict__ bias, int dout,
                                float* __restrict__ out) {
  __shared__ float M[kMaxPairs][25];
  __shared__ float Ma[kMaxAdds][5];
  __shared__ float av[8];
  const long long E = (long long)B * N * (N - 1);
  for (long long e = blockIdx.x; e < E; e += gridDim.x) {
    const long long g = e / ((long long)N * (N - 1));
    const int le = (int)(e - g * N * (N - 1));
    const int a = le / (N - 1), bb = le - a * (N - 1);
    const int b = bb < a ? bb : bb + 1;
    const float* yi = Y + (g * N + b) * ydim;  // receiver (x_i) side
    const float* yj = Y + (g * N + a) * ydim;  // sender (x_j) side
    const float* at = attr + e * d2;
    __syncthreads();  // previous edge fully consumed
    for (int t = threadIdx.x; t < n_pairs * 25; t += blockDim.x) {
      const int q = t / 25, i = (t - q * 25) / 5, k = t % 5;
      const int* in = pairs + q * kPairInts;
      float m = 0.f;
      if (i < in[3] && k < in[2])
        for (int j = 0; j < in[5]; ++j) m = fmaf(cg[q * kCgFloats + (i * 5 + j) * 5 + k], at[in[4] + j], m);
      M[q][i * 5 + k] = m;
    }
    for (int t = threadIdx.x; t < n_adds * 5; t += blockDim.x) {
      const int q = t / 5, k = t % 5;
      const int* in = adds + q * kAddInts;
      float m = 0.f;
      if (k < in[2])
        for (int j = 0; j < in[4]; ++j) m = fmaf(cg[(n_pairs + q) * kCgFloats + j * 5 + k], at[in[3] + j], m);
      Ma[q][k] = m;
    }
    if (threadIdx.x < d_add && threadIdx.x < 8) av[threadIdx.x] = add[e * d_add + threadIdx.x];
    __syncthreads();
    for (int item = threadIdx.x; item < n_items; item += blockDim.x) {
      int blk = 0, w = item;
      while (w >= blocks[blk * 3 + 1]) {
        w -= blocks[blk * 3 + 1];
        ++blk;
      }
      const int offo = blocks[blk * 3], mulo = blocks[blk * 3 + 1], dimo = blocks[blk * 3 + 2];
      float acc[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) acc[k] = (bias != nullptr && k < dimo) ? bias[offo + w * dimo + k] : 0.f;
      for (int q = 0; q < n_pairs; ++q) {
        const int* in = pairs + q * kPairInts;
        if (in[0] != offo) continue;
        const int dim1 = in[3];
        const float* pi = yi + in[6] + w * dim1;
        const float* pj = yj + in[7] + w * dim1;
        for (int i = 0; i < dim1; ++i) {
          const float sv = pi[i] + pj[i];
#pragma unroll
          for (int k = 0; k < 5; ++k) acc[k] = fmaf(M[q][i * 5 + k], sv, acc[k]);
        }
      }
      for (int q = 0; q < n_adds; ++q) {
        const int* in = adds + q * kAddInts;
        if (in[0] != offo) continue;
        float t = 0.f;
        for (int u = 0; u < in[6]; ++u) t = fmaf(weights[in[5] + (long long)u * mulo + w], av[u], t);
#pragma unroll
        for (int k = 0; k < 5; ++k) acc[k] = fmaf(Ma[q][k], t, acc[k]);
      }
      float* o = out + e * dout + offo + w * dimo;
#pragma unroll
      for (int k = 0; k < 5; ++k)
        if (k < dimo) o[k] = acc[k];
    }
  }
}

// e3nn Gate: x = [n_s scalars | n_g gates | gated]; out = [c_silu silu(scalars) | gated * c_sig sigmoid(gate)]
__global__ void generic_gate_kernel(const float* __restrict__ x, long long rows, int n_s, int n_g, int d_gated,
                                    const int* __restrict__ gate_index, float* __restrict__ out) {
  const int din = n_s + n_g + d_gated, dout = n_s + d_gated;
  const long long total = rows * dout;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / dout;
    const int c = (int)(idx - row * dout);
    const float* r = x + row * din;
    out[idx] = c < n_s ? silu_gate(r[c]) : r[n_s + n_g + (c - n_s)] * sig_gate(r[n_s + gate_index[c - n_s]]);
  }
}

// message input cat(x_i, x_j, add) in the reference edge order (graph-major, source ascending, target ascending):
// edge (g, a -> b): x_i = x[target b], x_j = x[source a]   (models/segnn/segnn.py:264-277)
__global__ void generic_message_input_kernel(const float* __restrict__ x, const float* __restrict__ add, int B, int N,
                                             int D, int d_add, float* __restrict__ out) {
  const int dout = 2 * D + d_add;
  const long long E = (long long)B * N * (N - 1);
  const long long total = E * dout;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long e = idx / dout;
    const int c = (int)(idx - e * dout);
    const long long g = e / ((long long)N * (N - 1));
    const int le = (int)(e - g * N * (N - 1));
    const int a = le / (N - 1), bb = le - a * (N - 1);
    const int b = bb < a ? bb : bb + 1;
    float v;
    if (c < D) v = x[(g * N + b) * D + c];
    else if (c < 2 * D) v = x[(g * N + a) * D + (c - D)];
    else v = add[e * d_add + (c - 2 * D)];
    out[idx] = v;
  }
}

// agg[target] = sum over sources (ascending) of m[edge(source -> target)]: deterministic scatter-sum (segnn.py:205)
__global__ void generic_aggregate_kernel(const float* __restrict__ m, int B, int N, int D, float* __restrict__ agg) {
  const long long total = (long long)B * N * D;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long node = idx / D;
    const int c = (int)(idx - node * D);
    const long long g = node / N;
    const int b = (int)(node - g * N);
    float s = 0.f;
    for (int a = 0; a < N; ++a) {
      if (a == b) continue;
      const long long e = g * N * (N - 1) + (long long)a * (N - 1) + (b < a ? b : b - 1);
      s += m[e * D + c];
    }
    agg[idx] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// O3Transform for lmax_attr <= 2 (o3_building_blocks.py:230-278): e3nn 'integral' harmonics of the unit vector, degrees
// 0..lmax in e3nn's component order; l = 2: sqrt(15) xz, sqrt(15) xy, sqrt(5) (y^2 - (x^2 + z^2) / 2), sqrt(15) yz,
// sqrt(15) / 2 (z^2 - x^2), all over sqrt(4 pi).  The fused kernels keep lmax_attr = 1 in registers; these feed the
// table-driven kernels above.
// ------------------------------------------------------------------------------------------------
constexpr float kY2a = 1.0925484305920792f;  // sqrt(15 / (4 pi))
constexpr float kY2b = 0.6307831305050401f;  // sqrt(5 / (4 pi))

__device__ __forceinline__ void harmonics_l2(float x, float y, float z, float* o) {
  o[0] = kY2a * x * z;
  o[1] = kY2a * x * y;
  o[2] = kY2b * (y * y - 0.5f * (x * x + z * z));
  o[3] = kY2a * y * z;
  o[4] = 0.5f * kY2a * (z * z - x * x);
}

__global__ void edge_attr_lmax_kernel(const float* __restrict__ pos, const float* __restrict__ mass, int B, int N,
                                      int lmax, float* __restrict__ edge_attr, float* __restrict__ add) {
  const int d = (lmax + 1) * (lmax + 1);
  const long long per = (long long)N * (N - 1);
  const long long E = per * B;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x) {
    const long long g = e / per;
    const long long r = e - g * per;
    const int a = (int)(r / (N - 1));
    int b = (int)(r - (long long)a * (N - 1));
    b += (b >= a);
    const long long s = g * N + a, t = g * N + b;  // sender (source), receiver (target)
    float ux, uy, uz, len;
    unit_vec(pos[s * 3 + 0] - pos[t * 3 + 0], pos[s * 3 + 1] - pos[t * 3 + 1], pos[s * 3 + 2] - pos[t * 3 + 2], ux,
             uy, uz, len);
    float* o = edge_attr + e * d;
    o[0] = kY0;
    if (lmax >= 1) {
      o[1] = kY1 * ux;
      o[2] = kY1 * uy;
      o[3] = kY1 * uz;
    }
    if (lmax >= 2) harmonics_l2(ux, uy, uz, o + 4);
    add[e * 2 + 0] = len;
    add[e * 2 + 1] = mass[s] * mass[t];
  }
}

// node_attr = mean over the senders j of Y(r_j - r_i) + Y(v_i), l = 0 slot set to 1 (segnn.py:148); x as in K1.
// One thread per receiver, senders read through L1 / L2 (this path is the checker-grade generic one, not tuned).
__global__ void prep_lmax_kernel(const float* __restrict__ pos, const float* __restrict__ vel, int B, int N, int lmax,
                                 float* __restrict__ x_in, float* __restrict__ node_attr) {
  const int d = (lmax + 1) * (lmax + 1);
  const long long nodes = (long long)B * N;
  for (long long node = blockIdx.x * (long long)blockDim.x + threadIdx.x; node < nodes;
       node += (long long)gridDim.x * blockDim.x) {
    const long long base = (node / N) * N;
    const int i = (int)(node - base);
    const float px = pos[node * 3 + 0], py = pos[node * 3 + 1], pz = pos[node * 3 + 2];
    float s1[3] = {0.f, 0.f, 0.f}, s2[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < N; ++j) {
      if (j == i) continue;
      float ux, uy, uz, len, h[5];
      unit_vec(pos[(base + j) * 3 + 0] - px, pos[(base + j) * 3 + 1] - py, pos[(base + j) * 3 + 2] - pz, ux, uy, uz,
               len);
      s1[0] += ux;
      s1[1] += uy;
      s1[2] += uz;
      if (lmax >= 2) {
        harmonics_l2(ux, uy, uz, h);
#pragma unroll
        for (int k = 0; k < 5; ++k) s2[k] += h[k];
      }
    }
    const float vx = vel[node * 3 + 0], vy = vel[node * 3 + 1], vz = vel[node * 3 + 2];
    float ux, uy, uz, vlen;
    unit_vec(vx, vy, vz, ux, uy, uz, vlen);
    const float inv_deg = N > 1 ? 1.0f / (float)(N - 1) : 0.0f;
    float* o = node_attr + node * d;
    o[0] = 1.0f;
    if (lmax >= 1) {
      o[1] = kY1 * (s1[0] * inv_deg) + kY1 * ux;
      o[2] = kY1 * (s1[1] * inv_deg) + kY1 * uy;
      o[3] = kY1 * (s1[2] * inv_deg) + kY1 * uz;
    }
    if (lmax >= 2) {
      float h[5];
      harmonics_l2(ux, uy, uz, h);
#pragma unroll
      for (int k = 0; k < 5; ++k) o[4 + k] = s2[k] * inv_deg + h[k];
    }
    const float m = (px + py + pz) / 3.0f;  // reference quirk: mean over xyz of the node (o3_building_blocks.py:274)
    float* x = x_in + node * 7;
    x[0] = px - m;
    x[1] = py - m;
    x[2] = pz - m;
    x[3] = vx;
    x[4] = vy;
    x[5] = vz;
    x[6] = vlen;
  }
}

// ------------------------------------------------------------------------------------------------
// Explicit edge lists (kNN graphs, utils/build_fully_connected_graph.py:42-80; num_neighbors < N - 1): the same
// reference formulation with gathers through edge_index (row 0 = source j, row 1 = target i) and a deterministic
// segment reduction over the edges of a target (CSR built by the host-side plumbing from a stable sort of the targets).
// ------------------------------------------------------------------------------------------------
__global__ void edge_attr_list_kernel(const float* __restrict__ pos, const float* __restrict__ mass,
                                      const long long* __restrict__ ei, long long E, int lmax,
                                      float* __restrict__ edge_attr, float* __restrict__ add) {
  const int d = (lmax + 1) * (lmax + 1);
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x) {
    const long long s = ei[e], t = ei[E + e];
    float ux, uy, uz, len;
    unit_vec(pos[s * 3 + 0] - pos[t * 3 + 0], pos[s * 3 + 1] - pos[t * 3 + 1], pos[s * 3 + 2] - pos[t * 3 + 2], ux,
             uy, uz, len);
    float* o = edge_attr + e * d;
    o[0] = kY0;
    if (lmax >= 1) {
      o[1] = kY1 * ux;
      o[2] = kY1 * uy;
      o[3] = kY1 * uz;
    }
    if (lmax >= 2) harmonics_l2(ux, uy, uz, o + 4);
    add[e * 2 + 0] = len;
    add[e * 2 + 1] = mass[s] * mass[t];
  }
}

// out[e] = [x[target] | x[source] | add[e]]  (segnn.py:264-279: cat(x_i, x_j, additional_message_features))
__global__ void message_input_list_kernel(const float* __restrict__ x, const float* __restrict__ add,
                                          const long long* __restrict__ ei, long long E, int D, int d_add,
                                          float* __restrict__ out) {
  const int dout = 2 * D + d_add;
  const long long total = E * dout;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long e = idx / dout;
    const int c = (int)(idx - e * dout);
    float v;
    if (c < D) v = x[ei[E + e] * D + c];
    else if (c < 2 * D) v = x[ei[e] * D + (c - D)];
    else v = add[e * d_add + (c - 2 * D)];
    out[idx] = v;
  }
}

// out[node][c] = sum (or mean; 0 for an empty segment) over k in [ptr[node], ptr[node + 1]) of values[order[k]][c], in
// that order: the scatter of segnn.py:205 / o3_building_blocks.py:257-263 without atomics, run-to-run identical
__global__ void segment_reduce_kernel(const float* __restrict__ values, const long long* __restrict__ order,
                                      const long long* __restrict__ ptr, long long nodes, int D, int mean,
                                      float* __restrict__ out) {
  const long long total = nodes * D;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long node = idx / D;
    const int c = (int)(idx - node * D);
    const long long k0 = ptr[node], k1 = ptr[node + 1];
    float s = 0.f;
    for (long long k = k0; k < k1; ++k) s += values[order[k] * D + c];
    if (mean && k1 > k0) s /= (float)(k1 - k0);
    out[idx] = s;
  }
}

// node_attr = mean_attr (scatter-mean of the incoming edge attributes) + Y(v), l = 0 slot set to 1; x as in K1
__global__ void prep_list_kernel(const float* __restrict__ pos, const float* __restrict__ vel,
                                 const float* __restrict__ mean_attr, long long nodes, int lmax,
                                 float* __restrict__ x_in, float* __restrict__ node_attr) {
  const int d = (lmax + 1) * (lmax + 1);
  for (long long node = blockIdx.x * (long long)blockDim.x + threadIdx.x; node < nodes;
       node += (long long)gridDim.x * blockDim.x) {
    const float px = pos[node * 3 + 0], py = pos[node * 3 + 1], pz = pos[node * 3 + 2];
    const float vx = vel[node * 3 + 0], vy = vel[node * 3 + 1], vz = vel[node * 3 + 2];
    float ux, uy, uz, vlen;
    unit_vec(vx, vy, vz, ux, uy, uz, vlen);
    const float* ma = mean_attr + node * d;
    float* o = node_attr + node * d;
    o[0] = 1.0f;
    if (lmax >= 1) {
      o[1] = ma[1] + kY1 * ux;
      o[2] = ma[2] + kY1 * uy;
      o[3] = ma[3] + kY1 * uz;
    }
    if (lmax >= 2) {
      float h[5];
      harmonics_l2(ux, uy, uz, h);
#pragma unroll
      for (int k = 0; k < 5; ++k) o[4 + k] = ma[4 + k] + h[k];
    }
    const float m = (px + py + pz) / 3.0f;  // o3_building_blocks.py:274
    float* x = x_in + node * 7;
    x[0] = px - m;
    x[1] = py - m;
    x[2] = pz - m;
    x[3] = vx;
    x[4] = vy;
    x[5] = vz;
    x[6] = vlen;
  }
}

static inline int generic_grid(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = 148LL * 64;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace segnn

using namespace segnn;

extern "C" {

int segnn_generic_tp(const float* x1, int d1, const float* x2, int d2, int64_t rows, const float* weights,
                     const int* instr, int n_instr, const float* cg, const float* bias, int dout, float* out,
                     segnn_stream_t stream) {
  SEGNN_CHECK_ARG(rows >= 0 && d1 >= 1 && d2 >= 1 && dout >= 1 && n_instr >= 1, "bad sizes");
  if (rows == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(x1 && x2 && weights && instr && cg && out, "null pointer");
  // rows per block: 8 when there are enough rows to fill the GPU with 8-row tiles, else 2 (node-level products on a
  // few hundred rows would otherwise run on a third of the SMs)
  const int R = rows >= 8LL * 2 * 148 ? 8 : 2;
  const size_t smem = (size_t)R * (d1 + d2) * sizeof(float);
  if (rows >= 4 * R && smem <= 200 * 1024) {
    auto kern = R == 8 ? generic_tp_tiled_kernel<8> : generic_tp_tiled_kernel<2>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) {
      set_error("segnn_generic_tp: cudaFuncSetAttribute: %s", cudaGetErrorString(err));
      return SEGNN_E_CUDA;
    }
    long long blocks = (rows + R - 1) / R;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    kern<<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(x1, d1, x2, d2, rows, weights, instr, n_instr, cg, bias,
                                                              dout, out);
  } else {
    generic_tp_kernel<<<generic_grid(rows * dout), 256, 0, (cudaStream_t)stream>>>(x1, d1, x2, d2, rows, weights,
                                                                                 instr, n_instr, cg, bias, dout, out);
  }
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_generic_tp_expand_ld(const float* x1, int d1, const float* x2, int d2, int64_t rows, const int* paths,
                               int n_paths, const float* cg, int dimo, int K, int64_t lda, float* A,
                               segnn_stream_t stream) {
  SEGNN_CHECK_ARG(rows >= 0 && d1 >= 1 && d2 >= 1 && n_paths >= 1 && dimo >= 1 && dimo <= 5 && K >= 1 && lda >= K,
                  "bad sizes");
  if (rows == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(x1 && x2 && paths && cg && A, "null pointer");
  generic_tp_expand_kernel<<<generic_grid(rows * K), 256, 0, (cudaStream_t)stream>>>(x1, d1, x2, d2, rows, paths,
                                                                                   n_paths, cg, dimo, K, lda, A);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_generic_tp_expand(const float* x1, int d1, const float* x2, int d2, int64_t rows, const int* paths,
                            int n_paths, const float* cg, int dimo, int K, float* A, segnn_stream_t stream) {
  return segnn_generic_tp_expand_ld(x1, d1, x2, d2, rows, paths, n_paths, cg, dimo, K, K, A, stream);
}

int segnn_generic_tp_scatter_ld(const float* Y, int64_t ldy, int64_t rows, int dimo, int mulo, int offo, int dout,
                                const float* bias, float* out, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(rows >= 0 && dimo >= 1 && mulo >= 1 && offo >= 0 && dout >= offo + mulo * dimo && ldy >= mulo,
                  "bad sizes");
  if (rows == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(Y && out, "null pointer");
  generic_tp_scatter_kernel<<<generic_grid(rows * mulo * dimo), 256, 0, (cudaStream_t)stream>>>(
      Y, ldy, rows, dimo, mulo, offo, dout, bias, out);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_generic_tp_scatter(const float* Y, int64_t rows, int dimo, int mulo, int offo, int dout, const float* bias,
                             float* out, segnn_stream_t stream) {
  return segnn_generic_tp_scatter_ld(Y, mulo, rows, dimo, mulo, offo, dout, bias, out, stream);
}

int segnn_generic_hoisted_msg1(const float* Y, int ydim, const float* attr, int d2, const float* add, int d_add, int B,
                               int N, const int* pairs, int n_pairs, const int* adds, int n_adds, const int* blocks,
                               int n_blocks, int n_items, const float* cg, const float* weights, const float* bias,
                               int dout, float* out, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(B >= 0 && N >= 2 && ydim >= 1 && d2 >= 1 && dout >= 1 && n_pairs >= 0 && n_adds >= 0, "bad sizes");
  SEGNN_CHECK_ARG(n_pairs <= kMaxPairs && n_adds <= kMaxAdds && d_add <= 8 && n_blocks >= 1 && n_items >= 1,
                  "too many instruction pairs / additional features for the per-edge tables");
  if (B == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(Y && attr && cg && out && blocks && (n_pairs == 0 || pairs) &&
                      (n_adds == 0 || (adds && add && weights)),
                  "null pointer");
  long long nblk = (long long)B * N * (N - 1);
  if (nblk > 148LL * 256) nblk = 148LL * 256;
  generic_hoisted_msg1_kernel<<<(unsigned)nblk, 128, 0, (cudaStream_t)stream>>>(
      Y, ydim, attr, d2, add, d_add, B, N, pairs, n_pairs, adds, n_adds, blocks, n_blocks, n_items, cg, weights, bias,
      dout, out);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_generic_gate(const float* x, int64_t rows, int n_scalars, int n_gates, int d_gated, const int* gate_index,
                       float* out, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(rows >= 0 && n_scalars >= 0 && n_gates >= 0 && d_gated >= 0, "bad sizes");
  if (rows == 0 || n_scalars + d_gated == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(x && out && (d_gated == 0 || gate_index), "null pointer");
  generic_gate_kernel<<<generic_grid(rows * (n_scalars + d_gated)), 256, 0, (cudaStream_t)stream>>>(
      x, rows, n_scalars, n_gates, d_gated, gate_index, out);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_generic_message_input(const float* x, const float* add, int B, int N, int D, int d_add, float* out,
                                segnn_stream_t stream) {
  SEGNN_CHECK_ARG(B >= 0 && N >= 2 && D >= 1 && d_add >= 0, "bad sizes");
  if (B == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(x && out && (d_add == 0 || add), "null pointer");
  generic_message_input_kernel<<<generic_grid((long long)B * N * (N - 1) * (2 * D + d_add)), 256, 0,
                                 (cudaStream_t)stream>>>(x, add, B, N, D, d_add, out);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_generic_aggregate(const float* m, int B, int N, int D, float* agg, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(B >= 0 && N >= 2 && D >= 1, "bad sizes");
  if (B == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(m && agg, "null pointer");
  generic_aggregate_kernel<<<generic_grid((long long)B * N * D), 256, 0, (cudaStream_t)stream>>>(m, B, N, D, agg);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_edge_attr_lmax(const float* pos, const float* mass, int B, int N, int lmax_attr, float* edge_attr, float* add,
                         segnn_stream_t stream) {
  SEGNN_CHECK_ARG(B >= 0 && N >= 2 && lmax_attr >= 0 && lmax_attr <= 2, "bad sizes (lmax_attr in 0..2)");
  if (B == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(pos && mass && edge_attr && add, "null pointer");
  edge_attr_lmax_kernel<<<generic_grid((long long)B * N * (N - 1)), 256, 0, (cudaStream_t)stream>>>(
      pos, mass, B, N, lmax_attr, edge_attr, add);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_prep_fwd_lmax(const float* pos, const float* vel, int B, int N, int lmax_attr, float* x_in, float* node_attr,
                        segnn_stream_t stream) {
  SEGNN_CHECK_ARG(B >= 0 && N >= 1 && lmax_attr >= 0 && lmax_attr <= 2, "bad sizes (lmax_attr in 0..2)");
  if (B == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(pos && vel && x_in && node_attr, "null pointer");
  prep_lmax_kernel<<<generic_grid((long long)B * N), 256, 0, (cudaStream_t)stream>>>(pos, vel, B, N, lmax_attr, x_in,
                                                                                    node_attr);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_edge_attr_list(const float* pos, const float* mass, const int64_t* edge_index, int64_t E, int lmax_attr,
                         float* edge_attr, float* add, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(E >= 0 && lmax_attr >= 0 && lmax_attr <= 2, "bad sizes (lmax_attr in 0..2)");
  if (E == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(pos && mass && edge_index && edge_attr && add, "null pointer");
  edge_attr_list_kernel<<<generic_grid(E), 256, 0, (cudaStream_t)stream>>>(
      pos, mass, reinterpret_cast<const long long*>(edge_index), E, lmax_attr, edge_attr, add);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_generic_message_input_list(const float* x, const float* add, const int64_t* edge_index, int64_t E, int D,
                                     int d_add, float* out, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(E >= 0 && D >= 1 && d_add >= 0, "bad sizes");
  if (E == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(x && edge_index && out && (d_add == 0 || add), "null pointer");
  message_input_list_kernel<<<generic_grid(E * (2 * D + d_add)), 256, 0, (cudaStream_t)stream>>>(
      x, add, reinterpret_cast<const long long*>(edge_index), E, D, d_add, out);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_segment_reduce(const float* values, const int64_t* order, const int64_t* ptr, int64_t nodes, int D, int mean,
                         float* out, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(nodes >= 0 && D >= 1, "bad sizes");
  if (nodes == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(values && order && ptr && out, "null pointer");
  segment_reduce_kernel<<<generic_grid(nodes * D), 256, 0, (cudaStream_t)stream>>>(
      values, reinterpret_cast<const long long*>(order), reinterpret_cast<const long long*>(ptr), nodes, D, mean, out);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_prep_fwd_list(const float* pos, const float* vel, const float* mean_attr, int64_t nodes, int lmax_attr,
                        float* x_in, float* node_attr, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(nodes >= 0 && lmax_attr >= 0 && lmax_attr <= 2, "bad sizes (lmax_attr in 0..2)");
  if (nodes == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(pos && vel && mean_attr && x_in && node_attr, "null pointer");
  prep_list_kernel<<<generic_grid(nodes), 256, 0, (cudaStream_t)stream>>>(pos, vel, mean_attr, nodes, lmax_attr, x_in,
                                                                          node_attr);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

}  // extern "C"
