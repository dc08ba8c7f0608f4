// Training-side node-level kernels (fp32): deterministic column reductions (BatchNorm statistics, bias and weight
// gradients), per-column affine / linear combination (train-mode BatchNorm forward and backward), backward of the
// attribute combine + gate, the weight-gradient GEMM, and the backward of the embedding and the head.
// Reference semantics: e3nn BatchNorm as used at models/segnn/segnn.py:233-235,257-261,282-283; autograd of
// o3_building_blocks.py:150-203; training/losses.py:22-45.
#include "segnn_common.cuh"

namespace segnn {

// ------------------------------------------------------------------------------------------------
// colsum: out[c] = sum_r f(x[r][c], y[r][c]);  mode 0: x, 1: x*x, 2: x*y.  Two deterministic stages.
// ------------------------------------------------------------------------------------------------
constexpr int kColsumRowsPerBlock = 64;  // rows per partial sum (fixed => the reduction order is deterministic)
constexpr int kColsumRowLanes = 8;       // threads that share one column inside a block

// block = 32 columns x 8 row lanes; lane ry sums rows r0 + ry, r0 + ry + 8, ...; the 8 lane sums are combined in
// shared memory in a fixed order.
__global__ void __launch_bounds__(256) colsum_stage1(const float* __restrict__ x, const float* __restrict__ y,
                                                   int64_t rows, int cols, int mode, double* __restrict__ partial) {
  __shared__ double red[kColsumRowLanes][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const int64_t r0 = (int64_t)blockIdx.y * kColsumRowsPerBlock;
  const int64_t r1 = min(rows, r0 + kColsumRowsPerBlock);
  // float64 accumulation: BatchNorm statistics feed var = E[x^2] - mean^2 and cancellation-prone gradient sums
  double acc = 0.0;
  if (c < cols) {
    for (int64_t r = r0 + ry; r < r1; r += kColsumRowLanes) {
      const double v = (double)x[r * cols + c];
      acc += mode == 0 ? v : (mode == 1 ? v * v : v * (double)y[r * cols + c]);
    }
  }
  red[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && c < cols) {
    double s = red[0][cx];
#pragma unroll
    for (int i = 1; i < kColsumRowLanes; ++i) s += red[i][cx];
    partial[(int64_t)blockIdx.y * cols + c] = s;
  }
}

// Few rows (training-size graphs): ONE launch.  Block = 32 columns x 32 row lanes; lane ry sums rows ry, ry + 32, ...
// in float64 and the 32 lane sums are added in lane order: deterministic (fixed association), one pass over the rows with
// all loads of a lane independent.  The training step of the README configuration makes ~190 column sums per step, each of
// which was two ~4 us launches on the critical chain.
__global__ void __launch_bounds__(1024) colsum_small(const float* __restrict__ x, const float* __restrict__ y,
                                                   int64_t rows, int cols, int mode, float* __restrict__ out) {
  __shared__ double red[32][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  double acc = 0.0;
  if (c < cols) {
#pragma unroll 4
    for (int64_t r = ry; r < rows; r += 32) {
      const double v = (double)x[r * cols + c];
      acc += mode == 0 ? v : (mode == 1 ? v * v : v * (double)y[r * cols + c]);
    }
  }
  red[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && c < cols) {
    double s = red[0][cx];
#pragma unroll
    for (int i = 1; i < 32; ++i) s += red[i][cx];
    out[c] = (float)s;
  }
}
constexpr int kColsumSmallParts = 32;  // up to 2048 rows take the single-launch path

// Two independent column sums in one launch (blockIdx.y selects the job): the statistics of a train-mode BatchNorm always
// come in pairs (sum and sum of squares, sum g and sum g * x), each pair was two launches on the critical chain.
struct ColsumJob {
  const float* x;
  const float* y;
  int64_t rows;
  int cols, mode;
  float* out;
};
__global__ void __launch_bounds__(1024) colsum_small2(ColsumJob a, ColsumJob b) {
  __shared__ double red[32][33];
  const ColsumJob j = blockIdx.y == 0 ? a : b;
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  double acc = 0.0;
  if (c < j.cols) {  // same loop, same association as colsum_small: bit-identical sums
#pragma unroll 4
    for (int64_t r = ry; r < j.rows; r += 32) {
      const double v = (double)j.x[r * j.cols + c];
      acc += j.mode == 0 ? v : (j.mode == 1 ? v * v : v * (double)j.y[r * j.cols + c]);
    }
  }
  red[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && c < j.cols) {
    double s = red[0][cx];
#pragma unroll
    for (int i = 1; i < 32; ++i) s += red[i][cx];
    j.out[c] = (float)s;
  }
}

template <typename T>
__global__ void colsum_stage2(const T* __restrict__ partial, int nparts, int cols, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  T acc = 0;
  for (int p = 0; p < nparts; ++p) acc += partial[(int64_t)p * cols + c];
  out[c] = (float)acc;
}

// Many partial rows (102,400-node batches give 1,600): 32 columns x 32 part lanes per block; lane ry adds parts ry,
// ry + 32, ... and the 32 lane sums are combined in lane order -- the same fixed association on every run.  One thread
// per column walking all the parts took 165 us per column sum of a train-mode BatchNorm rollout step (24 per step).
template <typename T>
__global__ void __launch_bounds__(1024) colsum_stage2_wide(const T* __restrict__ partial, int nparts, int cols,
                                                         float* __restrict__ out) {
  __shared__ T red[32][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  T acc = 0;
  if (c < cols) {
#pragma unroll 4
    for (int p = ry; p < nparts; p += 32) acc += partial[(int64_t)p * cols + c];
  }
  red[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && c < cols) {
    T s = red[0][cx];
#pragma unroll
    for (int i = 1; i < 32; ++i) s += red[i][cx];
    out[c] = (float)s;
  }
}

// out[r][c] = A[c] * dy[r][c] + B[c] * x[r][c] + C[c]   (x / B / C may be NULL)
__global__ void lincomb_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ A,
                               const float* __restrict__ B, const float* __restrict__ C, int64_t rows, int cols,
                               float* __restrict__ out) {
  const int64_t total = rows * cols;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cols);
    float v = A[c] * dy[idx];
    if (x != nullptr) v = fmaf(B[c], x[idx], v);
    if (C != nullptr) v += C[c];
    out[idx] = v;
  }
}

// The same on float4 columns: a thread keeps its four column coefficients in registers and walks the rows (no
// per-element modulo, 128-bit accesses): 2.4 -> ~5 TB/s on the [102400, 384] passes of a train-mode BatchNorm rollout.
__global__ void __launch_bounds__(1024)
    lincomb4_kernel(const float4* __restrict__ dy, const float4* __restrict__ x, const float4* __restrict__ A,
                    const float4* __restrict__ B, const float4* __restrict__ C, int64_t rows, int cols4,
                    float4* __restrict__ out) {
  for (int c4 = threadIdx.x; c4 < cols4; c4 += blockDim.x) {
    const float4 a = A[c4];
    const float4 b = x != nullptr ? B[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 c = C != nullptr ? C[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int64_t r = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; r < rows; r += (int64_t)gridDim.x * blockDim.y) {
      const float4 d = dy[r * cols4 + c4];
      float4 v = make_float4(fmaf(a.x, d.x, c.x), fmaf(a.y, d.y, c.y), fmaf(a.z, d.z, c.z), fmaf(a.w, d.w, c.w));
      if (x != nullptr) {
        const float4 xv = x[r * cols4 + c4];
        v.x = fmaf(b.x, xv.x, v.x);
        v.y = fmaf(b.y, xv.y, v.y);
        v.z = fmaf(b.z, xv.z, v.z);
        v.w = fmaf(b.w, xv.w, v.w);
      }
      out[r * cols4 + c4] = v;
    }
  }
}

// d/dx of c*silu(x) and c*sigmoid(x)
__device__ __forceinline__ float silu_gate_grad(float x) {
  const float s = sigmoid_acc(x);
  return kCSilu * s * (1.0f + x * (1.0f - s));
}
__device__ __forceinline__ float sig_gate_grad(float x) {
  const float s = sigmoid_acc(x);
  return kCSig * s * (1.0f - s);
}

// ------------------------------------------------------------------------------------------------
// backward of tp_combine (without residual / BN: those are handled by the caller)
//   dy [nodes][4][n0+n], dz0 [nodes][n0] (gradient w.r.t. the l=0 pre-activations = bias gradient rows)
// ------------------------------------------------------------------------------------------------
template <bool GATE>
__global__ void tp_combine_bwd_kernel(const float* __restrict__ y, const float* __restrict__ node_attr, int nodes,
                                      int n, const float* __restrict__ bias, const float* __restrict__ dout,
                                      float* __restrict__ dy, float* __restrict__ dz0) {
  const int n0 = GATE ? 2 * n : n;
  const int n_out = n0 + n;
  const int64_t total = (int64_t)nodes * n;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t node = idx / n;
    const int w = (int)(idx - node * n);
    const float a0 = node_attr[node * 4 + 0], ax = node_attr[node * 4 + 1], ay = node_attr[node * 4 + 2],
                az = node_attr[node * 4 + 3];
    const float* y0 = y + node * 4 * n_out;
    const float* y1 = y0 + n_out;
    const float* y2 = y1 + n_out;
    const float* y3 = y2 + n_out;
    const float* d = dout + node * 4 * n;
    const float ds = d[w], dvx = d[n + w], dvy = d[2 * n + w], dvz = d[3 * n + w];
    float dzs, dzg = 0.f, dzx, dzy, dzz;
    if (GATE) {
      float zs = a0 * y0[w] + ax * y1[w] + ay * y2[w] + az * y3[w];
      float zg = a0 * y0[n + w] + ax * y1[n + w] + ay * y2[n + w] + az * y3[n + w];
      if (bias != nullptr) {
        zs += bias[w];
        zg += bias[n + w];
      }
      const float t = y0[n0 + w];
      const float zx = ax * t + a0 * y1[n0 + w], zy = ay * t + a0 * y2[n0 + w], zz = az * t + a0 * y3[n0 + w];
      const float g = sig_gate(zg);
      dzs = ds * silu_gate_grad(zs);
      dzg = sig_gate_grad(zg) * (dvx * zx + dvy * zy + dvz * zz);
      dzx = g * dvx;
      dzy = g * dvy;
      dzz = g * dvz;
    } else {
      dzs = ds;
      dzx = dvx;
      dzy = dvy;
      dzz = dvz;
    }
    float* e0 = dy + node * 4 * n_out;
    float* e1 = e0 + n_out;
    float* e2 = e1 + n_out;
    float* e3 = e2 + n_out;
    e0[w] = a0 * dzs;
    e1[w] = ax * dzs;
    e2[w] = ay * dzs;
    e3[w] = az * dzs;
    if (GATE) {
      e0[n + w] = a0 * dzg;
      e1[n + w] = ax * dzg;
      e2[n + w] = ay * dzg;
      e3[n + w] = az * dzg;
    }
    e0[n0 + w] = ax * dzx + ay * dzy + az * dzz;
    e1[n0 + w] = a0 * dzx;
    e2[n0 + w] = a0 * dzy;
    e3[n0 + w] = a0 * dzz;
    if (dz0 != nullptr) {
      dz0[node * n0 + w] = dzs;
      if (GATE) dz0[node * n0 + n + w] = dzg;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// weight gradient of the node GEMM: dw_s[k][c] = sum_nodes x[node][0][k] * dy[node][0][c],
//                                   dw_v[k][c] = sum_nodes sum_{p=1..3} x[node][p][k] * dy[node][p][c]
// x = cat_K(x0, x1); dy = cat_cols(dy0 [.. split], dy1). 32x32 output tile per block, rows streamed through smem;
// grid.z splits the rows into slabs whose partial results are reduced by colsum_stage2 (deterministic).
// ------------------------------------------------------------------------------------------------
constexpr int kWgTile = 32, kWgRows = 32;

__global__ void __launch_bounds__(256) node_wgrad_kernel(const float* __restrict__ x0, const float* __restrict__ x1,
                                                       const float* __restrict__ dy0, const float* __restrict__ dy1,
                                                       int split, int nodes, int n_in, int n_out, int cls,
                                                       int64_t rows_per_slab, float* __restrict__ partial) {
  __shared__ float sx[kWgRows][kWgTile + 1];
  __shared__ float sd[kWgRows][kWgTile + 1];
  const int K = x1 ? 2 * n_in : n_in;
  const int k0 = blockIdx.x * kWgTile, c0 = blockIdx.y * kWgTile;
  const int64_t rows = cls == 0 ? (int64_t)nodes : (int64_t)nodes * 3;
  const int64_t r_begin = (int64_t)blockIdx.z * rows_per_slab;
  const int64_t r_end = min(rows, r_begin + rows_per_slab);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  float acc[4] = {0.f, 0.f, 0.f, 0.f};                       // output (k0 + ty*4 + i, c0 + tx)
  const int n_out1 = n_out - split;
  for (int64_t rb = r_begin; rb < r_end; rb += kWgRows) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rr = ty * 4 + i;
      const int64_t r = rb + rr;
      float vx = 0.f, vd = 0.f;
      if (r < r_end) {
        const int64_t plane = cls == 0 ? r * 4 : (r / 3) * 4 + 1 + (r % 3);
        const int k = k0 + tx, c = c0 + tx;
        if (k < K) vx = k < n_in ? x0[plane * n_in + k] : x1[plane * n_in + (k - n_in)];
        if (c < n_out) vd = c < split ? dy0[plane * split + c] : dy1[plane * n_out1 + (c - split)];
      }
      sx[rr][tx] = vx;
      sd[rr][tx] = vd;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < kWgRows; ++rr) {
      const float dv = sd[rr][tx];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(sx[rr][ty * 4 + i], dv, acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + ty * 4 + i, c = c0 + tx;
    if (k < K && c < n_out) partial[((int64_t)blockIdx.z * K + k) * n_out + c] = acc[i];
  }
}

// ------------------------------------------------------------------------------------------------
// embedding backward (inputs carry no gradient): per-node contributions to (w_embed [6][n], bias [n]) -> [nodes][7][n]
// ------------------------------------------------------------------------------------------------
__global__ void embed_bwd_kernel(const float* __restrict__ x_in, const float* __restrict__ node_attr,
                                 const float* __restrict__ dh, int nodes, int n, float* __restrict__ contrib) {
  const int64_t total = (int64_t)nodes * n;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t node = idx / n;
    const int c = (int)(idx - node * n);
    const float* x = x_in + node * 7;
    const float a0 = node_attr[node * 4 + 0], ax = node_attr[node * 4 + 1], ay = node_attr[node * 4 + 2],
                az = node_attr[node * 4 + 3];
    const float* d = dh + node * 4 * n;
    const float ds = d[c], dx = d[n + c], dy = d[2 * n + c], dz = d[3 * n + c];
    const float pdot = x[0] * ax + x[1] * ay + x[2] * az;
    const float vdot = x[3] * ax + x[4] * ay + x[5] * az;
    float* o = contrib + node * 7 * n;
    o[0 * n + c] = a0 * (x[0] * dx + x[1] * dy + x[2] * dz);  // W_vec0 -> 1o
    o[1 * n + c] = a0 * (x[3] * dx + x[4] * dy + x[5] * dz);  // W_vec1 -> 1o
    o[2 * n + c] = pdot * ds;                                 // W_vec0 -> 0e (/sqrt3 folded)
    o[3 * n + c] = vdot * ds;                                 // W_vec1 -> 0e
    o[4 * n + c] = a0 * x[6] * ds;                            // W_sc -> 0e
    o[5 * n + c] = x[6] * (ax * dx + ay * dy + az * dz);      // W_sc -> 1o
    o[6 * n + c] = ds;                                        // bias
  }
}

// head backward: dh [nodes][4][n] and per-node contributions to w_head [2][n][2] -> [nodes][4][n] (ws0, ws1, wv0, wv1)
__global__ void head_bwd_kernel(const float* __restrict__ h, const float* __restrict__ node_attr,
                                const float* __restrict__ w_head, const float* __restrict__ dpred, int nodes, int n,
                                float* __restrict__ dh, float* __restrict__ contrib) {
  const int64_t total = (int64_t)nodes * n;
  const float* ws = w_head;
  const float* wv = w_head + 2 * n;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t node = idx / n;
    const int u = (int)(idx - node * n);
    const float a0 = node_attr[node * 4 + 0];
    const float a1[3] = {node_attr[node * 4 + 1], node_attr[node * 4 + 2], node_attr[node * 4 + 3]};
    const float* dp = dpred + node * 6;
    const float* hn = h + node * 4 * n;
    const float s = hn[u], v[3] = {hn[n + u], hn[2 * n + u], hn[3 * n + u]};
    float dt[2], dd[2][3];
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      dt[o] = a1[0] * dp[o * 3 + 0] + a1[1] * dp[o * 3 + 1] + a1[2] * dp[o * 3 + 2];
#pragma unroll
      for (int k = 0; k < 3; ++k) dd[o][k] = a0 * dp[o * 3 + k];
    }
    float* dhn = dh + node * 4 * n;
    dhn[u] = ws[u * 2 + 0] * dt[0] + ws[u * 2 + 1] * dt[1];
#pragma unroll
    for (int k = 0; k < 3; ++k) dhn[(1 + k) * n + u] = wv[u * 2 + 0] * dd[0][k] + wv[u * 2 + 1] * dd[1][k];
    float* c = contrib + node * 4 * n;
    c[0 * n + u] = s * dt[0];
    c[1 * n + u] = s * dt[1];
    c[2 * n + u] = v[0] * dd[0][0] + v[1] * dd[0][1] + v[2] * dd[0][2];
    c[3 * n + u] = v[0] * dd[1][0] + v[1] * dd[1][1] + v[2] * dd[1][2];
  }
}

// out = a + b + c (c may be NULL): the three gradient streams that meet at a layer input
__global__ void add3_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                            int64_t count, float* __restrict__ out) {
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < count;
       idx += (int64_t)gridDim.x * blockDim.x) {
    float v = a[idx] + b[idx];
    if (c != nullptr) v += c[idx];
    out[idx] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// e3nn BatchNorm coefficients (models/segnn/segnn.py:233-235): one thread per channel turns the column sums into the
// folded affine of the forward pass (+ running-statistics update) or into the (A, B, C) coefficients of the backward
// pass and the parameter gradients.  Replaces ~60 tiny elementwise launches per BatchNorm and step.
// ------------------------------------------------------------------------------------------------
__global__ void bn_coeffs_fwd_kernel(const float* __restrict__ sums, const float* __restrict__ sq, int v_planes, int n,
                                     double rows, double deg, const float* __restrict__ weight,
                                     const float* __restrict__ bias, float* __restrict__ running_mean,
                                     float* __restrict__ running_var, double eps, double momentum, int training,
                                     int update, float* __restrict__ cols, float* __restrict__ stats) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n) return;
  double mean, var_s, var_v;
  if (training) {
    mean = (double)sums[w] / rows;
    var_s = fmax((double)sq[w] / rows - mean * mean, 0.0);
    double sv = 0.0;
    for (int p = 0; p < v_planes; ++p) sv += (double)sq[n + p * n + w];
    var_v = sv / (3.0 * rows);
  } else {
    mean = running_mean[w];
    var_s = running_var[w];
    var_v = running_var[n + w];
  }
  const double rs_s = rsqrt(var_s + eps), rs_v = rsqrt(var_v + eps);
  const double mul_s = weight[w] * rs_s, mul_v = weight[n + w] * rs_v;
  float* mulc = cols;
  float* addc = cols + 4 * n;
  mulc[w] = (float)mul_s;
  mulc[n + w] = mulc[2 * n + w] = mulc[3 * n + w] = (float)mul_v;
  addc[w] = (float)(deg * ((double)bias[w] - mean * mul_s));
  addc[n + w] = addc[2 * n + w] = addc[3 * n + w] = 0.f;
  stats[w] = (float)mean;
  stats[n + w] = (float)var_s;
  stats[2 * n + w] = (float)var_v;
  stats[3 * n + w] = (float)rs_s;
  stats[4 * n + w] = (float)rs_v;
  if (training && update) {
    running_mean[w] = (float)((1.0 - momentum) * running_mean[w] + momentum * mean);
    running_var[w] = (float)((1.0 - momentum) * running_var[w] + momentum * var_s);
    running_var[n + w] = (float)((1.0 - momentum) * running_var[n + w] + momentum * var_v);
  }
}

__global__ void bn_coeffs_bwd_kernel(const float* __restrict__ sum_g, const float* __restrict__ sum_gx, int n,
                                     double rows, double deg, const float* __restrict__ weight,
                                     const float* __restrict__ stats, int training, float* __restrict__ cols,
                                     float* __restrict__ edge, float* __restrict__ dparam) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n) return;
  const double mean = stats[w], rs_s = stats[3 * n + w], rs_v = stats[4 * n + w];
  const double w_s = weight[w], w_v = weight[n + w];
  const double sg_s = deg * (double)sum_g[w];
  const double sgx_s = sum_gx[w];
  const double sgx_v = (double)sum_gx[n + w] + (double)sum_gx[2 * n + w] + (double)sum_gx[3 * n + w];
  const double dgamma_s = rs_s * (sgx_s - mean * sg_s), dgamma_v = rs_v * sgx_v;
  const double A_s = w_s * rs_s, A_v = w_v * rs_v;
  double B_s = 0.0, B_v = 0.0, C_s = 0.0;
  if (training) {
    const double c1 = sg_s / rows, c2 = dgamma_s / rows;
    B_s = -w_s * rs_s * rs_s * c2;
    C_s = -w_s * rs_s * c1 - B_s * mean;
    B_v = -w_v * rs_v * rs_v * rs_v * sgx_v / (3.0 * rows);
  }
  float* A4 = cols;
  float* B4 = cols + 4 * n;
  float* C4 = cols + 8 * n;
  A4[w] = (float)A_s;
  B4[w] = (float)B_s;
  C4[w] = (float)C_s;
#pragma unroll
  for (int p = 1; p < 4; ++p) {
    A4[p * n + w] = (float)A_v;
    B4[p * n + w] = (float)B_v;
    C4[p * n + w] = 0.f;
  }
  edge[w] = (float)A_s;
  edge[n + w] = (float)A_v;
  edge[2 * n + w] = (float)B_s;
  edge[3 * n + w] = (float)B_v;
  edge[4 * n + w] = (float)C_s;
  dparam[w] = (float)dgamma_s;
  dparam[n + w] = (float)dgamma_v;
  dparam[2 * n + w] = (float)sg_s;
}

static inline int grid_for_train(int64_t total, int threads) {
  int64_t b = (total + threads - 1) / threads;
  const int64_t cap = 148LL * 32;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace segnn

using namespace segnn;

extern "C" {

int64_t segnn_colsum_workspace(int64_t rows, int cols) {
  const int64_t parts = (rows + kColsumRowsPerBlock - 1) / kColsumRowsPerBlock;
  return parts * cols * (int64_t)sizeof(double);
}

int segnn_colsum(const float* x, const float* y, int64_t rows, int cols, int mode, float* workspace, float* out,
                 segnn_stream_t stream) {
  SEGNN_CHECK_ARG(rows >= 0 && cols >= 1 && mode >= 0 && mode <= 2, "bad arguments");
  SEGNN_CHECK_ARG(out != nullptr, "null out");
  if (rows == 0) {
    cudaMemsetAsync(out, 0, sizeof(float) * cols, (cudaStream_t)stream);
    return SEGNN_OK;
  }
  SEGNN_CHECK_ARG(x && workspace && (mode != 2 || y), "null pointer");
  const int64_t parts = (rows + kColsumRowsPerBlock - 1) / kColsumRowsPerBlock;
  SEGNN_CHECK_ARG(parts <= 65535, "too many rows for one colsum call");
  if (parts <= kColsumSmallParts) {
    colsum_small<<<(cols + 31) / 32, 1024, 0, (cudaStream_t)stream>>>(x, y, rows, cols, mode, out);
    SEGNN_CHECK_LAUNCH();
    return SEGNN_OK;
  }
  dim3 grid((cols + 31) / 32, (unsigned)parts);
  double* partial = reinterpret_cast<double*>(workspace);
  colsum_stage1<<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, rows, cols, mode, partial);
  if (parts > 64)
    colsum_stage2_wide<double><<<(cols + 31) / 32, 1024, 0, (cudaStream_t)stream>>>(partial, (int)parts, cols, out);
  else
    colsum_stage2<double><<<(cols + 127) / 128, 128, 0, (cudaStream_t)stream>>>(partial, (int)parts, cols, out);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_colsum2(const float* xa, const float* ya, int64_t rows_a, int cols_a, int mode_a, float* out_a,
                  const float* xb, const float* yb, int64_t rows_b, int cols_b, int mode_b, float* out_b,
                  float* workspace, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(rows_a >= 0 && cols_a >= 1 && mode_a >= 0 && mode_a <= 2 && rows_b >= 0 && cols_b >= 1 &&
                      mode_b >= 0 && mode_b <= 2, "bad arguments");
  const int64_t parts_a = (rows_a + kColsumRowsPerBlock - 1) / kColsumRowsPerBlock;
  const int64_t parts_b = (rows_b + kColsumRowsPerBlock - 1) / kColsumRowsPerBlock;
  if (rows_a > 0 && rows_b > 0 && parts_a <= kColsumSmallParts && parts_b <= kColsumSmallParts) {
    SEGNN_CHECK_ARG(xa && xb && out_a && out_b && (mode_a != 2 || ya) && (mode_b != 2 || yb), "null pointer");
    const int cmax = cols_a > cols_b ? cols_a : cols_b;
    colsum_small2<<<dim3((cmax + 31) / 32, 2), 1024, 0, (cudaStream_t)stream>>>(
        ColsumJob{xa, ya, rows_a, cols_a, mode_a, out_a}, ColsumJob{xb, yb, rows_b, cols_b, mode_b, out_b});
    SEGNN_CHECK_LAUNCH();
    return SEGNN_OK;
  }
  // many rows: the two-stage reductions one after the other (the workspace is reused in stream order)
  int rc = segnn_colsum(xa, ya, rows_a, cols_a, mode_a, workspace, out_a, stream);
  if (rc != SEGNN_OK) return rc;
  return segnn_colsum(xb, yb, rows_b, cols_b, mode_b, workspace, out_b, stream);
}

int segnn_lincomb(const float* dy, const float* x, const float* A, const float* B, const float* C, int64_t rows,
                  int cols, float* out, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(rows >= 0 && cols >= 1, "bad sizes");
  if (rows == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(dy && A && out && (x == nullptr || B != nullptr), "null pointer");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if ((cols & 3) == 0 && rows >= 4096 && al16(dy) && al16(x) && al16(A) && al16(B) && al16(C) && al16(out)) {
    const int cols4 = cols / 4;
    const int bx = cols4 >= 256 ? 256 : ((cols4 + 31) & ~31);
    const int by = 1024 / bx;
    int64_t blocks = (rows + by - 1) / by;
    if (blocks > 8 * 148) blocks = 8 * 148;
    lincomb4_kernel<<<(unsigned)blocks, dim3(bx, by), 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(A),
        reinterpret_cast<const float4*>(B), reinterpret_cast<const float4*>(C), rows, cols4,
        reinterpret_cast<float4*>(out));
    SEGNN_CHECK_LAUNCH();
    return SEGNN_OK;
  }
  lincomb_kernel<<<grid_for_train(rows * cols, 256), 256, 0, (cudaStream_t)stream>>>(dy, x, A, B, C, rows, cols, out);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_bn_coeffs_fwd(const float* sums, const float* sq, int v_planes, int n, double rows, double deg,
                        const float* weight, const float* bias, float* running_mean, float* running_var, double eps,
                        double momentum, int training, int update, float* cols, float* stats, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(n >= 1 && rows > 0 && v_planes >= 1 && v_planes <= 3, "bad arguments");
  SEGNN_CHECK_ARG(weight && bias && running_mean && running_var && cols && stats && (!training || (sums && sq)),
                  "null pointer");
  bn_coeffs_fwd_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, sq, v_planes, n, rows, deg, weight, bias,
                                                                         running_mean, running_var, eps, momentum,
                                                                         training, update, cols, stats);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_bn_coeffs_bwd(const float* sum_g, const float* sum_gx, int n, double rows, double deg, const float* weight,
                        const float* stats, int training, float* cols, float* edge, float* dparam,
                        segnn_stream_t stream) {
  SEGNN_CHECK_ARG(n >= 1 && rows > 0, "bad arguments");
  SEGNN_CHECK_ARG(sum_g && sum_gx && weight && stats && cols && edge && dparam, "null pointer");
  bn_coeffs_bwd_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sum_g, sum_gx, n, rows, deg, weight, stats,
                                                                         training, cols, edge, dparam);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_add3(const float* a, const float* b, const float* c, int64_t count, float* out, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(count >= 0, "bad size");
  if (count == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(a && b && out, "null pointer");
  add3_kernel<<<grid_for_train(count, 256), 256, 0, (cudaStream_t)stream>>>(a, b, c, count, out);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_tp_combine_bwd(const float* y, const float* node_attr, int nodes, int n, int gate, const float* bias,
                         const float* dout, float* dy, float* dz0, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(nodes >= 0 && n >= 1, "bad sizes");
  if (nodes == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(y && node_attr && dout && dy, "null pointer");
  const int grid = grid_for_train((int64_t)nodes * n, 256);
  if (gate)
    tp_combine_bwd_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(y, node_attr, nodes, n, bias, dout, dy, dz0);
  else
    tp_combine_bwd_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(y, node_attr, nodes, n, bias, dout, dy, dz0);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

// Row slabs of the weight-gradient GEMM: enough CTAs to fill the GPU even for a few hundred nodes (the README training
// batch has 320), at least 32 rows per slab, at most 64 slabs. Depends on the sizes only => deterministic.
static int64_t wgrad_slabs(int nodes, int K, int n_out) {
  const int64_t rows = (int64_t)nodes * 3;
  const int64_t tiles = (int64_t)((K + kWgTile - 1) / kWgTile) * ((n_out + kWgTile - 1) / kWgTile);
  int64_t slabs = (2 * 148 + tiles - 1) / tiles;
  const int64_t by_rows = (rows + 4095) / 4096;
  if (slabs < by_rows) slabs = by_rows;
  const int64_t max_by_rows = (nodes + kWgRows - 1) / kWgRows;  // the scalar class has `nodes` rows
  if (slabs > max_by_rows) slabs = max_by_rows;
  if (slabs > 64) slabs = 64;
  if (slabs < 1) slabs = 1;
  return slabs;
}

int64_t segnn_node_gemm_wgrad_workspace(int nodes, int K, int n_out) {
  return wgrad_slabs(nodes, K, n_out) * (int64_t)K * n_out * (int64_t)sizeof(float);
}

int segnn_node_gemm_wgrad(const float* x0, const float* x1, const float* dy0, const float* dy1, int split, int nodes,
                          int n_in, int n_out, float* workspace, float* dw_s, float* dw_v, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(nodes >= 0 && n_in >= 1 && n_out >= 1, "bad sizes");
  const int K = x1 ? 2 * n_in : n_in;
  SEGNN_CHECK_ARG(dw_s && dw_v, "null output");
  if (nodes == 0) {
    cudaMemsetAsync(dw_s, 0, sizeof(float) * K * n_out, (cudaStream_t)stream);
    cudaMemsetAsync(dw_v, 0, sizeof(float) * K * n_out, (cudaStream_t)stream);
    return SEGNN_OK;
  }
  SEGNN_CHECK_ARG(x0 && dy0 && workspace, "null pointer");
  if (dy1 == nullptr) split = n_out;
  SEGNN_CHECK_ARG(split > 0 && split <= n_out, "split out of range");
  for (int cls = 0; cls < 2; ++cls) {
    const int64_t rows = cls == 0 ? (int64_t)nodes : (int64_t)nodes * 3;
    const int64_t slabs = wgrad_slabs(nodes, K, n_out);  // same count as the workspace query
    const int64_t per = ((rows + slabs - 1) / slabs + kWgRows - 1) / kWgRows * kWgRows;
    dim3 grid((K + kWgTile - 1) / kWgTile, (n_out + kWgTile - 1) / kWgTile, (unsigned)slabs);
    node_wgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x0, x1, dy0, dy1, split, nodes, n_in, n_out, cls, per,
                                                             workspace);
    colsum_stage2<float><<<(K * n_out + 127) / 128, 128, 0, (cudaStream_t)stream>>>(workspace, (int)slabs, K * n_out,
                                                                           cls == 0 ? dw_s : dw_v);
  }
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_embed_bwd(const float* x_in, const float* node_attr, const float* dh, int nodes, int n, float* contrib,
                    segnn_stream_t stream) {
  SEGNN_CHECK_ARG(nodes >= 0 && n >= 1, "bad sizes");
  if (nodes == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(x_in && node_attr && dh && contrib, "null pointer");
  embed_bwd_kernel<<<grid_for_train((int64_t)nodes * n, 256), 256, 0, (cudaStream_t)stream>>>(x_in, node_attr, dh,
                                                                                            nodes, n, contrib);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_head_bwd(const float* h, const float* node_attr, const float* w_head, const float* dpred, int nodes, int n,
                   float* dh, float* contrib, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(nodes >= 0 && n >= 1, "bad sizes");
  if (nodes == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(h && node_attr && w_head && dpred && dh && contrib, "null pointer");
  head_bwd_kernel<<<grid_for_train((int64_t)nodes * n, 256), 256, 0, (cudaStream_t)stream>>>(h, node_attr, w_head,
                                                                                           dpred, nodes, n, dh,
                                                                                           contrib);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

}  // extern "C"
