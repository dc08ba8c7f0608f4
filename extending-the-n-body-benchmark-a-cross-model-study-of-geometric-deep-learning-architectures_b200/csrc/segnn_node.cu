// Node-level kernels of the SEGNN hot path: graph/geometry materialisation (tests only), K1 prep,
// K2 embedding, the node GEMM, the attribute combine, the head and the self-feed integrator.
#include <stdarg.h>

#include <cuda_fp16.h>

#include "segnn_common.cuh"

namespace segnn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ------------------------------------------------------------------------------------------------
// graph enumeration + edge attributes (materialised for parity tests / legacy consumers only)
// ------------------------------------------------------------------------------------------------

__global__ void edge_index_kernel(int B, int N, int64_t* __restrict__ ei) {
  const int64_t per = (int64_t)N * (N - 1);
  const int64_t E = per * B;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t g = e / per;
    int64_t r = e - g * per;
    int a = (int)(r / (N - 1));
    int b = (int)(r - (int64_t)a * (N - 1));
    b += (b >= a);
    ei[e] = g * N + a;      // source / sender
    ei[E + e] = g * N + b;  // target / receiver
  }
}

__global__ void edge_attr_kernel(const float* __restrict__ pos, const float* __restrict__ mass, int B, int N,
                                 float* __restrict__ edge_attr, float* __restrict__ add) {
  const int64_t per = (int64_t)N * (N - 1);
  const int64_t E = per * B;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t g = e / per;
    int64_t r = e - g * per;
    int a = (int)(r / (N - 1));
    int b = (int)(r - (int64_t)a * (N - 1));
    b += (b >= a);
    int64_t s = g * N + a, t = g * N + b;
    float ux, uy, uz, len;
    unit_vec(pos[s * 3 + 0] - pos[t * 3 + 0], pos[s * 3 + 1] - pos[t * 3 + 1], pos[s * 3 + 2] - pos[t * 3 + 2], ux,
             uy, uz, len);
    edge_attr[e * 4 + 0] = kY0;
    edge_attr[e * 4 + 1] = kY1 * ux;
    edge_attr[e * 4 + 2] = kY1 * uy;
    edge_attr[e * 4 + 3] = kY1 * uz;
    add[e * 2 + 0] = len;
    add[e * 2 + 1] = mass[s] * mass[t];
  }
}

// ------------------------------------------------------------------------------------------------
// K1 prep: node_attr = mean_j Y(r_j - r_i) + Y(v_i) with the l=0 slot forced to 1; x = [pos - mean_xyz, vel, |vel|]
// One thread per receiver; the graph's positions are staged in shared memory in sender tiles.
// ------------------------------------------------------------------------------------------------

constexpr int kPrepThreads = 128;

__global__ void __launch_bounds__(kPrepThreads) prep_kernel(const float* __restrict__ pos,
                                                           const float* __restrict__ vel, int B, int N,
                                                           float* __restrict__ x_in, float* __restrict__ node_attr) {
  // grid.y = graph, grid.x = receiver tile inside the graph
  __shared__ float sp[kPrepThreads * 3];
  const int g = blockIdx.y;
  const int i = blockIdx.x * kPrepThreads + threadIdx.x;
  const int64_t base = (int64_t)g * N;
  const bool live = i < N;
  float px = 0.f, py = 0.f, pz = 0.f;
  if (live) {
    px = pos[(base + i) * 3 + 0];
    py = pos[(base + i) * 3 + 1];
    pz = pos[(base + i) * 3 + 2];
  }
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (int j0 = 0; j0 < N; j0 += kPrepThreads) {
    int cnt = min(kPrepThreads, N - j0);
    __syncthreads();
    for (int t = threadIdx.x; t < cnt * 3; t += kPrepThreads) sp[t] = pos[(base + j0) * 3 + t];
    __syncthreads();
    if (live) {
      for (int j = 0; j < cnt; ++j) {
        if (j0 + j == i) continue;
        float ux, uy, uz, len;
        unit_vec(sp[j * 3 + 0] - px, sp[j * 3 + 1] - py, sp[j * 3 + 2] - pz, ux, uy, uz, len);
        sx += ux;
        sy += uy;
        sz += uz;
      }
    }
  }
  if (!live) return;
  const int64_t node = base + i;
  float vx = vel[node * 3 + 0], vy = vel[node * 3 + 1], vz = vel[node * 3 + 2];
  float ux, uy, uz, vlen;
  unit_vec(vx, vy, vz, ux, uy, uz, vlen);
  float inv_deg = N > 1 ? 1.0f / (float)(N - 1) : 0.0f;
  node_attr[node * 4 + 0] = 1.0f;
  node_attr[node * 4 + 1] = kY1 * (sx * inv_deg) + kY1 * ux;
  node_attr[node * 4 + 2] = kY1 * (sy * inv_deg) + kY1 * uy;
  node_attr[node * 4 + 3] = kY1 * (sz * inv_deg) + kY1 * uz;
  float m = (px + py + pz) / 3.0f;  // reference quirk: mean over xyz of the node (o3_building_blocks.py:274)
  float* x = x_in + node * 7;
  x[0] = px - m;
  x[1] = py - m;
  x[2] = pz - m;
  x[3] = vx;
  x[4] = vy;
  x[5] = vz;
  x[6] = vlen;
}

// ------------------------------------------------------------------------------------------------
// K2 embedding: O3TensorProduct(2x1o + 1x0e -> n x0e + n x1o) steered by node_attr; K = 3, elementwise.
// ------------------------------------------------------------------------------------------------

__global__ void embed_kernel(const float* __restrict__ x_in, const float* __restrict__ node_attr,
                             const float* __restrict__ w, const float* __restrict__ bias, int nodes, int n,
                             float* __restrict__ h, __half* __restrict__ h16 = nullptr) {
  const int64_t total = (int64_t)nodes * n;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t node = idx / n;
    const int c = (int)(idx - node * n);
    const float* x = x_in + node * 7;
    const float a0 = node_attr[node * 4 + 0], ax = node_attr[node * 4 + 1], ay = node_attr[node * 4 + 2],
                az = node_attr[node * 4 + 3];
    const float w_p1 = w[0 * n + c], w_v1 = w[1 * n + c], w_p0 = w[2 * n + c], w_v0 = w[3 * n + c],
                w_s0 = w[4 * n + c], w_s1 = w[5 * n + c];
    const float pdot = x[0] * ax + x[1] * ay + x[2] * az;
    const float vdot = x[3] * ax + x[4] * ay + x[5] * az;
    const float s = x[6];
    float* out = h + node * 4 * n;
    const float o0 = a0 * w_s0 * s + w_p0 * pdot + w_v0 * vdot + bias[c];
    const float t = w_s1 * s;
    const float o1 = ax * t + a0 * (w_p1 * x[0] + w_v1 * x[3]);
    const float o2 = ay * t + a0 * (w_p1 * x[1] + w_v1 * x[4]);
    const float o3 = az * t + a0 * (w_p1 * x[2] + w_v1 * x[5]);
    out[c] = o0;
    out[1 * n + c] = o1;
    out[2 * n + c] = o2;
    out[3 * n + c] = o3;
    if (h16 != nullptr) {  // fp16 operand copy for the tensor-core node GEMMs (segnn_node_gemm_tc_x16)
      __half* o16 = h16 + node * 4 * n;
      o16[c] = __float2half_rn(o0);
      o16[1 * n + c] = __float2half_rn(o1);
      o16[2 * n + c] = __float2half_rn(o2);
      o16[3 * n + c] = __float2half_rn(o3);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Node GEMM (fp32 FFMA): y[node][c][:] = cat_K(x0[node][c], x1[node][c]) @ (c == 0 ? w_s : w_v)
// grid.z selects the row class (0: scalar planes, 1: the three vector planes), so one launch covers both.
// 64x64 tile, BK = 16, 256 threads, 4x4 outputs per thread.
// ------------------------------------------------------------------------------------------------

constexpr int kGemmBM = 64, kGemmBN = 64, kGemmBK = 16;

__global__ void __launch_bounds__(256) node_gemm_kernel(const float* __restrict__ x0, const float* __restrict__ x1,
                                                      int nodes, int n_in, const float* __restrict__ w_s,
                                                      const float* __restrict__ w_v, const float* __restrict__ bias,
                                                      int n_bias, int n_out, float* __restrict__ y0,
                                                      float* __restrict__ y1, int split) {
  __shared__ float As[kGemmBK][kGemmBM + 4];
  __shared__ float Bs[kGemmBK][kGemmBN + 4];
  const int cls = blockIdx.z;  // 0 scalar rows, 1 vector rows
  const int64_t rows = cls == 0 ? (int64_t)nodes : (int64_t)nodes * 3;
  const int64_t row0 = (int64_t)blockIdx.x * kGemmBM;
  if (row0 >= rows) return;
  const int col0 = blockIdx.y * kGemmBN;
  const float* __restrict__ W = cls == 0 ? w_s : w_v;
  const int K = x1 ? 2 * n_in : n_in;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;

  // A-tile loader: thread -> (row = tid / 4, 4 consecutive k at (tid % 4) * 4)
  const int a_row = tid >> 2, a_k = (tid & 3) * 4;
  const int64_t ar = row0 + a_row;
  int64_t a_plane = -1;  // (node * 4 + c)
  if (ar < rows) a_plane = cls == 0 ? ar * 4 : (ar / 3) * 4 + 1 + (ar % 3);
  // B-tile loader: thread -> (k = tid / 16, 4 consecutive cols at (tid % 16) * 4)
  const int b_k = tid >> 4, b_c = (tid & 15) * 4;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // register prefetch of the next K tile while the current one is being multiplied: at training sizes (a few hundred
  // rows) a launch is a handful of CTAs deep, so the global-load latency of every K step is exposed unless the loads of
  // step k + 1 are in flight during the FMAs of step k
  float ra[4], rb[4];
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = k0 + a_k + q;
      float v = 0.f;
      if (a_plane >= 0 && k < K) v = k < n_in ? x0[a_plane * n_in + k] : x1[a_plane * n_in + (k - n_in)];
      ra[q] = v;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = k0 + b_k, c = col0 + b_c + q;
      rb[q] = (k < K && c < n_out) ? W[(int64_t)k * n_out + c] : 0.f;
    }
  };
  load_tile(0);
  for (int k0 = 0; k0 < K; k0 += kGemmBK) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      As[a_k + q][a_row] = ra[q];
      Bs[b_k][b_c + q] = rb[q];
    }
    __syncthreads();
    if (k0 + kGemmBK < K) load_tile(k0 + kGemmBK);
#pragma unroll
    for (int k = 0; k < kGemmBK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t r = row0 + ty * 4 + i;
    if (r >= rows) continue;
    int64_t plane = cls == 0 ? r * 4 : (r / 3) * 4 + 1 + (r % 3);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = col0 + tx * 4 + j;
      if (c >= n_out) continue;
      float v = acc[i][j];
      if (cls == 0 && bias != nullptr && c < n_bias) v += bias[c];
      if (c < split) y0[plane * split + c] = v;
      else y1[plane * (n_out - split) + (c - split)] = v;
    }
  }
}

// Same GEMM for launches of a few hundred rows (training batches: 64 x 5 bodies = 1280 rows), where the 64 x 64 kernel is
// a handful of CTAs whose every K step waits a full global-load latency (measured 1.1 - 2.1 us per step of 16, 12 - 39 us
// per launch on the README training step): 32 x 64 tiles (twice the CTAs), K steps of 32 through a 4-stage cp.async
// ring, so up to four steps of loads are in flight while one is multiplied.  Same accumulation order (k ascending, one
// fmaf per product) as node_gemm_kernel: bit-identical results.  Needs n_in % 4 == 0 and n_out % 4 == 0 (16-byte copies).
constexpr int kSmallGemmMaxNodes = 4096;  // above: node_gemm_kernel (fp32 mode switches to segnn_node_gemm_tf32x3 there)
// K steps of 64 from K = 256 on (message_layer_1 data gradient, K = 6n: twice the bytes in flight, half the barriers).
constexpr int kSmBM = 32, kSmBN = 64, kSmStages = 4, kSmThreads = 128;
constexpr int kSmBStride = kSmBN + 4;  // floats; rows stay 16-byte aligned
template <int BK>
__host__ __device__ constexpr int sm_stage_floats() { return kSmBM * (BK + 4) + BK * kSmBStride; }

__device__ __forceinline__ void cp_async16(float* dst, const float* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int bytes = valid ? 16 : 0;  // src-size 0: the 16 bytes are zero-filled, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}

template <int kSmBK>
__global__ void __launch_bounds__(kSmThreads) node_gemm_small_kernel(
    const float* __restrict__ x0, const float* __restrict__ x1, int nodes, int n_in, const float* __restrict__ w_s,
    const float* __restrict__ w_v, const float* __restrict__ bias, int n_bias, int n_out, float* __restrict__ y0,
    float* __restrict__ y1, int split) {
  extern __shared__ __align__(16) float sm_ring[];
  constexpr int kSmAStride = kSmBK + 4, kSmStageFloats = sm_stage_floats<kSmBK>();
  constexpr int kAChunks = kSmBM * kSmBK / 4 / kSmThreads, kBChunks = kSmBK * kSmBN / 4 / kSmThreads;
  const int cls = blockIdx.z;
  const int64_t rows = cls == 0 ? (int64_t)nodes : (int64_t)nodes * 3;
  const int64_t row0 = (int64_t)blockIdx.x * kSmBM;
  if (row0 >= rows) return;
  const int col0 = blockIdx.y * kSmBN;
  const float* __restrict__ W = cls == 0 ? w_s : w_v;
  const int K = x1 ? 2 * n_in : n_in;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;  // 16 column groups x 8 row groups, 4 x 4 outputs each
  const int ksteps = (K + kSmBK - 1) / kSmBK;

  // A copies: 32 rows x BK / 4 chunks of 4 k (two / four per thread); B copies: BK k x 16 chunks of 4 columns
  int64_t a_plane[kAChunks];
  int a_row[kAChunks], a_k[kAChunks];
#pragma unroll
  for (int q = 0; q < kAChunks; ++q) {
    const int c = tid + q * kSmThreads;
    a_row[q] = c / (kSmBK / 4);
    a_k[q] = (c % (kSmBK / 4)) * 4;
    const int64_t r = row0 + a_row[q];
    a_plane[q] = r < rows ? (cls == 0 ? r * 4 : (r / 3) * 4 + 1 + (r % 3)) : -1;
  }
  auto issue = [&](int step) {
    if (step < ksteps) {
      float* As = sm_ring + (step % kSmStages) * kSmStageFloats;
      float* Bs = As + kSmBM * kSmAStride;
      const int k0 = step * kSmBK;
#pragma unroll
      for (int q = 0; q < kAChunks; ++q) {
        const int k = k0 + a_k[q];
        const bool ok = a_plane[q] >= 0 && k < K;
        const float* src = !ok ? x0 : (k < n_in ? x0 + a_plane[q] * n_in + k : x1 + a_plane[q] * n_in + (k - n_in));
        cp_async16(As + a_row[q] * kSmAStride + a_k[q], src, ok);
      }
#pragma unroll
      for (int q = 0; q < kBChunks; ++q) {
        const int c = tid + q * kSmThreads;
        const int bk = c >> 4, bc = (c & 15) * 4;
        const bool ok = k0 + bk < K && col0 + bc < n_out;
        cp_async16(Bs + bk * kSmBStride + bc, ok ? W + (int64_t)(k0 + bk) * n_out + col0 + bc : W, ok);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");  // one group per step, empty past the end: fixed wait depth
  };

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
  for (int s = 0; s < kSmStages - 1; ++s) issue(s);
  for (int step = 0; step < ksteps; ++step) {
    asm volatile("cp.async.wait_group %0;" ::"n"(kSmStages - 2) : "memory");  // this step's group has landed
    __syncthreads();  // ... for every thread; and every thread is done with the stage the next issue overwrites
    issue(step + kSmStages - 1);
    const float* As = sm_ring + (step % kSmStages) * kSmStageFloats;
    const float* Bs = As + kSmBM * kSmAStride;
#pragma unroll
    for (int k4 = 0; k4 < kSmBK; k4 += 4) {
      float4 a4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a4[i] = *reinterpret_cast<const float4*>(As + (ty * 4 + i) * kSmAStride + k4);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4 b = *reinterpret_cast<const float4*>(Bs + (k4 + kk) * kSmBStride + tx * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float a = kk == 0 ? a4[i].x : kk == 1 ? a4[i].y : kk == 2 ? a4[i].z : a4[i].w;
          acc[i][0] = fmaf(a, b.x, acc[i][0]);
          acc[i][1] = fmaf(a, b.y, acc[i][1]);
          acc[i][2] = fmaf(a, b.z, acc[i][2]);
          acc[i][3] = fmaf(a, b.w, acc[i][3]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = row0 + ty * 4 + i;
    if (r >= rows) continue;
    const int64_t plane = cls == 0 ? r * 4 : (r / 3) * 4 + 1 + (r % 3);
    const int c = col0 + tx * 4;
    if (c >= n_out) continue;  // n_out % 4 == 0: a group of four columns is inside or outside
    float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    if (cls == 0 && bias != nullptr) {
      if (c + 0 < n_bias) v.x += bias[c + 0];
      if (c + 1 < n_bias) v.y += bias[c + 1];
      if (c + 2 < n_bias) v.z += bias[c + 2];
      if (c + 3 < n_bias) v.w += bias[c + 3];
    }
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (c + j < split) y0[plane * split + c + j] = vv[j];
      else y1[plane * (n_out - split) + (c + j - split)] = vv[j];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Attribute combine + gate / residual / eval-BN epilogue of a node-level tensor product.
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ float ldy(const float* p) { return *p; }
__device__ __forceinline__ float ldy(const __half* p) { return __half2float(*p); }

template <bool GATE, typename YT>
__global__ void tp_combine_kernel(const YT* __restrict__ y, const float* __restrict__ node_attr, int nodes, int n,
                                  const float* __restrict__ bias, const float* __restrict__ residual,
                                  const float* __restrict__ bn_mul,
                                  const float* __restrict__ bn_add, float* __restrict__ out) {
  const int n0 = GATE ? 2 * n : n;
  const int n_out = n0 + n;
  const int64_t total = (int64_t)nodes * n;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t node = idx / n;
    const int w = (int)(idx - node * n);
    const float a0 = node_attr[node * 4 + 0], ax = node_attr[node * 4 + 1], ay = node_attr[node * 4 + 2],
                az = node_attr[node * 4 + 3];
    const YT* y0 = y + node * 4 * n_out;
    const YT* y1 = y0 + n_out;
    const YT* y2 = y1 + n_out;
    const YT* y3 = y2 + n_out;
    float zs = a0 * ldy(y0 + w) + ax * ldy(y1 + w) + ay * ldy(y2 + w) + az * ldy(y3 + w);
    if (bias != nullptr) zs += bias[w];
    const float t = ldy(y0 + n0 + w);
    float vx = ax * t + a0 * ldy(y1 + n0 + w);
    float vy = ay * t + a0 * ldy(y2 + n0 + w);
    float vz = az * t + a0 * ldy(y3 + n0 + w);
    if (GATE) {
      float zg = a0 * ldy(y0 + n + w) + ax * ldy(y1 + n + w) + ay * ldy(y2 + n + w) + az * ldy(y3 + n + w);
      if (bias != nullptr) zg += bias[n + w];
      float g = sig_gate(zg);
      zs = silu_gate(zs);
      vx *= g;
      vy *= g;
      vz *= g;
    }
    const int64_t o = node * 4 * n;
    if (residual != nullptr) {
      zs += residual[o + w];
      vx += residual[o + n + w];
      vy += residual[o + 2 * n + w];
      vz += residual[o + 3 * n + w];
    }
    if (bn_mul != nullptr) {
      const float ms = bn_mul[w], mv = bn_mul[n + w];
      zs = fmaf(zs, ms, bn_add[w]);
      vx *= mv;
      vy *= mv;
      vz *= mv;
    }
    out[o + w] = zs;
    out[o + n + w] = vx;
    out[o + 2 * n + w] = vy;
    out[o + 3 * n + w] = vz;
  }
}

// Same pass on fp16 rows, two channels per thread (half2 loads, float2 stores): the scalar version is bound by its
// load instructions once the rows are half as wide.
template <bool GATE>
__global__ void tp_combine_y16_kernel(const __half* __restrict__ y, const float* __restrict__ node_attr, int nodes, int n,
                                      const float* __restrict__ bias, const float* __restrict__ residual,
                                      const float* __restrict__ bn_mul, const float* __restrict__ bn_add,
                                      float* __restrict__ out, __half* __restrict__ out16) {
  const int n0 = GATE ? 2 * n : n;
  const int n_out = n0 + n;
  const int nh = n / 2;
  const int64_t total = (int64_t)nodes * nh;
  auto ld2 = [](const __half* p) { return __half22float2(*reinterpret_cast<const __half2*>(p)); };
  auto ldf2 = [](const float* p) { return *reinterpret_cast<const float2*>(p); };
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t node = idx / nh;
    const int w = 2 * (int)(idx - node * nh);
    const float4 a = *reinterpret_cast<const float4*>(node_attr + node * 4);
    const __half* y0 = y + node * 4 * n_out;
    const __half* y1 = y0 + n_out;
    const __half* y2 = y1 + n_out;
    const __half* y3 = y2 + n_out;
    auto mix = [&](int off) {  // a0 y0 + a1 . (y1, y2, y3) at column offset `off`
      const float2 p0 = ld2(y0 + off), p1 = ld2(y1 + off), p2 = ld2(y2 + off), p3 = ld2(y3 + off);
      return make_float2(a.x * p0.x + a.y * p1.x + a.z * p2.x + a.w * p3.x,
                         a.x * p0.y + a.y * p1.y + a.z * p2.y + a.w * p3.y);
    };
    float2 zs = mix(w);
    if (bias != nullptr) {
      const float2 b = ldf2(bias + w);
      zs.x += b.x;
      zs.y += b.y;
    }
    const float2 t = ld2(y0 + n0 + w), q1 = ld2(y1 + n0 + w), q2 = ld2(y2 + n0 + w), q3 = ld2(y3 + n0 + w);
    float2 vx = make_float2(a.y * t.x + a.x * q1.x, a.y * t.y + a.x * q1.y);
    float2 vy = make_float2(a.z * t.x + a.x * q2.x, a.z * t.y + a.x * q2.y);
    float2 vz = make_float2(a.w * t.x + a.x * q3.x, a.w * t.y + a.x * q3.y);
    if (GATE) {
      float2 zg = mix(n + w);
      if (bias != nullptr) {
        const float2 b = ldf2(bias + n + w);
        zg.x += b.x;
        zg.y += b.y;
      }
      const float gx = sig_gate(zg.x), gy = sig_gate(zg.y);
      zs.x = silu_gate(zs.x);
      zs.y = silu_gate(zs.y);
      vx.x *= gx; vy.x *= gx; vz.x *= gx;
      vx.y *= gy; vy.y *= gy; vz.y *= gy;
    }
    const int64_t o = node * 4 * n;
    if (residual != nullptr) {
      const float2 r0 = ldf2(residual + o + w), r1 = ldf2(residual + o + n + w), r2 = ldf2(residual + o + 2 * n + w),
                   r3 = ldf2(residual + o + 3 * n + w);
      zs.x += r0.x; zs.y += r0.y;
      vx.x += r1.x; vx.y += r1.y;
      vy.x += r2.x; vy.y += r2.y;
      vz.x += r3.x; vz.y += r3.y;
    }
    if (bn_mul != nullptr) {
      const float2 ms = ldf2(bn_mul + w), mv = ldf2(bn_mul + n + w), ad = ldf2(bn_add + w);
      zs.x = fmaf(zs.x, ms.x, ad.x);
      zs.y = fmaf(zs.y, ms.y, ad.y);
      vx.x *= mv.x; vy.x *= mv.x; vz.x *= mv.x;
      vx.y *= mv.y; vy.y *= mv.y; vz.y *= mv.y;
    }
    if (out != nullptr) {
      *reinterpret_cast<float2*>(out + o + w) = zs;
      *reinterpret_cast<float2*>(out + o + n + w) = vx;
      *reinterpret_cast<float2*>(out + o + 2 * n + w) = vy;
      *reinterpret_cast<float2*>(out + o + 3 * n + w) = vz;
    }
    if (out16 != nullptr) {  // the fp16 operand copy the next tensor-core node GEMM reads (same rounding as its loader)
      *reinterpret_cast<__half2*>(out16 + o + w) = __float22half2_rn(zs);
      *reinterpret_cast<__half2*>(out16 + o + n + w) = __float22half2_rn(vx);
      *reinterpret_cast<__half2*>(out16 + o + 2 * n + w) = __float22half2_rn(vy);
      *reinterpret_cast<__half2*>(out16 + o + 3 * n + w) = __float22half2_rn(vz);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Head: pre_pool2 (h -> 2x1o), one warp per node; and the self-feed integrator.
// ------------------------------------------------------------------------------------------------

__global__ void head_kernel(const float* __restrict__ h, const float* __restrict__ node_attr,
                            const float* __restrict__ w_head, int nodes, int n, float* __restrict__ pred) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float* ws = w_head;          // [n][2]  scalar -> 1o (times a1)
  const float* wv = w_head + 2 * n;  // [n][2]  vector -> 1o (times a0)
  for (int64_t node = warp; node < nodes; node += nwarps) {
    const float* hn = h + node * 4 * n;
    float t[2] = {0.f, 0.f};
    float d[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    for (int u = lane; u < n; u += 32) {
      float s = hn[u], vx = hn[n + u], vy = hn[2 * n + u], vz = hn[3 * n + u];
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        float a = ws[u * 2 + o], b = wv[u * 2 + o];
        t[o] = fmaf(a, s, t[o]);
        d[o][0] = fmaf(b, vx, d[o][0]);
        d[o][1] = fmaf(b, vy, d[o][1]);
        d[o][2] = fmaf(b, vz, d[o][2]);
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        t[o] += __shfl_xor_sync(0xffffffffu, t[o], off);
#pragma unroll
        for (int k = 0; k < 3; ++k) d[o][k] += __shfl_xor_sync(0xffffffffu, d[o][k], off);
      }
    }
    if (lane == 0) {
      const float a0 = node_attr[node * 4 + 0];
      const float a1[3] = {node_attr[node * 4 + 1], node_attr[node * 4 + 2], node_attr[node * 4 + 3]};
#pragma unroll
      for (int o = 0; o < 2; ++o)
#pragma unroll
        for (int k = 0; k < 3; ++k) pred[node * 6 + o * 3 + k] = a1[k] * t[o] + a0 * d[o][k];
    }
  }
}

__global__ void integrate_kernel(const float* __restrict__ pred, float* __restrict__ pos, float* __restrict__ vel,
                                 int nodes, float* __restrict__ traj_pos, float* __restrict__ traj_vel,
                                 const int* __restrict__ frame, int max_frames) {
  const int64_t total = (int64_t)nodes * 3;
  if (frame != nullptr) {  // frame slot chosen on the device so a captured CUDA graph can be replayed
    const int f = *frame;
    if (max_frames > 0 && (f < 0 || f >= max_frames)) {  // never write past the trajectory buffers
      traj_pos = nullptr;
      traj_vel = nullptr;
    }
    const int64_t off = (int64_t)f * total;
    if (traj_pos) traj_pos += off;
    if (traj_vel) traj_vel += off;
  }
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t node = idx / 3;
    const int k = (int)(idx - node * 3);
    float p = pos[idx] + pred[node * 6 + k];
    float v = pred[node * 6 + 3 + k];
    pos[idx] = p;
    vel[idx] = v;
    if (traj_pos) traj_pos[idx] = p;
    if (traj_vel) traj_vel[idx] = v;
  }
}

static inline int grid_for(int64_t total, int threads) {
  int64_t b = (total + threads - 1) / threads;
  const int64_t cap = 148LL * 32;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace segnn

using namespace segnn;

extern "C" {

int segnn_version(void) { return 100; }
const char* segnn_last_error(void) { return g_err; }

int segnn_edge_index(int B, int N, int64_t* edge_index, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(B >= 0 && N >= 1, "need B >= 0, N >= 1");
  int64_t E = (int64_t)B * N * (N - 1);
  if (E == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(edge_index != nullptr, "null edge_index");
  edge_index_kernel<<<grid_for(E, 256), 256, 0, (cudaStream_t)stream>>>(B, N, edge_index);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_edge_attr(const float* pos, const float* mass, int B, int N, float* edge_attr, float* add,
                    segnn_stream_t stream) {
  SEGNN_CHECK_ARG(B >= 0 && N >= 1, "need B >= 0, N >= 1");
  int64_t E = (int64_t)B * N * (N - 1);
  if (E == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(pos && mass && edge_attr && add, "null pointer");
  edge_attr_kernel<<<grid_for(E, 256), 256, 0, (cudaStream_t)stream>>>(pos, mass, B, N, edge_attr, add);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_prep_fwd(const float* pos, const float* vel, int B, int N, float* x_in, float* node_attr,
                   segnn_stream_t stream) {
  SEGNN_CHECK_ARG(B >= 0 && N >= 1, "need B >= 0, N >= 1");
  if (B == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(pos && vel && x_in && node_attr, "null pointer");
  SEGNN_CHECK_ARG(B <= 65535, "B > 65535 graphs per call: split the batch");
  dim3 grid((N + kPrepThreads - 1) / kPrepThreads, B);
  prep_kernel<<<grid, kPrepThreads, 0, (cudaStream_t)stream>>>(pos, vel, B, N, x_in, node_attr);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_embed_fwd(const float* x_in, const float* node_attr, const float* w_embed, const float* bias, int nodes,
                    int n, float* h_out, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(nodes >= 0 && n >= 1, "need nodes >= 0, n >= 1");
  if (nodes == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(x_in && node_attr && w_embed && bias && h_out, "null pointer");
  embed_kernel<<<grid_for((int64_t)nodes * n, 256), 256, 0, (cudaStream_t)stream>>>(x_in, node_attr, w_embed, bias,
                                                                                 nodes, n, h_out);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_embed_fwd_x16(const float* x_in, const float* node_attr, const float* w_embed, const float* bias, int nodes,
                        int n, float* h_out, void* h16_out, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(nodes >= 0 && n >= 1, "need nodes >= 0, n >= 1");
  if (nodes == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(x_in && node_attr && w_embed && bias && h_out && h16_out, "null pointer");
  embed_kernel<<<grid_for((int64_t)nodes * n, 256), 256, 0, (cudaStream_t)stream>>>(
      x_in, node_attr, w_embed, bias, nodes, n, h_out, reinterpret_cast<__half*>(h16_out));
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_node_gemm(const float* x0, const float* x1, int nodes, int n_in, const float* w_s, const float* w_v,
                    const float* bias, int n_bias, int n_out, float* y0, float* y1, int split,
                    segnn_stream_t stream) {
  SEGNN_CHECK_ARG(nodes >= 0 && n_in >= 1 && n_out >= 1, "bad sizes");
  if (nodes == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(x0 && w_s && w_v && y0, "null pointer");
  SEGNN_CHECK_ARG(bias == nullptr || (n_bias >= 0 && n_bias <= n_out), "n_bias out of range");
  if (y1 == nullptr) split = n_out;
  SEGNN_CHECK_ARG(split > 0 && split <= n_out, "split out of range");
  // a few hundred rows (training batches): the latency-tolerant small-tile kernel (same results bit for bit)
  if (nodes <= kSmallGemmMaxNodes && n_in % 4 == 0 && n_out % 4 == 0 && ((uintptr_t)x0 & 15) == 0 &&
      ((uintptr_t)x1 & 15) == 0 && ((uintptr_t)w_s & 15) == 0 && ((uintptr_t)w_v & 15) == 0) {
    static const cudaError_t attr32 = cudaFuncSetAttribute(node_gemm_small_kernel<32>,
                                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                           kSmStages * sm_stage_floats<32>() * (int)sizeof(float));
    static const cudaError_t attr64 = cudaFuncSetAttribute(node_gemm_small_kernel<64>,
                                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                           kSmStages * sm_stage_floats<64>() * (int)sizeof(float));
    if (attr32 != cudaSuccess || attr64 != cudaSuccess) {
      set_error("segnn_node_gemm: cudaFuncSetAttribute: %s",
                cudaGetErrorString(attr32 != cudaSuccess ? attr32 : attr64));
      return SEGNN_E_CUDA;
    }
    dim3 grid_s((unsigned)(((int64_t)nodes * 3 + kSmBM - 1) / kSmBM), (n_out + kSmBN - 1) / kSmBN, 2);
    const int K = x1 ? 2 * n_in : n_in;
    if (K >= 256)
      node_gemm_small_kernel<64><<<grid_s, kSmThreads, kSmStages * sm_stage_floats<64>() * sizeof(float),
                                   (cudaStream_t)stream>>>(x0, x1, nodes, n_in, w_s, w_v, bias, n_bias, n_out, y0, y1,
                                                           split);
    else
      node_gemm_small_kernel<32><<<grid_s, kSmThreads, kSmStages * sm_stage_floats<32>() * sizeof(float),
                                   (cudaStream_t)stream>>>(x0, x1, nodes, n_in, w_s, w_v, bias, n_bias, n_out, y0, y1,
                                                           split);
    SEGNN_CHECK_LAUNCH();
    return SEGNN_OK;
  }
  int64_t row_tiles = ((int64_t)nodes * 3 + kGemmBM - 1) / kGemmBM;
  SEGNN_CHECK_ARG(row_tiles < (1LL << 31), "too many rows");
  dim3 grid((unsigned)row_tiles, (n_out + kGemmBN - 1) / kGemmBN, 2);
  node_gemm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x0, x1, nodes, n_in, w_s, w_v, bias, n_bias, n_out, y0, y1,
                                                           split);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_tp_combine(const float* y, const float* node_attr, int nodes, int n, int gate, const float* bias,
                     const float* residual, const float* bn_mul, const float* bn_add, float* out,
                     segnn_stream_t stream) {
  SEGNN_CHECK_ARG(nodes >= 0 && n >= 1, "bad sizes");
  if (nodes == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(y && node_attr && out, "null pointer");
  SEGNN_CHECK_ARG((bn_mul == nullptr) == (bn_add == nullptr), "bn_mul and bn_add must be given together");
  int grid = grid_for((int64_t)nodes * n, 256);
  if (gate)
    tp_combine_kernel<true, float><<<grid, 256, 0, (cudaStream_t)stream>>>(y, node_attr, nodes, n, bias, residual,
                                                                          bn_mul, bn_add, out);
  else
    tp_combine_kernel<false, float><<<grid, 256, 0, (cudaStream_t)stream>>>(y, node_attr, nodes, n, bias, residual,
                                                                           bn_mul, bn_add, out);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

static int tp_combine_y16_launch(const void* y, const float* node_attr, int nodes, int n, int gate, const float* bias,
                                 const float* residual, const float* bn_mul, const float* bn_add, float* out,
                                 __half* out16, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(nodes >= 0 && n >= 1, "bad sizes");
  if (nodes == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(y && node_attr && (out || out16), "null pointer");
  SEGNN_CHECK_ARG((bn_mul == nullptr) == (bn_add == nullptr), "bn_mul and bn_add must be given together");
  const __half* yh = reinterpret_cast<const __half*>(y);
  if (n % 2 == 0) {  // every row offset is then a multiple of 2 elements: half2 / float2 accesses are aligned
    int grid = grid_for((int64_t)nodes * (n / 2), 256);
    if (gate)
      tp_combine_y16_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(yh, node_attr, nodes, n, bias, residual,
                                                                         bn_mul, bn_add, out, out16);
    else
      tp_combine_y16_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(yh, node_attr, nodes, n, bias, residual,
                                                                          bn_mul, bn_add, out, out16);
    SEGNN_CHECK_LAUNCH();
    return SEGNN_OK;
  }
  SEGNN_CHECK_ARG(out16 == nullptr && out != nullptr, "the fp16 feature copy needs an even hidden multiplicity");
  int grid = grid_for((int64_t)nodes * n, 256);
  if (gate)
    tp_combine_kernel<true, __half><<<grid, 256, 0, (cudaStream_t)stream>>>(yh, node_attr, nodes, n, bias, residual,
                                                                           bn_mul, bn_add, out);
  else
    tp_combine_kernel<false, __half><<<grid, 256, 0, (cudaStream_t)stream>>>(yh, node_attr, nodes, n, bias, residual,
                                                                            bn_mul, bn_add, out);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_tp_combine_y16(const void* y, const float* node_attr, int nodes, int n, int gate, const float* bias,
                         const float* residual, const float* bn_mul, const float* bn_add, float* out,
                         segnn_stream_t stream) {
  SEGNN_CHECK_ARG(out != nullptr, "null pointer");
  return tp_combine_y16_launch(y, node_attr, nodes, n, gate, bias, residual, bn_mul, bn_add, out, nullptr, stream);
}

int segnn_tp_combine_y16_x16(const void* y, const float* node_attr, int nodes, int n, int gate, const float* bias,
                             const float* residual, const float* bn_mul, const float* bn_add, float* out,
                             void* out16, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(out16 != nullptr, "null pointer");
  return tp_combine_y16_launch(y, node_attr, nodes, n, gate, bias, residual, bn_mul, bn_add, out,
                               reinterpret_cast<__half*>(out16), stream);
}

int segnn_head_fwd(const float* h, const float* node_attr, const float* w_head, int nodes, int n, float* pred,
                   segnn_stream_t stream) {
  SEGNN_CHECK_ARG(nodes >= 0 && n >= 1, "bad sizes");
  if (nodes == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(h && node_attr && w_head && pred, "null pointer");
  head_kernel<<<grid_for((int64_t)nodes * 32, 256), 256, 0, (cudaStream_t)stream>>>(h, node_attr, w_head, nodes, n,
                                                                                  pred);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

__global__ void counter_add_kernel(int* ctr, int delta) { *ctr += delta; }

int segnn_counter_add(int* counter, int delta, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(counter != nullptr, "null counter");
  counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, delta);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

int segnn_integrate(const float* pred, float* pos, float* vel, int nodes, float* traj_pos, float* traj_vel,
                    const int* frame, int max_frames, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(nodes >= 0, "bad sizes");
  if (nodes == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(pred && pos && vel, "null pointer");
  integrate_kernel<<<grid_for((int64_t)nodes * 3, 256), 256, 0, (cudaStream_t)stream>>>(pred, pos, vel, nodes,
                                                                                      traj_pos, traj_vel, frame,
                                                                                      max_frames);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

}  // extern "C"
