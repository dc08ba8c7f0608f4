// K3 (fp32 mode): fused SEGNN edge layer over the implicit fully-connected graph, FFMA path.
//
// One thread group (NT = round_up(n, 32) threads) owns one receiver i and streams its graph's senders j in
// blocks of kEB edges.  Thread w of the group is "channel w":
//   phase 1  z = P_i[w] + Q_j[w] (+ geometry terms)  -> gate -> (s', v'.a1, v') of the edge, staged in smem
//   phase 2  message_layer_2 as an FFMA contraction over the staged edge features (weights stream from L1/L2)
//            -> gate -> accumulate the receiver's sum in registers (no atomics, one store per receiver).
// The edge list, edge attributes, gathered x_i/x_j and per-edge messages never touch HBM.
#include "segnn_common.cuh"

namespace segnn {

constexpr int kEB = 8;  // edges per staged block
constexpr int kRB = 4;  // receivers (thread groups) per CTA

__device__ __forceinline__ void group_barrier(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

template <int NT>
__global__ void __launch_bounds__(NT* kRB)
    edge_layer_fp32_kernel(const float* __restrict__ pos, const float* __restrict__ mass, int nodes, int N, int n,
                           const float* __restrict__ pp, const float* __restrict__ qq,
                           const float* __restrict__ w_edge1,
                           const float* __restrict__ w2_ss, const float* __restrict__ w2_vs,
                           const float* __restrict__ w2_sv, const float* __restrict__ w2_vv,
                           const float* __restrict__ b2, const float* __restrict__ bn_mul,
                           const float* __restrict__ bn_add, float* __restrict__ agg, float* __restrict__ moments) {
  extern __shared__ __align__(16) float smem[];
  const int w = threadIdx.x, q = threadIdx.y;
  const int NP = (n + 3) & ~3;
  float* hb = smem + (size_t)q * (kEB * 5 * NP);
  float* gb = smem + (size_t)kRB * (kEB * 5 * NP) + q * (kEB * 4);

  const int64_t r = (int64_t)blockIdx.x * kRB + q;
  if (r >= nodes) return;  // whole group leaves together; barriers are per group
  const int64_t g = r / N;
  const int i = (int)(r - g * N);
  const int64_t base = g * N;
  const bool act = w < n;
  const int n3 = 3 * n;

  if (w >= n && w < NP) {
    for (int t = 0; t < kEB * 5; ++t) hb[t * NP + w] = 0.f;
  }

  const float pix = pos[r * 3 + 0], piy = pos[r * 3 + 1], piz = pos[r * 3 + 2];
  const float mi = mass[r];

  // receiver-side projections (P) and per-channel constants
  float p0s = 0.f, p0g = 0.f, p1 = 0.f, p0sk[3] = {0.f, 0.f, 0.f}, p0gk[3] = {0.f, 0.f, 0.f}, p1k[3] = {0.f, 0.f, 0.f};
  float wd0s = 0.f, wd0g = 0.f, wm0s = 0.f, wm0g = 0.f, wd1 = 0.f, wm1 = 0.f, b2s = 0.f, b2g = 0.f;
  if (act) {
    const float* pr = pp + r * 4 * n3;
    p0s = pr[w];
    p0g = pr[n + w];
    p1 = pr[2 * n + w];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float* prk = pr + (1 + k) * n3;
      p0sk[k] = prk[w];
      p0gk[k] = prk[n + w];
      p1k[k] = prk[2 * n + w];
    }
    wd0s = w_edge1[w];
    wd0g = w_edge1[n + w];
    wm0s = w_edge1[2 * n + w];
    wm0g = w_edge1[3 * n + w];
    wd1 = w_edge1[4 * n + w];
    wm1 = w_edge1[5 * n + w];
    b2s = b2[w];
    b2g = b2[n + w];
  }

  float as = 0.f, av[3] = {0.f, 0.f, 0.f}, m2s = 0.f, m2v = 0.f;

  for (int j0 = 0; j0 < N; j0 += kEB) {
    // ---- phase 1: message_layer_1 (hoisted form) + gate, staged per edge -------------------------------
#pragma unroll 2
    for (int e = 0; e < kEB; ++e) {
      const int jj = j0 + e;
      const int j = jj < N ? jj : N - 1;
      const int64_t s = base + j;
      float ux, uy, uz, len;
      unit_vec(pos[s * 3 + 0] - pix, pos[s * 3 + 1] - piy, pos[s * 3 + 2] - piz, ux, uy, uz, len);
      const float ax = kY1 * ux, ay = kY1 * uy, az = kY1 * uz;
      const float mm = mass[s] * mi;
      if (act) {
        const float* qr = qq + s * 4 * n3;
        const float* q1 = qr + n3;
        const float* q2 = q1 + n3;
        const float* q3 = q2 + n3;
        float zs = p0s + qr[w] + ax * (p0sk[0] + q1[w]) + ay * (p0sk[1] + q2[w]) + az * (p0sk[2] + q3[w]) +
                   len * wd0s + mm * wm0s;
        float zg = p0g + qr[n + w] + ax * (p0gk[0] + q1[n + w]) + ay * (p0gk[1] + q2[n + w]) +
                   az * (p0gk[2] + q3[n + w]) + len * wd0g + mm * wm0g;
        float t = p1 + qr[2 * n + w] + len * wd1 + mm * wm1;
        float zx = ax * t + p1k[0] + q1[2 * n + w];
        float zy = ay * t + p1k[1] + q2[2 * n + w];
        float zz = az * t + p1k[2] + q3[2 * n + w];
        const float sg = silu_gate(zs);
        const float gg = sig_gate(zg);
        zx *= gg;
        zy *= gg;
        zz *= gg;
        hb[(e * 5 + 0) * NP + w] = sg;
        hb[(e * 5 + 1) * NP + w] = ax * zx + ay * zy + az * zz;
        hb[(e * 5 + 2) * NP + w] = zx;
        hb[(e * 5 + 3) * NP + w] = zy;
        hb[(e * 5 + 4) * NP + w] = zz;
      }
      if (w == e) {
        gb[e * 4 + 0] = ax;
        gb[e * 4 + 1] = ay;
        gb[e * 4 + 2] = az;
        gb[e * 4 + 3] = (jj < N && jj != i) ? 1.0f : 0.0f;
      }
    }
    group_barrier(1 + q, NT);

    // ---- phase 2: message_layer_2 contraction + gate + in-register aggregation ----------------------------
    if (act) {
      float acc[kEB][6];
#pragma unroll
      for (int e = 0; e < kEB; ++e)
#pragma unroll
        for (int c = 0; c < 6; ++c) acc[e][c] = 0.f;

      for (int u0 = 0; u0 < n; u0 += 4) {
        float wss[4], wsg[4], wds[4], wdg[4], w1[4], w2[4];
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) {
          const int u = u0 + uu;
          const bool ok = u < n;
          wss[uu] = ok ? w2_ss[(int64_t)u * 2 * n + w] : 0.f;
          wsg[uu] = ok ? w2_ss[(int64_t)u * 2 * n + n + w] : 0.f;
          wds[uu] = ok ? w2_vs[(int64_t)u * 2 * n + w] : 0.f;
          wdg[uu] = ok ? w2_vs[(int64_t)u * 2 * n + n + w] : 0.f;
          w1[uu] = ok ? w2_sv[(int64_t)u * n + w] : 0.f;
          w2[uu] = ok ? w2_vv[(int64_t)u * n + w] : 0.f;
        }
#pragma unroll
        for (int e = 0; e < kEB; ++e) {
          const float4 hs = *reinterpret_cast<const float4*>(&hb[(e * 5 + 0) * NP + u0]);
          const float4 hd = *reinterpret_cast<const float4*>(&hb[(e * 5 + 1) * NP + u0]);
          const float4 hx = *reinterpret_cast<const float4*>(&hb[(e * 5 + 2) * NP + u0]);
          const float4 hy = *reinterpret_cast<const float4*>(&hb[(e * 5 + 3) * NP + u0]);
          const float4 hz = *reinterpret_cast<const float4*>(&hb[(e * 5 + 4) * NP + u0]);
          const float s4[4] = {hs.x, hs.y, hs.z, hs.w};
          const float d4[4] = {hd.x, hd.y, hd.z, hd.w};
          const float x4[4] = {hx.x, hx.y, hx.z, hx.w};
          const float y4[4] = {hy.x, hy.y, hy.z, hy.w};
          const float z4[4] = {hz.x, hz.y, hz.z, hz.w};
#pragma unroll
          for (int uu = 0; uu < 4; ++uu) {
            acc[e][0] = fmaf(wss[uu], s4[uu], acc[e][0]);
            acc[e][0] = fmaf(wds[uu], d4[uu], acc[e][0]);
            acc[e][1] = fmaf(wsg[uu], s4[uu], acc[e][1]);
            acc[e][1] = fmaf(wdg[uu], d4[uu], acc[e][1]);
            acc[e][2] = fmaf(w1[uu], s4[uu], acc[e][2]);
            acc[e][3] = fmaf(w2[uu], x4[uu], acc[e][3]);
            acc[e][4] = fmaf(w2[uu], y4[uu], acc[e][4]);
            acc[e][5] = fmaf(w2[uu], z4[uu], acc[e][5]);
          }
        }
      }
#pragma unroll
      for (int e = 0; e < kEB; ++e) {
        const float4 ge = *reinterpret_cast<const float4*>(&gb[e * 4]);
        const float ms = silu_gate(acc[e][0] + b2s);
        const float gt = sig_gate(acc[e][1] + b2g);
        const float mx = gt * fmaf(ge.x, acc[e][2], acc[e][3]);
        const float my = gt * fmaf(ge.y, acc[e][2], acc[e][4]);
        const float mz = gt * fmaf(ge.z, acc[e][2], acc[e][5]);
        if (ge.w != 0.f) {
          as += ms;
          av[0] += mx;
          av[1] += my;
          av[2] += mz;
          m2s = fmaf(ms, ms, m2s);
          m2v += mx * mx + my * my + mz * mz;
        }
      }
    }
    group_barrier(1 + q, NT);
  }

  if (act) {
    if (moments != nullptr) {
      moments[r * 2 * n + w] = m2s;
      moments[r * 2 * n + n + w] = m2v;
    }
    if (bn_mul != nullptr) {
      const float ms = bn_mul[w], mv = bn_mul[n + w];
      as = fmaf(as, ms, bn_add[w]);
      av[0] *= mv;
      av[1] *= mv;
      av[2] *= mv;
    }
    float* o = agg + r * 4 * n;
    o[w] = as;
    o[n + w] = av[0];
    o[2 * n + w] = av[1];
    o[3 * n + w] = av[2];
  }
}

template <int NT>
static int launch_fp32(const float* pos, const float* mass, int nodes, int N, int n, const float* pp, const float* qq,
                       const float* w_edge1, const float* w2_ss, const float* w2_vs, const float* w2_sv,
                       const float* w2_vv, const float* b2, const float* bn_mul, const float* bn_add, float* agg,
                       float* moments, cudaStream_t stream) {
  const int NP = (n + 3) & ~3;
  const size_t smem = sizeof(float) * ((size_t)kRB * kEB * 5 * NP + kRB * kEB * 4);
  auto kern = edge_layer_fp32_kernel<NT>;
  if (smem > 48 * 1024) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) {
      set_error("edge_layer_fp32: cudaFuncSetAttribute: %s", cudaGetErrorString(err));
      return SEGNN_E_CUDA;
    }
  }
  dim3 block(NT, kRB);
  unsigned grid = (unsigned)(((int64_t)nodes + kRB - 1) / kRB);
  kern<<<grid, block, smem, stream>>>(pos, mass, nodes, N, n, pp, qq, w_edge1, w2_ss, w2_vs, w2_sv, w2_vv, b2, bn_mul,
                                      bn_add, agg, moments);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error("edge_layer_fp32: launch: %s", cudaGetErrorString(err));
    return SEGNN_E_CUDA;
  }
  return SEGNN_OK;
}

int edge_layer_fp32(const float* pos, const float* mass, int B, int N, int n, const float* pp, const float* qq,
                    const float* w_edge1,
                    const float* w2_ss, const float* w2_vs, const float* w2_sv, const float* w2_vv, const float* b2,
                    const float* bn_mul, const float* bn_add, float* agg, float* moments, cudaStream_t stream) {
  const int64_t nodes64 = (int64_t)B * N;
  if (nodes64 > 0x7fffffff) {
    set_error("edge_layer_fp32: too many nodes");
    return SEGNN_E_INVALID;
  }
  const int nodes = (int)nodes64;
#define SEGNN_FP32_CASE(NT_)                                                                                   \
  return launch_fp32<NT_>(pos, mass, nodes, N, n, pp, qq, w_edge1, w2_ss, w2_vs, w2_sv, w2_vv, b2, bn_mul, bn_add, \
                          agg, moments, stream)
  if (n <= 32) SEGNN_FP32_CASE(32);
  if (n <= 64) SEGNN_FP32_CASE(64);
  if (n <= 96) SEGNN_FP32_CASE(96);
  if (n <= 128) SEGNN_FP32_CASE(128);
#undef SEGNN_FP32_CASE
  set_error("edge_layer_fp32: hidden multiplicity n=%d > 128 is not built", n);
  return SEGNN_E_UNSUPPORTED;
}

}  // namespace segnn
