// K3^T (fp32): backward of the fused SEGNN edge layer. Nothing per-edge is ever stored by the forward pass; every
// block of edges is recomputed here (message_layer_1 combine + gate, message_layer_2 contraction + gate) and then
// differentiated in place.
//
// Two passes over the same edge set, one thread group per *stationary* node, 8 edges per staged block:
//   pass 0  stationary = receiver i, streamed = senders j:  dP_i, d(message_layer_2 weights, bias), d(w_edge1)
//   pass 1  stationary = sender j,  streamed = receivers i: dQ_j
// message_layer_1's pre-activation depends on P_i + Q_j only, so both passes share one code path ("st" + "ot").
// The gradient that reaches every message of receiver i is  dm = A * dagg_i + B * m + C  (per channel), which covers
// the plain sum (A = 1), eval BatchNorm (A = mul) and train-mode BatchNorm (batch-statistics terms B, C computed on
// the host from node-level reductions).
// Weight gradients of message_layer_2 are reduced WITHOUT atomics: the grid is persistent (a CTA walks over several
// sets of stationary nodes), every thread group owns a private slab [6 n^2 + 2 n] of partial sums in a caller-provided
// workspace that it updates with plain read-modify-writes in program order, and a second kernel adds the slabs in a
// fixed order.  Gradients are therefore bit-identical from run to run.
#include "segnn_common.cuh"

namespace segnn {

constexpr int kBE = 8;  // edges per staged block
constexpr int kBG = 4;  // stationary nodes (thread groups) per CTA

__device__ __forceinline__ void bwd_group_barrier(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ float silu_gate_grad_e(float x) {
  const float s = sigmoid_acc(x);
  return kCSilu * s * (1.0f + x * (1.0f - s));
}
__device__ __forceinline__ float sig_gate_grad_e(float x) {
  const float s = sigmoid_acc(x);
  return kCSig * s * (1.0f - s);
}

struct EdgeBwdArgs {
  const float *pos, *mass, *pp, *qq, *w_edge1;
  const float *w2_ss, *w2_vs, *w2_sv, *w2_vv, *b2;      // [u][w]
  const float *w2t_ss, *w2t_vs, *w2t_sv, *w2t_vv;       // transposed [w][u]
  const float *bnA, *bnB, *bnC;                         // [2n], [2n], [n]
  const float* dagg;                                    // [nodes][4][n]
  float* dout;                                          // dP (pass 0) or dQ (pass 1): [nodes][4][3n]
  float* slabs;                                         // pass 0: [gridDim.x * kBG][6 n^2 + 2 n] partial weight grads
  float* dwe_partial;                                   // pass 0: [nodes][6n]
  int nodes, N, n;
};

// WG (pass 0 only): 2 = dP and the weight gradients in one sweep (large graphs: one recompute), 0 = dP / d(w_edge1)
// only, 1 = weight gradients only.  Training-size graphs launch (0) on the main stream and (1) next to the dQ pass on
// side streams: the slab read-modify-writes leave the critical dgrad chain.
template <int NT, int PASS, int WG>
__global__ void __launch_bounds__(NT* kBG) edge_layer_bwd_kernel(const EdgeBwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int w = threadIdx.x, q = threadIdx.y;
  const int n = a.n, N = a.N;
  const int NP = (n + 3) & ~3;
  float* hb = smem + (size_t)q * (kBE * 11 * NP);  // [8][5][NP] edge features
  float* db = hb + kBE * 5 * NP;                   // [8][6][NP] gradients of the message_layer_2 pre-activations
  float* gb = smem + (size_t)kBG * (kBE * 11 * NP) + q * (kBE * 8);  // [8][8]: ax, ay, az, valid, len, mm

  const int64_t slab_stride = (int64_t)6 * n * n + 2 * n;
  constexpr bool kWeights = PASS == 0 && WG != 0, kData = WG != 1;
  float* slab = kWeights ? a.slabs + ((int64_t)blockIdx.x * kBG + q) * slab_stride : nullptr;
  float* s_ss = slab;                       // [n][2n]
  float* s_vs = slab + (int64_t)2 * n * n;  // [n][2n]
  float* s_sv = slab + (int64_t)4 * n * n;  // [n][n]
  float* s_vv = slab + (int64_t)5 * n * n;  // [n][n]
  float* s_b = slab + (int64_t)6 * n * n;   // [2n]
  bool first_block = true, first_node = true;  // the first visit of a slab location stores, later visits accumulate
  // persistent grid: this thread group handles stationary nodes r = (blockIdx.x + k gridDim.x) * kBG + q in order
  for (int64_t r = (int64_t)blockIdx.x * kBG + q; r < a.nodes; r += (int64_t)gridDim.x * kBG) {
  const int64_t g = r / N;
  const int ir = (int)(r - g * N);
  const int64_t base = g * N;
  const bool act = w < n;
  const int n3 = 3 * n;

  if (w >= n && w < NP) {
    for (int t = 0; t < kBE * 11; ++t) hb[t * NP + w] = 0.f;
  }
  const float prx = a.pos[r * 3 + 0], pry = a.pos[r * 3 + 1], prz = a.pos[r * 3 + 2];
  const float mr = a.mass[r];

  // stationary-side projections st[plane][3]: plane 0 = (0s, 0g, 1), planes 1..3 = (0s_k, 0g_k, 1_k)
  float st[4][3], dst[4][3];
  float wd0s = 0.f, wd0g = 0.f, wm0s = 0.f, wm0g = 0.f, wd1 = 0.f, wm1 = 0.f, b2s = 0.f, b2g = 0.f;
  float As = 0.f, Av = 0.f, Bs = 0.f, Bv = 0.f, Cs = 0.f;
  float G[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int k = 0; k < 3; ++k) st[c][k] = dst[c][k] = 0.f;
  if (act) {
    const float* sr = (PASS == 0 ? a.pp : a.qq) + r * 4 * n3;
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int k = 0; k < 3; ++k) st[c][k] = sr[c * n3 + k * n + w];
    wd0s = a.w_edge1[w];
    wd0g = a.w_edge1[n + w];
    wm0s = a.w_edge1[2 * n + w];
    wm0g = a.w_edge1[3 * n + w];
    wd1 = a.w_edge1[4 * n + w];
    wm1 = a.w_edge1[5 * n + w];
    b2s = a.b2[w];
    b2g = a.b2[n + w];
    As = a.bnA[w];
    Av = a.bnA[n + w];
    Bs = a.bnB[w];
    Bv = a.bnB[n + w];
    Cs = a.bnC[w];
    if (PASS == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) G[c] = a.dagg[r * 4 * n + c * n + w];
    }
  }
  float dwe[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float db2s = 0.f, db2g = 0.f;

  for (int o0 = 0; o0 < N; o0 += kBE) {
    // ---- phase 1: recompute message_layer_1 combine + gate for the block's edges -------------------------------
#pragma unroll 2
    for (int e = 0; e < kBE; ++e) {
      const int oo = o0 + e;
      const int64_t on = base + (oo < N ? oo : N - 1);
      // rel_pos = pos[sender] - pos[receiver]
      const float sgn = PASS == 0 ? 1.0f : -1.0f;
      float ux, uy, uz, len;
      unit_vec(sgn * (a.pos[on * 3 + 0] - prx), sgn * (a.pos[on * 3 + 1] - pry), sgn * (a.pos[on * 3 + 2] - prz), ux,
               uy, uz, len);
      const float ax = kY1 * ux, ay = kY1 * uy, az = kY1 * uz;
      const float mm = a.mass[on] * mr;
      if (act) {
        const float* orow = (PASS == 0 ? a.qq : a.pp) + on * 4 * n3;
        float S[4][3];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int k = 0; k < 3; ++k) S[c][k] = st[c][k] + orow[c * n3 + k * n + w];
        const float zs = S[0][0] + ax * S[1][0] + ay * S[2][0] + az * S[3][0] + len * wd0s + mm * wm0s;
        const float zg = S[0][1] + ax * S[1][1] + ay * S[2][1] + az * S[3][1] + len * wd0g + mm * wm0g;
        const float t = S[0][2] + len * wd1 + mm * wm1;
        const float gg = sig_gate(zg);
        const float vx = gg * (ax * t + S[1][2]), vy = gg * (ay * t + S[2][2]), vz = gg * (az * t + S[3][2]);
        hb[(e * 5 + 0) * NP + w] = silu_gate(zs);
        hb[(e * 5 + 1) * NP + w] = ax * vx + ay * vy + az * vz;
        hb[(e * 5 + 2) * NP + w] = vx;
        hb[(e * 5 + 3) * NP + w] = vy;
        hb[(e * 5 + 4) * NP + w] = vz;
      }
      if (w == e) {
        gb[e * 8 + 0] = ax;
        gb[e * 8 + 1] = ay;
        gb[e * 8 + 2] = az;
        gb[e * 8 + 3] = (oo < N && oo != ir) ? 1.0f : 0.0f;
        gb[e * 8 + 4] = len;
        gb[e * 8 + 5] = mm;
      }
    }
    bwd_group_barrier(1 + q, NT);

    float acc[kBE][6];
    if (act) {
      // ---- phase 2: message_layer_2 forward contraction ------------------------------------------------------
#pragma unroll
      for (int e = 0; e < kBE; ++e)
#pragma unroll
        for (int c = 0; c < 6; ++c) acc[e][c] = 0.f;
      for (int u0 = 0; u0 < n; u0 += 4) {
        float wss[4], wsg[4], wds[4], wdg[4], w1[4], w2[4];
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) {
          const int u = u0 + uu;
          const bool ok = u < n;
          wss[uu] = ok ? a.w2_ss[(int64_t)u * 2 * n + w] : 0.f;
          wsg[uu] = ok ? a.w2_ss[(int64_t)u * 2 * n + n + w] : 0.f;
          wds[uu] = ok ? a.w2_vs[(int64_t)u * 2 * n + w] : 0.f;
          wdg[uu] = ok ? a.w2_vs[(int64_t)u * 2 * n + n + w] : 0.f;
          w1[uu] = ok ? a.w2_sv[(int64_t)u * n + w] : 0.f;
          w2[uu] = ok ? a.w2_vv[(int64_t)u * n + w] : 0.f;
        }
#pragma unroll
        for (int e = 0; e < kBE; ++e) {
          const float4 hs = *reinterpret_cast<const float4*>(&hb[(e * 5 + 0) * NP + u0]);
          const float4 hd = *reinterpret_cast<const float4*>(&hb[(e * 5 + 1) * NP + u0]);
          const float4 hx = *reinterpret_cast<const float4*>(&hb[(e * 5 + 2) * NP + u0]);
          const float4 hy = *reinterpret_cast<const float4*>(&hb[(e * 5 + 3) * NP + u0]);
          const float4 hz = *reinterpret_cast<const float4*>(&hb[(e * 5 + 4) * NP + u0]);
          const float s4[4] = {hs.x, hs.y, hs.z, hs.w}, d4[4] = {hd.x, hd.y, hd.z, hd.w};
          const float x4[4] = {hx.x, hx.y, hx.z, hx.w}, y4[4] = {hy.x, hy.y, hy.z, hy.w};
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
      // ---- phase 3: gate backward -> gradients of the six pre-activations (kept in acc and staged in db) --------
#pragma unroll
      for (int e = 0; e < kBE; ++e) {
        const float4 ge = *reinterpret_cast<const float4*>(&gb[e * 8]);
        float Ge[4] = {G[0], G[1], G[2], G[3]};
        if (PASS == 1) {
          const int oo = o0 + e;
          const int64_t on = base + (oo < N ? oo : N - 1);
#pragma unroll
          for (int c = 0; c < 4; ++c) Ge[c] = a.dagg[on * 4 * n + c * n + w];
        }
        const float ys = acc[e][0] + b2s, yg = acc[e][1] + b2g, t1 = acc[e][2];
        const float ms = silu_gate(ys), gt = sig_gate(yg);
        const float ux_ = fmaf(ge.x, t1, acc[e][3]), uy_ = fmaf(ge.y, t1, acc[e][4]), uz_ = fmaf(ge.z, t1, acc[e][5]);
        const float valid = ge.w;
        const float dms = valid * (As * Ge[0] + Bs * ms + Cs);
        const float dmx = valid * (Av * Ge[1] + Bv * gt * ux_);
        const float dmy = valid * (Av * Ge[2] + Bv * gt * uy_);
        const float dmz = valid * (Av * Ge[3] + Bv * gt * uz_);
        const float dys = dms * silu_gate_grad_e(ys);
        const float dyg = sig_gate_grad_e(yg) * (dmx * ux_ + dmy * uy_ + dmz * uz_);
        const float dux = gt * dmx, duy = gt * dmy, duz = gt * dmz;
        acc[e][0] = dys;
        acc[e][1] = dyg;
        acc[e][2] = ge.x * dux + ge.y * duy + ge.z * duz;
        acc[e][3] = dux;
        acc[e][4] = duy;
        acc[e][5] = duz;
#pragma unroll
        for (int c = 0; c < 6; ++c) db[(e * 6 + c) * NP + w] = acc[e][c];
        if (kWeights) {
          db2s += dys;
          db2g += dyg;
        }
      }
    }
    bwd_group_barrier(1 + q, NT);

    if (act) {
      // ---- phase 3b (pass 0): message_layer_2 weight gradients, output column w ---------------------------------
      if (kWeights) {
        // the slab's current partial sums of rows u0 .. u0 + 3 are fetched one iteration ahead (software pipeline), so
        // their L2 latency is covered by the 256 FMAs of the previous iteration; zero on the first visit of the slab
        float nxt[24];
        auto fetch = [&](int u0) {
#pragma unroll
          for (int uu = 0; uu < 4; ++uu) {
            const int u = u0 + uu;
            const bool ld = !first_block && u < n;
            nxt[uu * 6 + 0] = ld ? s_ss[(int64_t)u * 2 * n + w] : 0.f;
            nxt[uu * 6 + 1] = ld ? s_ss[(int64_t)u * 2 * n + n + w] : 0.f;
            nxt[uu * 6 + 2] = ld ? s_vs[(int64_t)u * 2 * n + w] : 0.f;
            nxt[uu * 6 + 3] = ld ? s_vs[(int64_t)u * 2 * n + n + w] : 0.f;
            nxt[uu * 6 + 4] = ld ? s_sv[(int64_t)u * n + w] : 0.f;
            nxt[uu * 6 + 5] = ld ? s_vv[(int64_t)u * n + w] : 0.f;
          }
        };
        fetch(0);
        for (int u0 = 0; u0 < n; u0 += 4) {
          float cur[24];
#pragma unroll
          for (int k = 0; k < 24; ++k) cur[k] = nxt[k];
          if (u0 + 4 < n) fetch(u0 + 4);
          float g_ss_s[4] = {0.f, 0.f, 0.f, 0.f}, g_ss_g[4] = {0.f, 0.f, 0.f, 0.f}, g_vs_s[4] = {0.f, 0.f, 0.f, 0.f},
                g_vs_g[4] = {0.f, 0.f, 0.f, 0.f}, g_sv[4] = {0.f, 0.f, 0.f, 0.f}, g_vv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int e = 0; e < kBE; ++e) {
            const float4 hs = *reinterpret_cast<const float4*>(&hb[(e * 5 + 0) * NP + u0]);
            const float4 hd = *reinterpret_cast<const float4*>(&hb[(e * 5 + 1) * NP + u0]);
            const float4 hx = *reinterpret_cast<const float4*>(&hb[(e * 5 + 2) * NP + u0]);
            const float4 hy = *reinterpret_cast<const float4*>(&hb[(e * 5 + 3) * NP + u0]);
            const float4 hz = *reinterpret_cast<const float4*>(&hb[(e * 5 + 4) * NP + u0]);
            const float s4[4] = {hs.x, hs.y, hs.z, hs.w}, d4[4] = {hd.x, hd.y, hd.z, hd.w};
            const float x4[4] = {hx.x, hx.y, hx.z, hx.w}, y4[4] = {hy.x, hy.y, hy.z, hy.w};
            const float z4[4] = {hz.x, hz.y, hz.z, hz.w};
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) {
              g_ss_s[uu] = fmaf(s4[uu], acc[e][0], g_ss_s[uu]);
              g_ss_g[uu] = fmaf(s4[uu], acc[e][1], g_ss_g[uu]);
              g_vs_s[uu] = fmaf(d4[uu], acc[e][0], g_vs_s[uu]);
              g_vs_g[uu] = fmaf(d4[uu], acc[e][1], g_vs_g[uu]);
              g_sv[uu] = fmaf(s4[uu], acc[e][2], g_sv[uu]);
              g_vv[uu] = fmaf(x4[uu], acc[e][3], g_vv[uu]);
              g_vv[uu] = fmaf(y4[uu], acc[e][4], g_vv[uu]);
              g_vv[uu] = fmaf(z4[uu], acc[e][5], g_vv[uu]);
            }
          }
#pragma unroll
          for (int uu = 0; uu < 4; ++uu) {
            const int u = u0 + uu;
            if (u < n) {  // private slab: only this thread ever touches these addresses
              s_ss[(int64_t)u * 2 * n + w] = cur[uu * 6 + 0] + g_ss_s[uu];
              s_ss[(int64_t)u * 2 * n + n + w] = cur[uu * 6 + 1] + g_ss_g[uu];
              s_vs[(int64_t)u * 2 * n + w] = cur[uu * 6 + 2] + g_vs_s[uu];
              s_vs[(int64_t)u * 2 * n + n + w] = cur[uu * 6 + 3] + g_vs_g[uu];
              s_sv[(int64_t)u * n + w] = cur[uu * 6 + 4] + g_sv[uu];
              s_vv[(int64_t)u * n + w] = cur[uu * 6 + 5] + g_vv[uu];
            }
          }
        }
        first_block = false;
      }
      if (kData) {
      // ---- phase 4: gradients of the edge features (input channel = this thread) --------------------------------
      float dx[kBE][5];
#pragma unroll
      for (int e = 0; e < kBE; ++e)
#pragma unroll
        for (int c = 0; c < 5; ++c) dx[e][c] = 0.f;
      for (int v0 = 0; v0 < n; v0 += 4) {
        float tss[4], tsg[4], tds[4], tdg[4], t1[4], t2[4];
#pragma unroll
        for (int vv = 0; vv < 4; ++vv) {
          const int v = v0 + vv;
          const bool ok = v < n;
          tss[vv] = ok ? a.w2t_ss[(int64_t)v * n + w] : 0.f;        // W_ss[u = w][out v]
          tsg[vv] = ok ? a.w2t_ss[(int64_t)(n + v) * n + w] : 0.f;  // W_ss[u = w][out n + v]
          tds[vv] = ok ? a.w2t_vs[(int64_t)v * n + w] : 0.f;
          tdg[vv] = ok ? a.w2t_vs[(int64_t)(n + v) * n + w] : 0.f;
          t1[vv] = ok ? a.w2t_sv[(int64_t)v * n + w] : 0.f;
          t2[vv] = ok ? a.w2t_vv[(int64_t)v * n + w] : 0.f;
        }
#pragma unroll
        for (int e = 0; e < kBE; ++e) {
          const float4 f0 = *reinterpret_cast<const float4*>(&db[(e * 6 + 0) * NP + v0]);
          const float4 f1 = *reinterpret_cast<const float4*>(&db[(e * 6 + 1) * NP + v0]);
          const float4 f2 = *reinterpret_cast<const float4*>(&db[(e * 6 + 2) * NP + v0]);
          const float4 f3 = *reinterpret_cast<const float4*>(&db[(e * 6 + 3) * NP + v0]);
          const float4 f4 = *reinterpret_cast<const float4*>(&db[(e * 6 + 4) * NP + v0]);
          const float4 f5 = *reinterpret_cast<const float4*>(&db[(e * 6 + 5) * NP + v0]);
          const float ys4[4] = {f0.x, f0.y, f0.z, f0.w}, yg4[4] = {f1.x, f1.y, f1.z, f1.w};
          const float tt4[4] = {f2.x, f2.y, f2.z, f2.w}, dx4[4] = {f3.x, f3.y, f3.z, f3.w};
          const float dy4[4] = {f4.x, f4.y, f4.z, f4.w}, dz4[4] = {f5.x, f5.y, f5.z, f5.w};
#pragma unroll
          for (int vv = 0; vv < 4; ++vv) {
            dx[e][0] = fmaf(tss[vv], ys4[vv], dx[e][0]);
            dx[e][0] = fmaf(tsg[vv], yg4[vv], dx[e][0]);
            dx[e][0] = fmaf(t1[vv], tt4[vv], dx[e][0]);
            dx[e][1] = fmaf(tds[vv], ys4[vv], dx[e][1]);
            dx[e][1] = fmaf(tdg[vv], yg4[vv], dx[e][1]);
            dx[e][2] = fmaf(t2[vv], dx4[vv], dx[e][2]);
            dx[e][3] = fmaf(t2[vv], dy4[vv], dx[e][3]);
            dx[e][4] = fmaf(t2[vv], dz4[vv], dx[e][4]);
          }
        }
      }
      // ---- phase 5: gate / combine backward of message_layer_1, accumulate the stationary side -----------------
#pragma unroll 2
      for (int e = 0; e < kBE; ++e) {
        const float4 ge = *reinterpret_cast<const float4*>(&gb[e * 8]);
        const float len = gb[e * 8 + 4], mm = gb[e * 8 + 5];
        const int oo = o0 + e;
        const int64_t on = base + (oo < N ? oo : N - 1);
        const float* orow = (PASS == 0 ? a.qq : a.pp) + on * 4 * n3;
        float S[4][3];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int k = 0; k < 3; ++k) S[c][k] = st[c][k] + orow[c * n3 + k * n + w];
        const float zs = S[0][0] + ge.x * S[1][0] + ge.y * S[2][0] + ge.z * S[3][0] + len * wd0s + mm * wm0s;
        const float zg = S[0][1] + ge.x * S[1][1] + ge.y * S[2][1] + ge.z * S[3][1] + len * wd0g + mm * wm0g;
        const float t = S[0][2] + len * wd1 + mm * wm1;
        const float zx = ge.x * t + S[1][2], zy = ge.y * t + S[2][2], zz = ge.z * t + S[3][2];
        const float gg = sig_gate(zg);
        const float tvx = dx[e][2] + ge.x * dx[e][1], tvy = dx[e][3] + ge.y * dx[e][1], tvz = dx[e][4] + ge.z * dx[e][1];
        const float dzs = ge.w * silu_gate_grad_e(zs) * dx[e][0];
        const float dzg = ge.w * sig_gate_grad_e(zg) * (zx * tvx + zy * tvy + zz * tvz);
        const float dzx = ge.w * gg * tvx, dzy = ge.w * gg * tvy, dzz = ge.w * gg * tvz;
        const float dt = ge.x * dzx + ge.y * dzy + ge.z * dzz;
        dst[0][0] += dzs;
        dst[1][0] += ge.x * dzs;
        dst[2][0] += ge.y * dzs;
        dst[3][0] += ge.z * dzs;
        dst[0][1] += dzg;
        dst[1][1] += ge.x * dzg;
        dst[2][1] += ge.y * dzg;
        dst[3][1] += ge.z * dzg;
        dst[0][2] += dt;
        dst[1][2] += dzx;
        dst[2][2] += dzy;
        dst[3][2] += dzz;
        if (PASS == 0) {
          dwe[0] += len * dzs;
          dwe[1] += len * dzg;
          dwe[2] += mm * dzs;
          dwe[3] += mm * dzg;
          dwe[4] += len * dt;
          dwe[5] += mm * dt;
        }
      }
      }  // kData
    }
    bwd_group_barrier(1 + q, NT);
  }

  if (act) {
    if (kData) {
      float* o = a.dout + r * 4 * n3;
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int k = 0; k < 3; ++k) o[c * n3 + k * n + w] = dst[c][k];
      if (PASS == 0) {
        float* pw = a.dwe_partial + r * 6 * n;
#pragma unroll
        for (int c = 0; c < 6; ++c) pw[c * n + w] = dwe[c];
      }
    }
    if (kWeights) {
      s_b[w] = first_node ? db2s : s_b[w] + db2s;
      s_b[n + w] = first_node ? db2g : s_b[n + w] + db2g;
      first_node = false;
    }
  }
  bwd_group_barrier(1 + q, NT);  // the staging buffers are reused by the next stationary node
  }  // persistent loop
}

// Fixed-order sum of the slabs, all five gradient blocks in one launch: element i of a slab goes to the output block
// whose [offset, offset + count) contains i.  Eight independent partial sums keep eight loads in flight per thread;
// their association is fixed, so the result is bit-identical from run to run.
struct SlabSegs {
  int64_t off[6];
  float* out[5];
};
__global__ void slab_reduce_kernel(const float* __restrict__ slabs, int n_slabs, int64_t stride, SlabSegs segs) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < stride; i += (int64_t)gridDim.x * blockDim.x) {
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float* p = slabs + i;
    int sidx = 0;
    for (; sidx + 8 <= n_slabs; sidx += 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] += p[(int64_t)(sidx + k) * stride];
    }
    for (int k = 0; sidx < n_slabs; ++sidx, ++k) a[k] += p[(int64_t)sidx * stride];
    const float acc = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    int seg = 0;
#pragma unroll
    for (int k = 1; k < 5; ++k) seg += i >= segs.off[k];
    segs.out[seg][i - segs.off[seg]] = acc;
  }
}

static inline unsigned bwd_grid(int64_t nodes) {
  const int64_t want = (nodes + kBG - 1) / kBG;
  return (unsigned)(want < 2 * 148 ? want : 2 * 148);
}

template <int NT, int PASS, int WG>
static int launch_bwd(const EdgeBwdArgs& args, cudaStream_t stream) {
  const int NP = (args.n + 3) & ~3;
  const size_t smem = sizeof(float) * ((size_t)kBG * kBE * 11 * NP + kBG * kBE * 8);
  auto kern = edge_layer_bwd_kernel<NT, PASS, WG>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) {
    set_error("edge_layer_bwd: cudaFuncSetAttribute(%zu bytes): %s", smem, cudaGetErrorString(err));
    return SEGNN_E_CUDA;
  }
  dim3 block(NT, kBG);
  const unsigned grid = bwd_grid(args.nodes);
  kern<<<grid, block, smem, stream>>>(args);
  err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error("edge_layer_bwd: launch: %s", cudaGetErrorString(err));
    return SEGNN_E_CUDA;
  }
  return SEGNN_OK;
}

}  // namespace segnn

using namespace segnn;

extern "C" int segnn_edge_layer_bwd(int pass, const float* pos, const float* mass, int B, int N, int n, const float* p,
                                    const float* q, const float* w_edge1, const float* w2_ss, const float* w2_vs,
                                    const float* w2_sv, const float* w2_vv, const float* b2, const float* w2t_ss,
                                    const float* w2t_vs, const float* w2t_sv, const float* w2t_vv, const float* bn_a,
                                    const float* bn_b, const float* bn_c, const float* dagg, float* dout,
                                    float* dw2_ss, float* dw2_vs, float* dw2_sv, float* dw2_vv, float* db2,
                                    float* dwe_partial, float* workspace, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(pass >= 0 && pass <= 3,
                  "pass must be 0 (dP + weight gradients), 1 (dQ), 2 (dP only) or 3 (weight gradients only)");
  SEGNN_CHECK_ARG(B >= 0 && N >= 1 && n >= 1, "bad sizes");
  if (B == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(pos && mass && p && q && w_edge1 && w2_ss && w2_vs && w2_sv && w2_vv && b2 && w2t_ss && w2t_vs &&
                      w2t_sv && w2t_vv && bn_a && bn_b && bn_c && dagg && (dout || pass == 3),
                  "null pointer");
  SEGNN_CHECK_ARG((pass != 0 && pass != 3) || (dw2_ss && dw2_vs && dw2_sv && dw2_vv && db2 && workspace),
                  "passes 0 and 3 need the weight-gradient outputs and the workspace");
  SEGNN_CHECK_ARG((pass != 0 && pass != 2) || dwe_partial, "passes 0 and 2 need dwe_partial");
  const int64_t nodes64 = (int64_t)B * N;
  SEGNN_CHECK_ARG(nodes64 <= 0x7fffffff, "too many nodes");
  if (n > 96) {
    set_error("segnn_edge_layer_bwd: hidden multiplicity n=%d > 96 is not built", n);
    return SEGNN_E_UNSUPPORTED;
  }
  EdgeBwdArgs a{pos, mass, p, q, w_edge1, w2_ss, w2_vs, w2_sv, w2_vv, b2, w2t_ss, w2t_vs, w2t_sv, w2t_vv,
                bn_a, bn_b, bn_c, dagg, dout, workspace, dwe_partial, (int)nodes64, N, n};
  cudaStream_t s = (cudaStream_t)stream;
  // slab k belongs to the thread group whose first stationary node is k: only the first min(groups, nodes) slabs are
  // ever written, each location first by a store (no zero-fill needed), and only those are reduced
  const int64_t groups = (int64_t)bwd_grid(nodes64) * kBG;
  const int n_slabs = (int)(groups < nodes64 ? groups : nodes64);
  const int64_t stride = (int64_t)6 * n * n + 2 * n;
  int rc;
#define SEGNN_BWD_CASE(NT_)                                                                       \
  rc = pass == 0 ? launch_bwd<NT_, 0, 2>(a, s) : pass == 1 ? launch_bwd<NT_, 1, 0>(a, s)          \
     : pass == 2 ? launch_bwd<NT_, 0, 0>(a, s) : launch_bwd<NT_, 0, 1>(a, s)
  if (n <= 32) SEGNN_BWD_CASE(32);
  else if (n <= 64) SEGNN_BWD_CASE(64);
  else SEGNN_BWD_CASE(96);
#undef SEGNN_BWD_CASE
  if (rc != SEGNN_OK || pass == 1 || pass == 2) return rc;
  // fixed-order reduction of the per-group slabs into the gradient blocks
  SlabSegs segs;
  const int64_t nn = (int64_t)n * n;
  segs.off[0] = 0; segs.off[1] = 2 * nn; segs.off[2] = 4 * nn; segs.off[3] = 5 * nn; segs.off[4] = 6 * nn;
  segs.off[5] = stride;
  segs.out[0] = dw2_ss; segs.out[1] = dw2_vs; segs.out[2] = dw2_sv; segs.out[3] = dw2_vv; segs.out[4] = db2;
  slab_reduce_kernel<<<(unsigned)((stride + 127) / 128), 128, 0, s>>>(workspace, n_slabs, stride, segs);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

extern "C" int64_t segnn_edge_layer_bwd_workspace(int B, int N, int n) {
  if (B < 0 || N < 1 || n < 1 || n > 96) return -1;
  return (int64_t)sizeof(float) * bwd_grid((int64_t)B * N) * kBG * ((int64_t)6 * n * n + 2 * n);
}
