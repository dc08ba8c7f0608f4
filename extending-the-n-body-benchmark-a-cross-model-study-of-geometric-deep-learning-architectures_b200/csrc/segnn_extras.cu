// Callers and data formats either side of the hot path, each a small kernel:
//   * charged-particle ground-truth simulator   datasets/nbody_offline/datagen/system.py:78-123 (+ physical_objects.py:49-57)
//   * group-collision macro                     datasets/nbody/visualization_utils.py:1455-1610
//   * k-nearest-neighbour edge list             utils/build_fully_connected_graph.py:42-80
//   * InstanceNorm                              models/segnn/instance_norm.py:53-129
// None of them is performance critical; they exist so that a user of the reference finds the same functionality on
// the device, checked against vectors produced by running the reference's own code (tests/golden/ref_*.pt).
#include "segnn_common.cuh"

namespace segnn {

// ---------------------------------------------------------------------------------------------------------------------
// Charged system, isolated bodies: F_i = sum_j k q_i q_j (x_i - x_j) / |x_i - x_j|^3, clamped component-wise to
// +-max_force, then v += F dt, x += v dt (semi-implicit Euler).  float64; one CTA per simulation, one thread per body.
// |x_i - x_j|^2 is formed as |x_i|^2 + |x_j|^2 - 2 x_i.x_j like System._l2 (system.py:78-83).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void charged_sim_kernel(double* __restrict__ pos, double* __restrict__ vel, const double* __restrict__ charge,
                                   int B, int N, double strength, double dt, double max_force, int steps,
                                   int sample_freq, double* __restrict__ traj_pos, double* __restrict__ traj_vel) {
  extern __shared__ double sh[];
  double* sp = sh;          // [N][3]
  double* sq = sh + 3 * N;  // [N] charges
  double* sn = sh + 4 * N;  // [N] |x|^2
  const int sim = blockIdx.x, i = threadIdx.x;
  const bool live = i < N;
  const long long node = (long long)sim * N + i;
  const long long nodes = (long long)B * N;
  double x = 0, y = 0, z = 0, vx = 0, vy = 0, vz = 0, q = 0;
  if (live) {
    x = pos[node * 3];
    y = pos[node * 3 + 1];
    z = pos[node * 3 + 2];
    vx = vel[node * 3];
    vy = vel[node * 3 + 1];
    vz = vel[node * 3 + 2];
    q = charge[node];
    sq[i] = q;
  }
  int frame = 0;
  for (int s = 0; s < steps; ++s) {
    __syncthreads();
    if (live) {
      sp[3 * i] = x;
      sp[3 * i + 1] = y;
      sp[3 * i + 2] = z;
      sn[i] = x * x + y * y + z * z;
    }
    __syncthreads();
    if (live) {
      double fx = 0.0, fy = 0.0, fz = 0.0;
      const double ni = sn[i];
      for (int j = 0; j < N; ++j) {
        if (j == i) continue;  // np.fill_diagonal(forces_size, 0)
        const double xj = sp[3 * j], yj = sp[3 * j + 1], zj = sp[3 * j + 2];
        const double d2 = ni + sn[j] - 2.0 * (x * xj + y * yj + z * zj);
        const double size = strength * q * sq[j] / (d2 * sqrt(d2));
        fx += size * (x - xj);
        fy += size * (y - yj);
        fz += size * (z - zj);
      }
      fx = fmin(fmax(fx, -max_force), max_force);
      fy = fmin(fmax(fy, -max_force), max_force);
      fz = fmin(fmax(fz, -max_force), max_force);
      vx += fx * dt;
      vy += fy * dt;
      vz += fz * dt;
      x += vx * dt;
      y += vy * dt;
      z += vz * dt;
      if ((s + 1) % sample_freq == 0) {
        const long long o = ((long long)frame * nodes + node) * 3;
        traj_pos[o] = x;
        traj_pos[o + 1] = y;
        traj_pos[o + 2] = z;
        traj_vel[o] = vx;
        traj_vel[o + 1] = vy;
        traj_vel[o + 2] = vz;
      }
    }
    if ((s + 1) % sample_freq == 0) ++frame;
  }
  if (live) {
    pos[node * 3] = x;
    pos[node * 3 + 1] = y;
    pos[node * 3 + 2] = z;
    vel[node * 3] = vx;
    vel[node * 3 + 1] = vy;
    vel[node * 3 + 2] = vz;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Group collisions.  Per simulation: close[t][i][j] = |x_i - x_j| <= d; a pair / triplet is "stuck" on every maximal
// run of >= time_threshold consecutive steps in which all its members are mutually close; for every (stuck pair
// interval, stuck interval of a disjoint triplet) with overlapping lifetimes the reference counts ONE group collision if
// some body of the pair is close to some body of the triplet at any step >= the start of the overlap.  Maximal runs are
// separated by at least one step, so the overlaps of distinct interval pairs are disjoint segments of
// (pair stuck) & (triplet stuck): count the segments whose start is <= the last step at which the two groups touch.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void gc_close_kernel(const float* __restrict__ traj_pos, int frames, int B, int N, float dist,
                                uint8_t* __restrict__ close) {
  const long long total = (long long)B * frames * N * N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % N);
    const int i = (int)((idx / N) % N);
    const int t = (int)((idx / ((long long)N * N)) % frames);
    const int s = (int)(idx / ((long long)N * N * frames));
    const float* a = traj_pos + (((long long)t * B + s) * N + i) * 3;
    const float* b = traj_pos + (((long long)t * B + s) * N + j) * 3;
    // float64 accumulation of float32 differences: the comparison is then exact for float32-representable input
    const double dx = (double)a[0] - (double)b[0], dy = (double)a[1] - (double)b[1], dz = (double)a[2] - (double)b[2];
    close[idx] = (i != j && sqrt(dx * dx + dy * dy + dz * dz) <= (double)dist) ? 1 : 0;
  }
}

// stuck[s][group][t] for pairs (groups 0 .. P-1, P = N(N-1)/2) followed by triplets; one thread per (s, group)
__device__ __forceinline__ void unrank_pair(int r, int N, int& i, int& j) {
  i = 0;
  while (r >= N - 1 - i) {
    r -= N - 1 - i;
    ++i;
  }
  j = i + 1 + r;
}
__device__ __forceinline__ void unrank_triplet(int r, int N, int& i, int& j, int& k) {
  i = 0;
  for (;;) {
    const int m = N - 1 - i;            // bodies after i
    const int cnt = m * (m - 1) / 2;    // triplets starting with i
    if (r < cnt) break;
    r -= cnt;
    ++i;
  }
  int jj, kk;
  unrank_pair(r, N - 1 - i, jj, kk);
  j = i + 1 + jj;
  k = i + 1 + kk;
}

__global__ void gc_stuck_kernel(const uint8_t* __restrict__ close, int frames, int B, int N, int time_threshold,
                                uint8_t* __restrict__ stuck) {
  const int P = N * (N - 1) / 2, T3 = N * (N - 1) * (N - 2) / 6, groups = P + T3;
  const long long total = (long long)B * groups;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % groups);
    const int s = (int)(idx / groups);
    int i, j, k = -1;
    if (g < P)
      unrank_pair(g, N, i, j);
    else
      unrank_triplet(g - P, N, i, j, k);
    const uint8_t* c = close + (long long)s * frames * N * N;
    uint8_t* out = stuck + idx * frames;
    int run = 0;
    for (int t = 0; t <= frames; ++t) {
      bool on = false;
      if (t < frames) {
        const uint8_t* ct = c + (long long)t * N * N;
        on = ct[i * N + j] != 0;
        if (k >= 0) on = on && ct[i * N + k] != 0 && ct[j * N + k] != 0;
        out[t] = 0;
      }
      if (on) {
        ++run;
      } else {
        if (run >= time_threshold)
          for (int u = t - run; u < t; ++u) out[u] = 1;
        run = 0;
      }
    }
  }
}

__global__ void gc_count_kernel(const uint8_t* __restrict__ close, const uint8_t* __restrict__ stuck, int frames, int B,
                                int N, int* __restrict__ counts) {
  const int P = N * (N - 1) / 2, T3 = N * (N - 1) * (N - 2) / 6, groups = P + T3;
  const long long total = (long long)B * P * T3;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int tr = (int)(idx % T3);
    const int pr = (int)((idx / T3) % P);
    const int s = (int)(idx / ((long long)T3 * P));
    int a, b, i, j, k;
    unrank_pair(pr, N, a, b);
    unrank_triplet(tr, N, i, j, k);
    if (a == i || a == j || a == k || b == i || b == j || b == k) continue;  // not disjoint
    const uint8_t* sp = stuck + ((long long)s * groups + pr) * frames;
    const uint8_t* st = stuck + ((long long)s * groups + P + tr) * frames;
    const uint8_t* c = close + (long long)s * frames * N * N;
    int last_touch = -1;
    for (int t = frames - 1; t >= 0 && last_touch < 0; --t) {
      const uint8_t* ct = c + (long long)t * N * N;
      if (ct[a * N + i] | ct[a * N + j] | ct[a * N + k] | ct[b * N + i] | ct[b * N + j] | ct[b * N + k]) last_touch = t;
    }
    if (last_touch < 0) continue;
    int n = 0;
    bool prev = false;
    for (int t = 0; t <= last_touch; ++t) {
      const bool both = sp[t] && st[t];
      if (both && !prev) ++n;
      prev = both;
    }
    if (n) atomicAdd(counts + s, n);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// kNN edge list: for node i of graph g the k nearest other nodes of the same graph in ascending distance (ties by
// index); edge e = (g N + i) k + r : row 0 = g N + i, row 1 = g N + neighbour_r  (build_fully_connected_graph.py:72-80).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void knn_kernel(const double* __restrict__ loc, int B, int N, int dim, int k, int64_t* __restrict__ edge_index) {
  const long long nodes = (long long)B * N;
  const long long E = nodes * k;
  for (long long node = blockIdx.x * (long long)blockDim.x + threadIdx.x; node < nodes;
       node += (long long)gridDim.x * blockDim.x) {
    const long long g0 = (node / N) * N;
    const int i = (int)(node - g0);
    double last_d = -1.0;
    int last_j = -1;
    for (int r = 0; r < k; ++r) {
      double best = 1.0e300;
      int best_j = -1;
      for (int j = 0; j < N; ++j) {
        if (j == i) continue;
        double d = 0.0;
        for (int c = 0; c < dim; ++c) {
          const double df = loc[node * dim + c] - loc[(g0 + j) * dim + c];
          d += df * df;
        }
        const bool after_last = d > last_d || (d == last_d && j > last_j);
        if (after_last && (d < best || (d == best && j < best_j))) {
          best = d;
          best_j = j;
        }
      }
      edge_index[node * k + r] = node;
      edge_index[E + node * k + r] = g0 + best_j;
      last_d = best;
      last_j = best_j;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// InstanceNorm: per graph and irrep channel u: (l = 0) subtract the graph mean; scale by
// weight[u] / sqrt(mean_nodes(mean_m x[u, m]^2) + eps); (d = 1) add bias[u].   grid = (graphs, irrep blocks)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void instance_norm_kernel(const float* __restrict__ x, const int64_t* __restrict__ ptr, int dim,
                                     const int* __restrict__ blocks, const float* __restrict__ weight,
                                     const float* __restrict__ bias, float eps, float* __restrict__ out) {
  const int g = blockIdx.x, blk = blockIdx.y;
  const int off = blocks[blk * 6], mul = blocks[blk * 6 + 1], d = blocks[blk * 6 + 2], l = blocks[blk * 6 + 3];
  const int woff = blocks[blk * 6 + 4], boff = blocks[blk * 6 + 5];
  const long long r0 = ptr[g], r1 = ptr[g + 1];
  const double rows = (double)(r1 - r0);
  for (int u = threadIdx.x; u < mul; u += blockDim.x) {
    double mean = 0.0;
    if (l == 0) {
      for (long long r = r0; r < r1; ++r) mean += x[r * dim + off + u];
      mean /= rows;
    }
    double sq = 0.0;
    for (long long r = r0; r < r1; ++r)
      for (int m = 0; m < d; ++m) {
        const double v = (double)x[r * dim + off + u * d + m] - mean;
        sq += v * v;
      }
    const double norm = sq / (rows * d);
    const float scale = (float)(1.0 / sqrt(norm + (double)eps)) * (weight ? weight[woff + u] : 1.0f);
    const float b = (bias && d == 1) ? bias[boff + u] : 0.0f;
    for (long long r = r0; r < r1; ++r)
      for (int m = 0; m < d; ++m) {
        const long long c = r * dim + off + u * d + m;
        out[c] = (x[c] - (float)mean) * scale + b;
      }
  }
}

}  // namespace segnn

using namespace segnn;

extern "C" int segnn_sim_charged(double* pos, double* vel, const double* charge, int B, int N,
                                 double interaction_strength, double dt, double max_force, int steps, int sample_freq,
                                 double* traj_pos, double* traj_vel, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(B >= 0 && N >= 1 && steps >= 0 && sample_freq >= 1, "bad sizes");
  SEGNN_CHECK_ARG(steps % sample_freq == 0, "steps must be a multiple of sample_freq");
  if (B == 0 || steps == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(pos && vel && charge && traj_pos && traj_vel, "null pointer");
  if (N > 1024) {
    set_error("segnn_sim_charged: N=%d exceeds one thread per body per CTA (1024)", N);
    return SEGNN_E_UNSUPPORTED;
  }
  const int threads = ((N + 31) / 32) * 32;
  charged_sim_kernel<<<B, threads, (size_t)N * 5 * sizeof(double), (cudaStream_t)stream>>>(
      pos, vel, charge, B, N, interaction_strength, dt, max_force, steps, sample_freq, traj_pos, traj_vel);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

extern "C" int64_t segnn_macros_group_collisions_workspace(int frames, int B, int N) {
  if (frames < 0 || B < 0 || N < 0) return -1;
  const int64_t groups = (int64_t)N * (N - 1) / 2 + (int64_t)N * (N - 1) * (N - 2) / 6;
  return (int64_t)B * frames * ((int64_t)N * N + groups);
}

extern "C" int segnn_macros_group_collisions(const float* traj_pos, int frames, int B, int N, int time_threshold,
                                             float distance_threshold, void* workspace, int* out_counts,
                                             segnn_stream_t stream) {
  SEGNN_CHECK_ARG(frames >= 0 && B >= 0 && N >= 0 && time_threshold >= 1, "bad sizes");
  if (B == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(out_counts, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(out_counts, 0, sizeof(int) * B, st);
  if (frames == 0 || N < 5) return SEGNN_OK;  // a pair and a disjoint triplet need five bodies
  SEGNN_CHECK_ARG(traj_pos && workspace, "null pointer");
  if (N > 64) {
    set_error("segnn_macros_group_collisions: N=%d: the triplet table grows like N^3 (built for N <= 64)", N);
    return SEGNN_E_UNSUPPORTED;
  }
  uint8_t* close = (uint8_t*)workspace;
  uint8_t* stuck = close + (int64_t)B * frames * N * N;
  const int64_t P = (int64_t)N * (N - 1) / 2, T3 = (int64_t)N * (N - 1) * (N - 2) / 6;
  auto blocks = [](int64_t work) { return (unsigned)((work + 255) / 256 > 148 * 16 ? 148 * 16 : (work + 255) / 256); };
  gc_close_kernel<<<blocks((int64_t)B * frames * N * N), 256, 0, st>>>(traj_pos, frames, B, N, distance_threshold, close);
  gc_stuck_kernel<<<blocks(B * (P + T3)), 256, 0, st>>>(close, frames, B, N, time_threshold, stuck);
  gc_count_kernel<<<blocks(B * P * T3), 256, 0, st>>>(close, stuck, frames, B, N, out_counts);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

extern "C" int segnn_knn_edge_index(const double* loc, int B, int N, int dim, int k, int64_t* edge_index,
                                    segnn_stream_t stream) {
  SEGNN_CHECK_ARG(B >= 0 && N >= 1 && dim >= 1, "bad sizes");
  if (k >= N || k < 0) {
    set_error("segnn_knn_edge_index: Graph cannot have more neighbors than there are nodes in simulation - 1");
    return SEGNN_E_INVALID;
  }
  if (B == 0 || k == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(loc && edge_index, "null pointer");
  const long long nodes = (long long)B * N;
  const unsigned grid = (unsigned)((nodes + 127) / 128 > 148 * 8 ? 148 * 8 : (nodes + 127) / 128);
  knn_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(loc, B, N, dim, k, edge_index);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

extern "C" int segnn_instance_norm(const float* x, const int64_t* graph_ptr, int graphs, int dim, const int* blocks,
                                   int n_blocks, const float* weight, const float* bias, float eps, float* out,
                                   segnn_stream_t stream) {
  SEGNN_CHECK_ARG(graphs >= 0 && dim >= 0 && n_blocks >= 0, "bad sizes");
  if (graphs == 0 || n_blocks == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(x && graph_ptr && blocks && out, "null pointer");
  instance_norm_kernel<<<dim3((unsigned)graphs, (unsigned)n_blocks), 128, 0, (cudaStream_t)stream>>>(
      x, graph_ptr, dim, blocks, weight, bias, eps, out);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}
