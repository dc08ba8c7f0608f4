// Rollout macros on the device: per (frame, simulation) kinetic energy, softened gravitational
// potential energy and total-momentum magnitude of a trajectory that is already resident in HBM (the rollout's
// trajectory buffers), so the acceptance statistics of trainer.py:888-927 (`_compute_nbody_energies`) and
// datasets/nbody/visualization_utils.py:959-960 (momentum) never need the [B, T, N, 3] arrays on the host or the
// reference's Python loops over simulations.  All-pairs structure like K1: positions of one system staged in shared
// memory, fixed-order block reduction (deterministic).
#include "segnn_common.cuh"

namespace segnn {

constexpr int kMacroThreads = 128;

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMacroThreads / 32; ++i) s += red[i];
  return s;
}

// traj_pos, traj_vel [frames][B*N][3]; out [frames][B][3] = (kinetic, potential, |sum_i v_i|); unit masses as in the
// reference ("masses assumed 1", trainer.py:921).
__global__ void __launch_bounds__(kMacroThreads)
    macro_energy_momentum_kernel(const float* __restrict__ traj_pos, const float* __restrict__ traj_vel, int B, int N,
                                 float G, float soft2, float* __restrict__ out) {
  extern __shared__ float sp[];  // [N][3]
  __shared__ float red[kMacroThreads / 32];
  const long long sys = blockIdx.x;  // frame * B + simulation
  const float* p = traj_pos + sys * N * 3;
  const float* v = traj_vel + sys * N * 3;
  for (int i = threadIdx.x; i < N * 3; i += kMacroThreads) sp[i] = p[i];
  __syncthreads();
  float pot = 0.f, kin = 0.f, mx = 0.f, my = 0.f, mz = 0.f;
  for (int i = threadIdx.x; i < N; i += kMacroThreads) {
    const float xi = sp[i * 3], yi = sp[i * 3 + 1], zi = sp[i * 3 + 2];
    float acc = 0.f;
    for (int j = i + 1; j < N; ++j) {
      const float dx = sp[j * 3] - xi, dy = sp[j * 3 + 1] - yi, dz = sp[j * 3 + 2] - zi;
      acc += rsqrtf(dx * dx + dy * dy + dz * dz + soft2);
    }
    pot += acc;
    const float vx = v[i * 3], vy = v[i * 3 + 1], vz = v[i * 3 + 2];
    kin += vx * vx + vy * vy + vz * vz;
    mx += vx;
    my += vy;
    mz += vz;
  }
  pot = block_sum(pot, red);
  kin = block_sum(kin, red);
  mx = block_sum(mx, red);
  my = block_sum(my, red);
  mz = block_sum(mz, red);
  if (threadIdx.x == 0) {
    out[sys * 3 + 0] = 0.5f * kin;
    out[sys * 3 + 1] = -G * pot;
    out[sys * 3 + 2] = sqrtf(mx * mx + my * my + mz * mz);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Event counters of datasets/nbody/visualization_utils.py, one CTA per simulation, frames visited in order:
//   :1093-1124 count_stickings_and_collisions  (pair contact run lengths: collision on first contact, upgraded to a
//               sticking when the run reaches time_threshold)
//   :1145-1167 count_balls_leaving_defined_area (bodies whose final run of steps farther than the threshold from the
//               centre of mass is longer than 10)
//   :1170-1187 get_max_distance_of_com_from_starting_position
//   :1201-1222 count_sharp_turns (angle between consecutive velocities above the threshold)
// The reference runs Python triple loops over (simulation, step, pair); here the per-pair run lengths live in shared
// memory (uint16 [N][N], upper triangle) and every frame is staged once.
// out_counts [B][4] int32 = (stickings, collisions, bodies_left, sharp_turns); out_com [B] float.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    macro_counters_kernel(const float* __restrict__ traj_pos, const float* __restrict__ traj_vel, int frames, int B,
                          int N, int time_threshold, float contact_dist, float leave_dist, float cos_turn,
                          int* __restrict__ out_counts, float* __restrict__ out_com) {
  extern __shared__ unsigned char raw[];
  float* cur = reinterpret_cast<float*>(raw);             // [N][3] positions of the frame
  float* vprev = cur + 3 * N;                             // [N][3] velocities of the previous frame
  int* outside = reinterpret_cast<int*>(vprev + 3 * N);   // [N] current run length outside the area
  unsigned short* contact = reinterpret_cast<unsigned short*>(outside + N);  // [N][N]
  __shared__ float red[8][3];
  __shared__ int ired[8];
  __shared__ float com[3], com0[3];
  const int sim = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long nodes = (long long)B * N;
  for (int i = tid; i < N * N; i += 256) contact[i] = 0;
  for (int i = tid; i < N; i += 256) outside[i] = 0;
  int stick = 0, coll = 0, turns = 0;
  float maxd = 0.f;
  for (int f = 0; f < frames; ++f) {
    const float* p = traj_pos + ((long long)f * nodes + (long long)sim * N) * 3;
    const float* v = traj_vel + ((long long)f * nodes + (long long)sim * N) * 3;
    __syncthreads();
    for (int i = tid; i < 3 * N; i += 256) cur[i] = p[i];
    __syncthreads();
    // centre of mass (equal masses)
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int i = tid; i < N; i += 256) {
      sx += cur[3 * i];
      sy += cur[3 * i + 1];
      sz += cur[3 * i + 2];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, o);
      sy += __shfl_xor_sync(0xffffffffu, sy, o);
      sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    if (lane == 0) {
      red[warp][0] = sx;
      red[warp][1] = sy;
      red[warp][2] = sz;
    }
    __syncthreads();
    if (tid < 3) {
      float s = 0.f;
      for (int w = 0; w < 8; ++w) s += red[w][tid];
      com[tid] = s / (float)N;
      if (f == 0) com0[tid] = s / (float)N;
    }
    __syncthreads();
    if (f > 0) {
      if (tid == 0) {
        const float dx = com[0] - com0[0], dy = com[1] - com0[1], dz = com[2] - com0[2];
        maxd = fmaxf(maxd, sqrtf(dx * dx + dy * dy + dz * dz));
      }
      for (int i = tid; i < N; i += 256) {
        const float dx = cur[3 * i] - com[0], dy = cur[3 * i + 1] - com[1], dz = cur[3 * i + 2] - com[2];
        outside[i] = sqrtf(dx * dx + dy * dy + dz * dz) > leave_dist ? outside[i] + 1 : 0;
        const float ax = v[3 * i], ay = v[3 * i + 1], az = v[3 * i + 2];
        const float bx = vprev[3 * i], by = vprev[3 * i + 1], bz = vprev[3 * i + 2];
        const float c = (ax * bx + ay * by + az * bz) / (sqrtf(ax * ax + ay * ay + az * az) * sqrtf(bx * bx + by * by + bz * bz));
        if (c < cos_turn) ++turns;  // angle > threshold; NaN (a zero velocity) compares false like in numpy
      }
      for (int i = warp; i < N; i += 8) {
        const float xi = cur[3 * i], yi = cur[3 * i + 1], zi = cur[3 * i + 2];
        for (int j = i + 1 + lane; j < N; j += 32) {
          const float dx = xi - cur[3 * j], dy = yi - cur[3 * j + 1], dz = zi - cur[3 * j + 2];
          unsigned short c = contact[i * N + j];
          if (sqrtf(dx * dx + dy * dy + dz * dz) <= contact_dist) {
            c = c == 0xffffu ? c : (unsigned short)(c + 1);
            if (c == 1) ++coll;
            if (c == time_threshold) {
              ++stick;
              --coll;
            }
          } else {
            c = 0;
          }
          contact[i * N + j] = c;
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < 3 * N; i += 256) vprev[i] = v[i];
  }
  __syncthreads();
  int left = 0;
  for (int i = tid; i < N; i += 256) left += outside[i] > 10 ? 1 : 0;
  int vals[4] = {stick, coll, left, turns};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int s = vals[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __syncthreads();
    if (lane == 0) ired[warp] = s;
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < 8; ++w) t += ired[w];
      out_counts[sim * 4 + q] = t;
    }
  }
  if (tid == 0) out_com[sim] = maxd;
}

}  // namespace segnn

using namespace segnn;

extern "C" int segnn_macros_counters(const float* traj_pos, const float* traj_vel, int frames, int B, int N,
                                     int time_threshold, float contact_distance, float leave_distance,
                                     float turn_angle_degrees, int* out_counts, float* out_com, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(frames >= 1 && B >= 0 && N >= 1 && time_threshold >= 1, "bad sizes");
  if (B == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(traj_pos && traj_vel && out_counts && out_com, "null pointer");
  const size_t smem = (size_t)N * 6 * sizeof(float) + (size_t)N * sizeof(int) + (size_t)N * N * sizeof(unsigned short);
  if (smem > 200 * 1024 || frames > 65000) {
    set_error("segnn_macros_counters: N=%d / frames=%d exceed the shared-memory pair table (N <= ~310, frames <= 65000)", N,
              frames);
    return SEGNN_E_UNSUPPORTED;
  }
  cudaError_t err = cudaFuncSetAttribute(macro_counters_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) {
    set_error("segnn_macros_counters: cudaFuncSetAttribute: %s", cudaGetErrorString(err));
    return SEGNN_E_CUDA;
  }
  const float cos_turn = cosf(turn_angle_degrees * 0.017453292519943295f);
  macro_counters_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(traj_pos, traj_vel, frames, B, N, time_threshold,
                                                                contact_distance, leave_distance, cos_turn, out_counts,
                                                                out_com);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

extern "C" int segnn_macros_energy_momentum(const float* traj_pos, const float* traj_vel, int frames, int B, int N,
                                            float G, float softening, float* out, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(frames >= 0 && B >= 0 && N >= 1, "bad sizes");
  if (frames == 0 || B == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(traj_pos && traj_vel && out, "null pointer");
  const size_t smem = (size_t)N * 3 * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("segnn_macros_energy_momentum: N=%d does not fit the shared-memory staging", N);
    return SEGNN_E_UNSUPPORTED;
  }
  const long long systems = (long long)frames * B;
  SEGNN_CHECK_ARG(systems <= 0x7fffffffLL, "too many (frame, simulation) pairs for one launch");
  if (smem > 48 * 1024) {
    cudaError_t err = cudaFuncSetAttribute(macro_energy_momentum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem);
    if (err != cudaSuccess) {
      set_error("segnn_macros_energy_momentum: cudaFuncSetAttribute: %s", cudaGetErrorString(err));
      return SEGNN_E_CUDA;
    }
  }
  macro_energy_momentum_kernel<<<(unsigned)systems, kMacroThreads, smem, (cudaStream_t)stream>>>(
      traj_pos, traj_vel, B, N, G, softening * softening, out);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}
