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

}  // namespace segnn

using namespace segnn;

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
