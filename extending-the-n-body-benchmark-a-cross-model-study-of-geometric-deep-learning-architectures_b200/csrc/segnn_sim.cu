// Ground-truth N-body simulator on the device: the reference's GravitySim
// (datasets/nbody/dataset/synthetic_sim.py:305-420: softened gravity, kick-drift-kick leapfrog, float64) runs one
// Python/NumPy process per simulation for 10,000 steps; here one CTA integrates one simulation for the whole
// trajectory inside a single launch (state in registers, positions exchanged through shared memory, all-pairs force
// per step) and writes the sampled frames straight into [frames][B*N][3] trajectory buffers.  float64 like the
// reference: the systems are chaotic, and parity with the NumPy integrator over hundreds of steps needs it.
#include "segnn_common.cuh"

namespace segnn {

// acceleration of body i from all bodies in shared memory (synthetic_sim.py:319-342 compute_acceleration)
__device__ __forceinline__ void gravity_acc(const double* __restrict__ sp, const double* __restrict__ sm, int N,
                                            double xi, double yi, double zi, double G, double soft2, double& ax,
                                            double& ay, double& az) {
  double sx = 0.0, sy = 0.0, sz = 0.0;
  for (int j = 0; j < N; ++j) {
    const double dx = sp[3 * j] - xi, dy = sp[3 * j + 1] - yi, dz = sp[3 * j + 2] - zi;
    const double r2 = dx * dx + dy * dy + dz * dz + soft2;
    const double inv = r2 > 0.0 ? 1.0 / (r2 * sqrt(r2)) : 0.0;  // inv_r3[inv_r3 > 0] ** (-1.5)
    const double w = inv * sm[j];
    sx += dx * w;
    sy += dy * w;
    sz += dz * w;
  }
  ax = G * sx;
  ay = G * sy;
  az = G * sz;
}

// one CTA per simulation, one thread per body (N <= blockDim.x)
__global__ void gravity_sim_kernel(double* __restrict__ pos, double* __restrict__ vel, const double* __restrict__ mass,
                                   int B, int N, double G, double soft2, double dt, int steps, int sample_freq,
                                   double* __restrict__ traj_pos, double* __restrict__ traj_vel,
                                   double* __restrict__ traj_force) {
  extern __shared__ double sh[];
  double* sp = sh;          // [N][3]
  double* sm = sh + 3 * N;  // [N]
  const int sim = blockIdx.x, i = threadIdx.x;
  const bool live = i < N;
  const long long node = (long long)sim * N + i;
  const long long nodes = (long long)B * N;
  double x = 0, y = 0, z = 0, vx = 0, vy = 0, vz = 0, m = 0, ax = 0, ay = 0, az = 0;
  if (live) {
    x = pos[node * 3];
    y = pos[node * 3 + 1];
    z = pos[node * 3 + 2];
    vx = vel[node * 3];
    vy = vel[node * 3 + 1];
    vz = vel[node * 3 + 2];
    m = mass[node];
    sp[3 * i] = x;
    sp[3 * i + 1] = y;
    sp[3 * i + 2] = z;
    sm[i] = m;
  }
  __syncthreads();
  if (live) gravity_acc(sp, sm, N, x, y, z, G, soft2, ax, ay, az);  // initial accelerations (:375)
  int frame = 0;
  for (int s = 0; s < steps; ++s) {
    if (s % sample_freq == 0) {  // :396-400: save, then step
      if (live) {
        const long long o = ((long long)frame * nodes + node) * 3;
        traj_pos[o] = x;
        traj_pos[o + 1] = y;
        traj_pos[o + 2] = z;
        traj_vel[o] = vx;
        traj_vel[o + 1] = vy;
        traj_vel[o + 2] = vz;
        if (traj_force != nullptr) {
          traj_force[o] = ax * m;
          traj_force[o + 1] = ay * m;
          traj_force[o + 2] = az * m;
        }
      }
      ++frame;
    }
    // simulate_step (:344-358): half kick, drift, new accelerations, half kick
    vx += ax * dt / 2.0;
    vy += ay * dt / 2.0;
    vz += az * dt / 2.0;
    x += vx * dt;
    y += vy * dt;
    z += vz * dt;
    __syncthreads();  // everyone has read the previous positions
    if (live) {
      sp[3 * i] = x;
      sp[3 * i + 1] = y;
      sp[3 * i + 2] = z;
    }
    __syncthreads();
    if (live) gravity_acc(sp, sm, N, x, y, z, G, soft2, ax, ay, az);
    vx += ax * dt / 2.0;
    vy += ay * dt / 2.0;
    vz += az * dt / 2.0;
  }
  if (live) {  // final state back (lets the caller continue the trajectory, :387-388)
    pos[node * 3] = x;
    pos[node * 3 + 1] = y;
    pos[node * 3 + 2] = z;
    vel[node * 3] = vx;
    vel[node * 3 + 1] = vy;
    vel[node * 3 + 2] = vz;
  }
}

}  // namespace segnn

using namespace segnn;

extern "C" int segnn_sim_gravity(double* pos, double* vel, const double* mass, int B, int N, double G, double softening,
                                 double dt, int steps, int sample_freq, double* traj_pos, double* traj_vel,
                                 double* traj_force, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(B >= 0 && N >= 1 && steps >= 0 && sample_freq >= 1, "bad sizes");
  SEGNN_CHECK_ARG(steps % sample_freq == 0, "steps must be a multiple of sample_freq (synthetic_sim.py:364)");
  if (B == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(pos && vel && mass && traj_pos && traj_vel, "null pointer");
  if (N > 1024) {
    set_error("segnn_sim_gravity: N=%d exceeds one thread per body per CTA (1024)", N);
    return SEGNN_E_UNSUPPORTED;
  }
  const int threads = ((N + 31) / 32) * 32;
  const size_t smem = (size_t)N * 4 * sizeof(double);
  gravity_sim_kernel<<<B, threads, smem, (cudaStream_t)stream>>>(pos, vel, mass, B, N, G, softening * softening, dt, steps,
                                                               sample_freq, traj_pos, traj_vel, traj_force);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}
