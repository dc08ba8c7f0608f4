// Shared helpers for the sm_100a SEGNN kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/segnn_b200.h"

namespace segnn {

// e3nn 'integral' spherical harmonics of unit vectors: Y_0 = 1/(2 sqrt(pi)), Y_1 = sqrt(3/(4 pi)) * r_hat
// (models/segnn/o3_building_blocks.py:243-245).
constexpr float kY0 = 0.28209479177387814f;
constexpr float kY1 = 0.4886025119029199f;
// e3nn normalize2mom constants of SiLU / sigmoid (1e6 float64 samples, seed 0) used by e3nn Gate
// (o3_building_blocks.py:186-193).
constexpr float kCSilu = 1.6791767923989418f;
constexpr float kCSig = 1.8467055342154763f;
constexpr float kInvSqrt3 = 0.5773502691896258f;

void set_error(const char* fmt, ...);

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float silu_gate(float x) { return kCSilu * x * sigmoid_acc(x); }
__device__ __forceinline__ float sig_gate(float x) { return kCSig * sigmoid_acc(x); }

// 1 MUFU each: sigmoid(x) = 0.5 * tanh(0.5 x) + 0.5
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

// unit vector of r with e3nn / torch.nn.functional.normalize semantics: r / max(|r|, 1e-12)
__device__ __forceinline__ void unit_vec(float x, float y, float z, float& ux, float& uy, float& uz, float& len) {
  len = sqrtf(x * x + y * y + z * z);
  float inv = 1.0f / fmaxf(len, 1e-12f);
  ux = x * inv;
  uy = y * inv;
  uz = z * inv;
}

}  // namespace segnn

#define SEGNN_CHECK_ARG(cond, msg)                       \
  do {                                                   \
    if (!(cond)) {                                       \
      segnn::set_error("%s: %s", __func__, msg);         \
      return SEGNN_E_INVALID;                            \
    }                                                    \
  } while (0)

#define SEGNN_CHECK_LAUNCH()                                                     \
  do {                                                                           \
    cudaError_t err__ = cudaGetLastError();                                      \
    if (err__ != cudaSuccess) {                                                  \
      segnn::set_error("%s: CUDA error: %s", __func__, cudaGetErrorString(err__)); \
      return SEGNN_E_CUDA;                                                       \
    }                                                                            \
  } while (0)
