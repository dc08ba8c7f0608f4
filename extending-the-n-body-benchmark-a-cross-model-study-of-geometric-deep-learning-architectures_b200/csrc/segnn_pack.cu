// segnn_pack_weights: the reference's flat e3nn parameters -> kernel operand blocks, on the device, in one launch.
//
// A consumer of the C ABI that is not this repo's Python package holds what a reference checkpoint holds: per
// O3TensorProduct one flat `tp.weight` (the [mul1, 1, mul_out] instruction views concatenated in e3nn's instruction
// order; models/segnn/o3_building_blocks.py:86,96) and one `biases` vector.  For the irreps of the SEGNN path the
// instruction order is fixed: with in1 = n_blocks x (n x0e + n x1o) [+ e x0e], in2 = 1x0e + 1x1o,
// out = n0 x0e + n x1o, block b starts at b (2 n n0 + 2 n^2) and holds
//     ss [n][n0] (0e x 0e -> 0e) | sv [n][n] (0e x 1o -> 1o) | vv [n][n] (1o x 0e -> 1o) | vs [n][n0] (1o x 1o -> 0e)
// followed by the extra-scalar block  ss_add [e][n0] | sv_add [e][n].  Net coupling constants (e3nn path weight x the
// reference's sqrt_k_correction, o3_building_blocks.py:150-162): identities, and 1/sqrt(3) for 1o x 1o -> 0e; for edge
// tensor products the constant Y_0 of the edge attribute is folded into the weights that multiply it.
// The layouts written here are those of packing.py (pack_msg1, pack_msg2, pack_node_tp, pack_embedding, pack_head,
// fold_batchnorm), which stays as the differentiable twin used by the training path.
#include "segnn_common.cuh"

namespace segnn {

struct TpLayout {
  int n, n0, n_blocks;
  __host__ __device__ int64_t block() const { return (int64_t)2 * n * n0 + (int64_t)2 * n * n; }
  __host__ __device__ int64_t ss(int b, int u, int c) const { return b * block() + (int64_t)u * n0 + c; }
  __host__ __device__ int64_t sv(int b, int u, int c) const { return b * block() + (int64_t)n * n0 + (int64_t)u * n + c; }
  __host__ __device__ int64_t vv(int b, int u, int c) const {
    return b * block() + (int64_t)n * n0 + (int64_t)n * n + (int64_t)u * n + c;
  }
  __host__ __device__ int64_t vs(int b, int u, int c) const {
    return b * block() + (int64_t)n * n0 + (int64_t)2 * n * n + (int64_t)u * n0 + c;
  }
  __host__ __device__ int64_t ss_add(int e, int c) const { return n_blocks * block() + (int64_t)e * n0 + c; }
  __host__ __device__ int64_t sv_add(int e, int c, int extra) const {
    return n_blocks * block() + (int64_t)extra * n0 + (int64_t)e * n + c;
  }
};

constexpr float kY0f = 0.28209479177387814f;

__global__ void pack_weights_kernel(int kind, int n, const float* __restrict__ w, const float* __restrict__ bias,
                                    float* __restrict__ out, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float v = 0.f;
    if (kind == SEGNN_PACK_MSG1) {
      const TpLayout L{n, 2 * n, 2};
      const int64_t wsz = (int64_t)n * 6 * n;
      if (i < 2 * wsz) {  // w_s | w_v, each [n][6n] = (P0 2n, P1 n, Q0 2n, Q1 n)
        const bool vec = i >= wsz;
        const int64_t j = vec ? i - wsz : i;
        const int u = (int)(j / (6 * n)), c = (int)(j % (6 * n));
        const int b = c >= 3 * n, cc = c - b * 3 * n;
        if (!vec) v = cc < 2 * n ? kY0f * w[L.ss(b, u, cc)] : w[L.sv(b, u, cc - 2 * n)];
        else v = cc < 2 * n ? kInvSqrt3 * w[L.vs(b, u, cc)] : kY0f * w[L.vv(b, u, cc - 2 * n)];
      } else if (i < 2 * wsz + 2 * n) {
        v = bias[i - 2 * wsz];
      } else {  // w_edge [6n] = (Y0 ss_add[dist] 2n, Y0 ss_add[m_i m_j] 2n, sv_add[dist] n, sv_add[m_i m_j] n)
        const int64_t j = i - 2 * wsz - 2 * n;
        if (j < 4 * n) v = kY0f * w[L.ss_add((int)(j / (2 * n)), (int)(j % (2 * n)))];
        else v = w[L.sv_add((int)((j - 4 * n) / n), (int)((j - 4 * n) % n), 2)];
      }
    } else if (kind == SEGNN_PACK_MSG2) {  // ss [n][2n] | vs [n][2n] | sv [n][n] | vv [n][n] | b [2n]
      const TpLayout L{n, 2 * n, 1};
      const int64_t a = (int64_t)2 * n * n, b1 = (int64_t)n * n;
      if (i < a) v = kY0f * w[L.ss(0, (int)(i / (2 * n)), (int)(i % (2 * n)))];
      else if (i < 2 * a) v = kInvSqrt3 * w[L.vs(0, (int)((i - a) / (2 * n)), (int)((i - a) % (2 * n)))];
      else if (i < 2 * a + b1) v = w[L.sv(0, (int)((i - 2 * a) / n), (int)((i - 2 * a) % n))];
      else if (i < 2 * a + 2 * b1) v = kY0f * w[L.vv(0, (int)((i - 2 * a - b1) / n), (int)((i - 2 * a - b1) % n))];
      else v = bias[i - 2 * a - 2 * b1];
    } else if (kind == SEGNN_PACK_UPDATE1 || kind == SEGNN_PACK_UPDATE2 || kind == SEGNN_PACK_POOL1) {
      // w_s | w_v [n_blocks n][n0 + n] | bias [n0]
      const int nb = kind == SEGNN_PACK_UPDATE1 ? 2 : 1;
      const int n0 = kind == SEGNN_PACK_UPDATE2 ? n : 2 * n;
      const TpLayout L{n, n0, nb};
      const int cols = n0 + n;
      const int64_t wsz = (int64_t)nb * n * cols;
      if (i < 2 * wsz) {
        const bool vec = i >= wsz;
        const int64_t j = vec ? i - wsz : i;
        const int row = (int)(j / cols), c = (int)(j % cols);
        const int b = row / n, u = row % n;
        if (!vec) v = c < n0 ? w[L.ss(b, u, c)] : w[L.sv(b, u, c - n0)];
        else v = c < n0 ? kInvSqrt3 * w[L.vs(b, u, c)] : w[L.vv(b, u, c - n0)];
      } else {
        v = bias[i - 2 * wsz];
      }
    } else if (kind == SEGNN_PACK_EMBED) {  // w [6][n] | bias [n]; rows 2, 3 (2x1o x 1o -> 0e) carry 1/sqrt(3)
      if (i < 6 * n) v = (i >= 2 * n && i < 4 * n ? kInvSqrt3 : 1.0f) * w[i];
      else v = bias[i - 6 * n];
    } else if (kind == SEGNN_PACK_HEAD) {  // w_head [2][n][2]: already in instruction order
      v = w[i];
    }
    out[i] = v;
  }
}

__global__ void fold_batchnorm_kernel(const float* __restrict__ weight, const float* __restrict__ bias,
                                      const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                      int n, float eps, float degree, float* __restrict__ mul, float* __restrict__ add) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * n) {
    const float m = weight[i] * rsqrtf(running_var[i] + eps);
    mul[i] = m;
    if (i < n) add[i] = degree * (bias[i] - running_mean[i] * m);
  }
}

static int64_t pack_size(int kind, int64_t n) {
  switch (kind) {
    case SEGNN_PACK_MSG1: return 12 * n * n + 2 * n + 6 * n;
    case SEGNN_PACK_MSG2: return 6 * n * n + 2 * n;
    case SEGNN_PACK_UPDATE1: return 2 * (2 * n) * (3 * n) + 2 * n;
    case SEGNN_PACK_UPDATE2: return 2 * n * (2 * n) + n;
    case SEGNN_PACK_POOL1: return 2 * n * (3 * n) + 2 * n;
    case SEGNN_PACK_EMBED: return 7 * n;
    case SEGNN_PACK_HEAD: return 4 * n;
    default: return -1;
  }
}

}  // namespace segnn

using namespace segnn;

extern "C" int64_t segnn_pack_weights_size(int kind, int n) {
  if (n < 1) return -1;
  return pack_size(kind, n);
}

extern "C" int segnn_pack_weights(int kind, int n, const float* tp_weight, const float* biases, float* out,
                                  segnn_stream_t stream) {
  const int64_t total = n >= 1 ? pack_size(kind, n) : -1;
  SEGNN_CHECK_ARG(total > 0, "unknown kind or bad multiplicity");
  SEGNN_CHECK_ARG(tp_weight && out && (biases || kind == SEGNN_PACK_HEAD), "null pointer");
  const int64_t blocks = (total + 255) / 256;
  pack_weights_kernel<<<(unsigned)(blocks > 148 * 8 ? 148 * 8 : blocks), 256, 0, (cudaStream_t)stream>>>(
      kind, n, tp_weight, biases, out, total);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}

extern "C" int segnn_fold_batchnorm(const float* weight, const float* bias, const float* running_mean,
                                    const float* running_var, int n, float eps, float degree, float* mul, float* add,
                                    segnn_stream_t stream) {
  SEGNN_CHECK_ARG(n >= 1, "bad sizes");
  SEGNN_CHECK_ARG(weight && bias && running_mean && running_var && mul && add, "null pointer");
  fold_batchnorm_kernel<<<(2 * n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(weight, bias, running_mean, running_var,
                                                                               n, eps, degree, mul, add);
  SEGNN_CHECK_LAUNCH();
  return SEGNN_OK;
}
