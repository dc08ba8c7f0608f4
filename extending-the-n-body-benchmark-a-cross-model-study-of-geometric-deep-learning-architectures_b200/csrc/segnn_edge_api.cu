// C entry points of the fused edge layer (K3): mode dispatch between the FFMA and the tcgen05 kernels.
#include "segnn_common.cuh"

namespace segnn {
int edge_layer_fp32(const float* pos, const float* mass, int B, int N, int n, const float* pp, const float* qq,
                    const float* w_edge1,
                    const float* w2_ss, const float* w2_vs, const float* w2_sv, const float* w2_vv, const float* b2,
                    const float* bn_mul, const float* bn_add, float* agg, float* moments, cudaStream_t stream);
int edge_layer_tc(const float* pos, const float* mass, int B, int N, int n, const float* pp, const float* qq,
                  const float* w_edge1,
                  const float* b2, const void* w2_tc, const float* bn_mul, const float* bn_add, float* agg, int half,
                  cudaStream_t stream);
int edge_layer_h2(const float* pos, const float* mass, int B, int N, int n, const void* pp, const void* qq,
                  const float* w_edge1, const float* b2, const void* w2_tc, const float* bn_mul, const float* bn_add,
                  float* agg, float* moments, void* agg16, cudaStream_t stream);
int64_t pack_w2_tc(const float* ss, const float* vs, const float* sv, const float* vv, int n, int half, void* out,
                   cudaStream_t stream);
}  // namespace segnn

using namespace segnn;

extern "C" {

int segnn_edge_layer_fwd(int mode, const float* pos, const float* mass, int B, int N, int n, const float* p,
                         const float* q, const float* w_edge1, const float* w2_ss, const float* w2_vs, const float* w2_sv,
                         const float* w2_vv, const float* b2, const void* w2_tc, const float* bn_mul,
                         const float* bn_add, float* agg_out, float* moments, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(B >= 0 && N >= 1 && n >= 1, "bad sizes");
  if (B == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(pos && mass && p && q && w_edge1 && b2 && agg_out, "null pointer");
  SEGNN_CHECK_ARG((bn_mul == nullptr) == (bn_add == nullptr), "bn_mul and bn_add must be given together");
  if (mode == SEGNN_MODE_FP32) {
    SEGNN_CHECK_ARG(w2_ss && w2_vs && w2_sv && w2_vv, "fp32 mode needs the four message_layer_2 weight blocks");
    return edge_layer_fp32(pos, mass, B, N, n, p, q, w_edge1, w2_ss, w2_vs, w2_sv, w2_vv, b2, bn_mul, bn_add, agg_out,
                           moments, (cudaStream_t)stream);
  }
  if (mode == SEGNN_MODE_BF16_TC || mode == SEGNN_MODE_FP16_TC) {
    SEGNN_CHECK_ARG(w2_tc != nullptr, "tensor-core mode needs the packed weight image (segnn_pack_w2_tc)");
    SEGNN_CHECK_ARG(moments == nullptr, "tensor-core mode does not emit train-mode moments");
    SEGNN_CHECK_ARG(N >= 2, "tensor-core mode needs N >= 2");
    return edge_layer_tc(pos, mass, B, N, n, p, q, w_edge1, b2, w2_tc, bn_mul, bn_add, agg_out,
                         mode == SEGNN_MODE_FP16_TC ? 1 : 0, (cudaStream_t)stream);
  }
  if (mode == SEGNN_MODE_FP16_PACKED) {
    SEGNN_CHECK_ARG(w2_tc != nullptr, "tensor-core mode needs the packed weight image (segnn_pack_w2_tc, fp16)");
    SEGNN_CHECK_ARG(N >= 2, "tensor-core mode needs N >= 2");
    // the packed-half kernel also emits the train-mode BatchNorm moments (second template instance)
    return edge_layer_h2(pos, mass, B, N, n, p, q, w_edge1, b2, w2_tc, bn_mul, bn_add, agg_out, moments, nullptr,
                         (cudaStream_t)stream);
  }
  set_error("segnn_edge_layer_fwd: unknown mode %d", mode);
  return SEGNN_E_INVALID;
}

int segnn_edge_layer_fwd_out16(const float* pos, const float* mass, int B, int N, int n, const void* p, const void* q,
                               const float* w_edge1, const float* b2, const void* w2_tc, const float* bn_mul,
                               const float* bn_add, void* agg16_out, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(B >= 0 && N >= 2 && n >= 1, "bad sizes");
  if (B == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(pos && mass && p && q && w_edge1 && b2 && w2_tc && agg16_out, "null pointer");
  SEGNN_CHECK_ARG((bn_mul == nullptr) == (bn_add == nullptr), "bn_mul and bn_add must be given together");
  return edge_layer_h2(pos, mass, B, N, n, p, q, w_edge1, b2, w2_tc, bn_mul, bn_add, nullptr, nullptr, agg16_out,
                       (cudaStream_t)stream);
}

int64_t segnn_pack_w2_tc(const float* w2_ss, const float* w2_vs, const float* w2_sv, const float* w2_vv, int n,
                         int operand, void* out, segnn_stream_t stream) {
  if (operand != SEGNN_OPERAND_BF16 && operand != SEGNN_OPERAND_FP16) {
    set_error("segnn_pack_w2_tc: unknown operand format %d", operand);
    return SEGNN_E_INVALID;
  }
  return pack_w2_tc(w2_ss, w2_vs, w2_sv, w2_vv, n, operand == SEGNN_OPERAND_FP16 ? 1 : 0, out, (cudaStream_t)stream);
}

}  // extern "C"
