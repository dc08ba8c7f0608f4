// C entry points of the fused edge layer (K3): mode dispatch between the FFMA and the tcgen05 kernels.
#include "segnn_common.cuh"

namespace segnn {
int edge_layer_fp32(const float* pos, const float* mass, int B, int N, int n, const float* pq, const float* w_edge1,
                    const float* w2_ss, const float* w2_vs, const float* w2_sv, const float* w2_vv, const float* b2,
                    const float* bn_mul, const float* bn_add, float* agg, float* moments, cudaStream_t stream);
}  // namespace segnn

using namespace segnn;

extern "C" {

int segnn_edge_layer_fwd(int mode, const float* pos, const float* mass, int B, int N, int n, const float* pq,
                         const float* w_edge1, const float* w2_ss, const float* w2_vs, const float* w2_sv,
                         const float* w2_vv, const float* b2, const void* w2_tc, const float* bn_mul,
                         const float* bn_add, float* agg_out, float* moments, segnn_stream_t stream) {
  SEGNN_CHECK_ARG(B >= 0 && N >= 1 && n >= 1, "bad sizes");
  if (B == 0) return SEGNN_OK;
  SEGNN_CHECK_ARG(pos && mass && pq && w_edge1 && b2 && agg_out, "null pointer");
  SEGNN_CHECK_ARG((bn_mul == nullptr) == (bn_add == nullptr), "bn_mul and bn_add must be given together");
  if (mode == SEGNN_MODE_FP32) {
    SEGNN_CHECK_ARG(w2_ss && w2_vs && w2_sv && w2_vv, "fp32 mode needs the four message_layer_2 weight blocks");
    return edge_layer_fp32(pos, mass, B, N, n, pq, w_edge1, w2_ss, w2_vs, w2_sv, w2_vv, b2, bn_mul, bn_add, agg_out,
                           moments, (cudaStream_t)stream);
  }
  (void)w2_tc;
  set_error("segnn_edge_layer_fwd: mode %d is not built", mode);
  return SEGNN_E_UNSUPPORTED;
}

int64_t segnn_pack_w2_tc(const float* w2_ss, const float* w2_vs, const float* w2_sv, const float* w2_vv, int n,
                         void* out, segnn_stream_t stream) {
  (void)w2_ss; (void)w2_vs; (void)w2_sv; (void)w2_vv; (void)n; (void)out; (void)stream;
  set_error("segnn_pack_w2_tc: tensor-core mode is not built");
  return SEGNN_E_UNSUPPORTED;
}

}  // extern "C"
