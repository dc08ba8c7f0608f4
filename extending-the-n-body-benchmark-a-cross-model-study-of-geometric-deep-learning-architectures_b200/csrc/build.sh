#!/usr/bin/env bash
# Builds libsegnn_b200.so (sm_100a) next to the Python package. No torch dependency: plain C ABI.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${here}/../libsegnn_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"${NVCC}" -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC -shared ${SEGNN_NVCC_EXTRA:-} \
  "${here}/segnn_node.cu" "${here}/segnn_edge_fp32.cu" "${here}/segnn_edge_tc.cu" "${here}/segnn_edge_tc_h2.cu" "${here}/segnn_node_gemm_tc.cu" "${here}/segnn_edge_api.cu" "${here}/segnn_train.cu" "${here}/segnn_edge_bwd.cu" "${here}/segnn_macros.cu" "${here}/segnn_generic.cu" "${here}/segnn_sim.cu" \
  -o "${out}"
echo "built ${out}"
