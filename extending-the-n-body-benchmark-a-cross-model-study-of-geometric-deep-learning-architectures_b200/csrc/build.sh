#!/usr/bin/env bash
# Builds libsegnn_b200.so (sm_100a) next to the Python package. No torch dependency: plain C ABI.
# One object per source, compiled in parallel, rebuilt only when the source or a header is newer.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${here}/../libsegnn_b200.so"
obj="${here}/build"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC ${SEGNN_NVCC_EXTRA:-})
SOURCES=(segnn_node.cu segnn_edge_fp32.cu segnn_edge_tc.cu segnn_edge_tc_h2.cu segnn_node_gemm_tc.cu segnn_edge_api.cu
         segnn_train.cu segnn_edge_bwd.cu segnn_macros.cu segnn_generic.cu segnn_sim.cu segnn_extras.cu segnn_pack.cu
         segnn_gemm_tf32x3.cu segnn_edge_gemm.cu segnn_l2_rows.cu)
mkdir -p "${obj}"
flag_stamp="${obj}/.flags"
if [[ ! -f "${flag_stamp}" ]] || [[ "$(cat "${flag_stamp}")" != "${FLAGS[*]}" ]]; then
  rm -f "${obj}"/*.o
  echo "${FLAGS[*]}" > "${flag_stamp}"
fi
newest_header="$(ls -t "${here}"/*.cuh "${here}/../../include/segnn_b200.h" | head -1)"
pids=()
objects=()
for src in "${SOURCES[@]}"; do
  o="${obj}/${src%.cu}.o"
  objects+=("${o}")
  if [[ ! -f "${o}" || "${here}/${src}" -nt "${o}" || "${newest_header}" -nt "${o}" ]]; then
    "${NVCC}" "${FLAGS[@]}" -c "${here}/${src}" -o "${o}" &
    pids+=($!)
  fi
done
for pid in "${pids[@]:-}"; do
  [[ -n "${pid}" ]] && wait "${pid}"
done
"${NVCC}" -shared -gencode arch=compute_100a,code=sm_100a "${objects[@]}" -o "${out}"
echo "built ${out}"
