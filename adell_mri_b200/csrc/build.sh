#!/usr/bin/env bash
# Builds libadell_b200.so in-tree for sm_100a (B200).  No other architecture is targeted.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${OUT:-${here}/../libadell_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"${NVCC}" -shared -Xcompiler -fPIC -O3 -std=c++17 -lineinfo \
  -gencode arch=compute_100a,code=sm_100a \
  -Xptxas -v \
  -o "${out}" "${here}/capi.cu" "${here}/gather.cu" "${here}/stats.cu" "${here}/batch.cu" "${here}/resize.cu" "${here}/compose.cu" "$@"
echo "built ${out}"
