#!/usr/bin/env bash
# Builds libadell_b200.so in-tree for sm_100a (B200).  No other architecture is targeted.
# Host code: -mfma turns the composers' fmaf() into the instruction (it was a libm call: 59 call sites, ~1 ms per 3 000
# matrices in adell_affine_compose); -ffp-contract=off keeps every separate multiply / add of the restated fp32 chains unfused.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${OUT:-${here}/../libadell_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"${NVCC}" -shared -Xcompiler -fPIC -O3 -std=c++17 -lineinfo \
  -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -mfma,-ffp-contract=off \
  -Xptxas -v \
  -o "${out}" "${here}/capi.cu" "${here}/gather.cu" "${here}/stats.cu" "${here}/batch.cu" "${here}/resize.cu" "${here}/compose.cu" "$@"
echo "built ${out}"
