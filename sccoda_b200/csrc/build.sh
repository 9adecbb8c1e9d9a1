#!/usr/bin/env bash
# Builds libsccoda_b200.so (in-tree, sm_100a only).  Used by __graft_entry__.build().
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${here}/../libsccoda_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"${NVCC}" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
    -Xcompiler -fPIC -shared ${SCCODA_NVCC_EXTRA:-} \
    -o "${out}" "${here}/capi.cu" "${here}/kernels.cu" "${here}/summary.cu" "${here}/peaks.cu"
echo "built ${out}"
