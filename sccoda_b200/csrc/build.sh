#!/usr/bin/env bash
# Builds libsccoda_b200.so (in-tree, sm_100a only).  Used by __graft_entry__.build().
# The translation units are compiled in parallel and linked into one shared library.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${SCCODA_LIB_OUT:-${here}/../libsccoda_b200.so}"
obj="${SCCODA_OBJ_DIR:-${here}/build}"
mkdir -p "${obj}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Wno-deprecated-gpu-targets -Xcompiler -fPIC ${SCCODA_NVCC_EXTRA:-})
pids=()
for src in capi kernels kernels_f32 kernels_f64 kernels_f32g kernels_f64g kernels_cc kernels_mg summary predict peaks pack; do
    ext=cu; [ -f "${here}/${src}.cpp" ] && ext=cpp
    "${NVCC}" "${FLAGS[@]}" -c -o "${obj}/${src}.o" "${here}/${src}.${ext}" > "${obj}/${src}.log" 2>&1 &
    pids+=($!)
done
fail=0
for pid in "${pids[@]}"; do wait "${pid}" || fail=1; done
cat "${obj}"/*.log
if [ "${fail}" -ne 0 ]; then echo "nvcc failed" >&2; exit 1; fi
"${NVCC}" -shared -Wno-deprecated-gpu-targets -o "${out}" "${obj}"/capi.o "${obj}"/kernels.o "${obj}"/kernels_f32.o "${obj}"/kernels_f64.o \
    "${obj}"/kernels_f32g.o "${obj}"/kernels_f64g.o "${obj}"/kernels_cc.o "${obj}"/kernels_mg.o \
    "${obj}"/summary.o "${obj}"/predict.o "${obj}"/peaks.o "${obj}"/pack.o -Xcompiler -pthread
echo "built ${out}"
