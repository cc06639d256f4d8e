#!/usr/bin/env bash
# Builds libcats_b200.so (the C-ABI library) in-tree for sm_100a.
# -fmad=false: the model spec rounds every * and + separately (docs/MODEL.md section 4); this is
# what makes the kernels bit-exact against the CPU oracle.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${CATS_OUT:-${here}/../libcats_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"${NVCC}" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 \
    -Xcompiler -fPIC -shared ${CATS_NVCC_EXTRA:-} \
    -o "${out}" "${here}/cats_kernels.cu"
echo "built ${out}"
