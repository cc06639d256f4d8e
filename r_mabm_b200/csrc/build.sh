#!/usr/bin/env bash
# Builds libcats_b200.so (the C-ABI library) in-tree for sm_100a.
# -fmad=false: the model spec rounds every * and + separately (docs/MODEL.md section 4); this is
# what makes the kernels bit-exact against the CPU oracle.
# -dlcm=cg: global loads go straight to L2.  The kernel runs with the whole unified L1 carved out as
# shared memory and streams the tail of each economy through L2; a 28 KB L1 shared by 24 warps only thrashes.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${CATS_OUT:-${here}/../libcats_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"${NVCC}" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 \
    -Xptxas -dlcm=cg -Xcompiler -fPIC -shared ${CATS_NVCC_EXTRA:-} \
    -o "${out}" "${here}/cats_kernels.cu"
echo "built ${out}"
