#!/usr/bin/env bash
# Builds libcats_b200.so (the C-ABI library) in-tree for sm_100a.
# -fmad=false: the model spec rounds every * and + separately (docs/MODEL.md section 4); this is
# what makes the kernels bit-exact against the CPU oracle.
# -dlcm=cg: global loads go straight to L2.  The kernel runs with the whole unified L1 carved out as
# shared memory and streams the tail of each economy through L2; a 28 KB L1 shared by 24 warps only thrashes.
# The period kernel's instantiations live in six translation units that compile in parallel.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${CATS_OUT:-${here}/../libcats_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
obj="$(mktemp -d)"
trap 'rm -rf "${obj}"' EXIT
flags=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xptxas -dlcm=cg
       -Xcompiler -fPIC ${CATS_NVCC_EXTRA:-})
pids=()
for tu in cats_kernels cats_inst_debug cats_inst_prod cats_inst_static cats_inst_strict cats_inst_mw; do
    "${NVCC}" "${flags[@]}" -c -o "${obj}/${tu}.o" "${here}/${tu}.cu" > "${obj}/${tu}.log" 2>&1 &
    pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=1; done
cat "${obj}"/*.log
[ "$rc" -eq 0 ] || { echo "nvcc failed" >&2; exit 1; }
"${NVCC}" -gencode arch=compute_100a,code=sm_100a -shared -o "${out}" "${obj}"/*.o
echo "built ${out}"
