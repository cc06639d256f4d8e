#!/usr/bin/env bash
# build_variant.sh <name> <nvcc -D flags...>: an experimental library build/lib_<name>.so whose default-size kernels
# (cats_inst_static, cats_inst_mw) are compiled with the given flags; the other translation units come from a cached
# baseline build (build/obj_base).  Development aid for A/B runs (scripts/variants.sh).
set -euo pipefail
name=$1; shift
here="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
src=$here/r_mabm_b200/csrc
NVCC=/usr/local/cuda/bin/nvcc
flags=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xptxas -dlcm=cg -Xcompiler -fPIC)
base=$here/build/obj_base
mkdir -p $base $here/build/obj_$name
for tu in cats_kernels cats_inst_debug cats_inst_prod cats_inst_strict; do
    [ -f $base/$tu.o ] || $NVCC "${flags[@]}" -c -o $base/$tu.o $src/$tu.cu &
done
for tu in cats_inst_static cats_inst_mw; do
    $NVCC "${flags[@]}" "$@" -c -o $here/build/obj_$name/$tu.o $src/$tu.cu &
done
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $here/build/lib_$name.so $base/*.o $here/build/obj_$name/*.o
echo "built build/lib_$name.so"
