#!/bin/bash
# A/B builds of the CUDA library into build/variants/libpsp_<name>.so, selected at run time
# with PSP_B200_LIB.
#   scripts/build_variants.sh name "-DPSP_PERSIST_CTAS=3 ..." [name flags]...    persistent PCG kernel only
#   scripts/build_variants.sh --all name "-DPSP_LL_ALLREDUCE" [name flags]...     every source file
# Known switches: PSP_PERSIST_CTAS, PSP_PERSIST_STAGES, PSP_VEC_UNROLL, PSP_PERSIST_PROF (phase
# timing printed by CTA 0), PSP_BLOCK_UNPACK (block-local halo unpack), PSP_LL_ALLREDUCE
# (flag-in-data all-reduce in the per-phase kernels; changes a shared struct: needs --all).
set -e
cd "$(dirname "$0")/.."
python -m pysparse_b200.build >/dev/null
mkdir -p build/variants
all=0
if [ "$1" == "--all" ]; then all=1; shift; fi
NV="nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v -Iinclude"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  if [ $all == 1 ]; then
    objs=""
    for src in pysparse_b200/csrc/*.cu; do
      o=build/variants/$(basename $src .cu)_$name.o
      $NV $flags -c $src -o $o 2> build/variants/$(basename $src .cu)_$name.log
      objs="$objs $o"
    done
    grep -h -A2 "k_pcg_persist\|k_bicgstab_persist" build/variants/psp_persist*_$name.log | grep -i "spill\|registers" | sed "s/^/[$name] /"
  else
    $NV $flags -c pysparse_b200/csrc/psp_persist.cu -o build/variants/persist_$name.o 2> build/variants/persist_$name.log
    grep -A2 "k_pcg_persist" build/variants/persist_$name.log | grep -i "spill\|registers" | sed "s/^/[$name] /"
    objs="$(ls build/obj/*.o | grep -v 'psp_persist.cu.o') build/variants/persist_$name.o"
  fi
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/variants/libpsp_$name.so $objs -ldl
done
