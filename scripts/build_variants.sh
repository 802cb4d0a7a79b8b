#!/bin/bash
# A/B builds of the persistent kernel: build/variants/libpsp_<name>.so, selected at run time
# with PSP_B200_LIB.  usage: scripts/build_variants.sh name "-DPSP_PERSIST_CTAS=3 ..." [name flags]...
set -e
cd "$(dirname "$0")/.."
python -m pysparse_b200.build >/dev/null
mkdir -p build/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v -Iinclude $flags \
    -c pysparse_b200/csrc/psp_persist.cu -o build/variants/persist_$name.o 2> build/variants/persist_$name.log
  grep -A2 "k_pcg_persist" build/variants/persist_$name.log | grep -i "spill\|registers" | sed "s/^/[$name] /"
  objs=$(ls build/obj/*.o | grep -v psp_persist)
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/variants/libpsp_$name.so $objs build/variants/persist_$name.o -ldl
done
