#!/bin/bash
# ncu capture of the run-format SpMV kernel (256^3 CSR or SSS): scripts/ncu_runs.sh csr|sss tag
fmt=${1:-csr}; tag=${2:-runs}
ncu --set full --clock-control none --import-source on -k regex:k_spmv_runs -s 4 -c 2 -f -o gpurun_out/${tag}_${fmt} \
    python scripts/spmv_bench.py 256 0 6 $fmt > gpurun_out/${tag}_${fmt}.log 2>&1
tail -3 gpurun_out/${tag}_${fmt}.log
