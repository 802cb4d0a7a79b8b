#!/bin/bash
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_spmv_runs -s 4 -c 1 -f -o gpurun_out/r2_spmv_runs_sss_384 python scripts/spmv_bench.py 384 0 6 sss > gpurun_out/n384_sss.log 2>&1; echo "ncu sss exit $?"; tail -2 gpurun_out/n384_sss.log
timeout 300 python scripts/runs_probe.py --stats 384 > gpurun_out/n384_stats.log 2>&1; tail -5 gpurun_out/n384_stats.log
