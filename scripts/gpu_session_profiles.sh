#!/bin/bash
# round 2: the profiler evidence kept under profiles/ (launch list of the bench command, full
# captures of the dominant kernels) and the sanitizer passes over the new kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --sweep-sizes 256 --powerlaw-rows 0 > gpurun_out/r2_launches_bench.log 2>&1
echo "launch list exit $?"; tail -2 gpurun_out/r2_launches.csv | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:k_pcg_persist -c 1 -f -o gpurun_out/r2_pcg_persist_runs \
    python scripts/pcg_one.py 256 -1 1 > gpurun_out/r2_pcg_persist_runs.log 2>&1; echo "pcg capture exit $?"; tail -1 gpurun_out/r2_pcg_persist_runs.log
ncu --set full --clock-control none --import-source on -k regex:k_spmv_runs -s 4 -c 1 -f -o gpurun_out/r2_spmv_runs_csr \
    python scripts/spmv_bench.py 256 0 6 csr > gpurun_out/r2_spmv_runs_csr.log 2>&1; echo "spmv csr capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:k_spmv_runs -s 4 -c 1 -f -o gpurun_out/r2_spmv_runs_sss \
    python scripts/spmv_bench.py 256 0 6 sss > gpurun_out/r2_spmv_runs_sss.log 2>&1; echo "spmv sss capture exit $?"
scripts/sanitize.sh r2b
