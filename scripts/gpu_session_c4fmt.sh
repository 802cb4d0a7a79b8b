#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/configs_c4fmt0.jsonl gpurun_out/configs_c4fmt1.jsonl
timeout 300 python scripts/run_configs.py c4fmt0 c4 > gpurun_out/c4_fmt0.log 2>&1; echo "fmt0 exit $?"; cut -c1-330 gpurun_out/configs_c4fmt0.jsonl
PSP_OPTS="spmv_format=1" timeout 300 python scripts/run_configs.py c4fmt1 c4 > gpurun_out/c4_fmt1.log 2>&1; echo "fmt1 exit $?"; cut -c1-330 gpurun_out/configs_c4fmt1.jsonl
PSP_OPTS="spmv_format=1" timeout 600 python -m pytest tests/test_gpu_baseline_sizes.py -q > gpurun_out/c4_pytest_fmt1.log 2>&1; echo "pytest fmt1 exit $?"; tail -4 gpurun_out/c4_pytest_fmt1.log
