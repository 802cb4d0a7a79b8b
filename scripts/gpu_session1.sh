#!/bin/bash
# round 2, GPU session 1: regression + the pieces round 1 left unmeasured + sanitizer passes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/s1_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/s1_pytest.log
timeout 300 python scripts/bicg_persist_probe.py --time > gpurun_out/s1_bicg_probe.log 2>&1; echo "bicg probe exit $?"; tail -4 gpurun_out/s1_bicg_probe.log
timeout 300 python scripts/gmres_persist_probe.py --time > gpurun_out/s1_gmres_probe.log 2>&1; echo "gmres probe exit $?"; tail -4 gpurun_out/s1_gmres_probe.log
timeout 300 python bench.py > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err; echo "bench exit $?"; cat gpurun_out/s1_bench.json | cut -c1-600
scripts/sanitize.sh r2
