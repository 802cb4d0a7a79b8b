#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/runs_probe.py --time 128 256 > gpurun_out/s2_runs_probe.log 2>&1; echo "runs probe exit $?"; cat gpurun_out/s2_runs_probe.log | tail -45
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/s2_pytest.log
timeout 300 python scripts/bicg_persist_probe.py --time > gpurun_out/s2_bicg_probe.log 2>&1; echo "bicg probe exit $?"; tail -4 gpurun_out/s2_bicg_probe.log
