#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/powerlaw_probe.py 21 22 > gpurun_out/p1_powerlaw.log 2>&1; echo "powerlaw exit $?"; cat gpurun_out/p1_powerlaw.log | tail -5
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/p1_pytest.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/p1_pytest.log
