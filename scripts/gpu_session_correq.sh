#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_correq.py tests/test_gpu_dropin.py -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/c1_pytest.log
timeout 600 python scripts/jdsym_ab.py 64 > gpurun_out/c1_jdsym_64.log 2>&1; echo "jdsym 64 exit $?"; cat gpurun_out/c1_jdsym_64.log | tail -4
timeout 900 python scripts/jdsym_ab.py 128 > gpurun_out/c1_jdsym_128.log 2>&1; echo "jdsym 128 exit $?"; cat gpurun_out/c1_jdsym_128.log | tail -4
timeout 300 python scripts/gather_probe.py > gpurun_out/c1_gather.log 2>&1; echo "gather exit $?"; cat gpurun_out/c1_gather.log | tail -5
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_sweep_p2p -s 2 -c 1 -f -o gpurun_out/r2_sweep_p2p python scripts/ssor_one.py 4096 4 2 > gpurun_out/c1_ncu_p2p.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/c1_ncu_p2p.log
