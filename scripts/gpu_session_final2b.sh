#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "two_gpus" > gpurun_out/g2_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/g2_pytest.log
timeout 300 python scripts/multi_check.py 2 > gpurun_out/r2_multi_check_2.log 2>&1; echo "multi_check exit $?"; tail -2 gpurun_out/r2_multi_check_2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29761 bench.py --gpus 2 > gpurun_out/r2_scale_2.json 2> gpurun_out/g2_bench.err; echo "bench2 exit $?"; tail -1 gpurun_out/r2_scale_2.json | cut -c1-200
