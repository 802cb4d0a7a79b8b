#!/bin/bash
# 8-GPU check of the final tree: the bench line as the driver launches it, and the parity log of the
# single-process multi-GPU path.
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 8 > gpurun_out/r2_scale_8.json 2> gpurun_out/f8_bench.err; echo "bench8 exit $?"; tail -1 gpurun_out/r2_scale_8.json | cut -c1-900; tail -2 gpurun_out/f8_bench.err
timeout 300 python scripts/multi_check.py 8 > gpurun_out/r2_multi_check_8.log 2>&1; echo "multi_check exit $?"; tail -3 gpurun_out/r2_multi_check_8.log
