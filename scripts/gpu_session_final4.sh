#!/bin/bash
mkdir -p gpurun_out
for n in 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2974$n bench.py --gpus $n > gpurun_out/r2_scale_$n.json 2> gpurun_out/f4_bench_$n.err; echo "bench$n exit $?"; tail -1 gpurun_out/r2_scale_$n.json | cut -c1-160
done
timeout 300 python bench.py --gpus 1 --no-sweep --no-cpu > gpurun_out/r2_scale_1.json 2> gpurun_out/f4_bench_1.err; echo "bench1 exit $?"; tail -1 gpurun_out/r2_scale_1.json | cut -c1-160
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29750 scripts/dist_check.py > gpurun_out/r2_dist_check_4.log 2>&1; echo "dist_check exit $?"; tail -2 gpurun_out/r2_dist_check_4.log
