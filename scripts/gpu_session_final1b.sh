#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/f1_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/f1_smoke.log
timeout 600 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/f1_bench.err; echo "bench exit $?"; tail -1 gpurun_out/r2_bench_1gpu.json | cut -c1-300
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_bench_reference.json 2>&1; echo "ref exit $?"; tail -1 gpurun_out/r2_bench_reference.json | cut -c1-300
