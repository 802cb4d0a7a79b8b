#!/bin/bash
# 1-GPU validation of the final tree: full GPU suite, smoke(), the default bench line, configs 1 and 4.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/f1_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/f1_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/f1_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/f1_smoke.log
timeout 600 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/f1_bench.err; echo "bench exit $?"; cut -c1-1500 gpurun_out/r2_bench_1gpu.json; tail -3 gpurun_out/f1_bench.err
rm -f gpurun_out/configs_r2final.jsonl
timeout 300 python scripts/run_configs.py r2final c1,c4 > gpurun_out/f1_configs.log 2>&1; echo "configs exit $?"; tail -6 gpurun_out/f1_configs.log | cut -c1-400
