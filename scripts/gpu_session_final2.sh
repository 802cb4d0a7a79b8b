#!/bin/bash
# 2-GPU validation of the final tree: full GPU suite, smoke(), both sharded entry points with kept logs.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/f2_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/f2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/f2_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/f2_smoke.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 scripts/dist_check.py > gpurun_out/r2_dist_check_2.log 2>&1; echo "dist_check exit $?"; tail -4 gpurun_out/r2_dist_check_2.log
timeout 600 python scripts/multi_check.py 2 > gpurun_out/r2_multi_check_2.log 2>&1; echo "multi_check exit $?"; tail -4 gpurun_out/r2_multi_check_2.log
