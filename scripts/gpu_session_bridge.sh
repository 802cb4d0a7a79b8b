#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/b1_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/b1_pytest.log
timeout 500 python scripts/format_crossover.py > gpurun_out/b1_crossover.log 2>&1; echo "crossover exit $?"; cat gpurun_out/b1_crossover.log
timeout 400 python bench.py --no-cpu --steps 2 --warmup 3 --powerlaw-rows 0 > gpurun_out/b1_bench.json 2> gpurun_out/b1_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/b1_bench.json").read().strip().splitlines()[-1])
print("pcg", d["value"], "e2e", d["e2e"]["value"], "sss", d["pcg_sss"]["value"])
for r in d["spmv"]:
    print(r["matrix"], r["op"], round(r["us"], 1), round(r["frac"], 3))
PY
