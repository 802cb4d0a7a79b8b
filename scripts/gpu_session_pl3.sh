#!/bin/bash
mkdir -p gpurun_out
for v in "" seq48 seq32; do
  if [ -n "$v" ]; then export PSP_B200_LIB=$PWD/build/variants/libpsp_$v.so; else unset PSP_B200_LIB; fi
  for var in 3 4 5; do
    echo "== sequential rows '${v:-seq96}', spmv_variant $var (stages $((var-1)))"
    PSP_OPTS="spmv_variant=$var" timeout 300 python scripts/powerlaw_probe.py 21 2>&1 | tail -2
  done
done > gpurun_out/p3_powerlaw_ab.log 2>&1
unset PSP_B200_LIB
echo "== default build, automatic choice, 2^22" >> gpurun_out/p3_powerlaw_ab.log
timeout 300 python scripts/powerlaw_probe.py 22 2>&1 | tail -2 >> gpurun_out/p3_powerlaw_ab.log
cat gpurun_out/p3_powerlaw_ab.log
