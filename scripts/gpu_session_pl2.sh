#!/bin/bash
mkdir -p gpurun_out
for v in "" plain b8 b2 plainb8; do
  if [ -n "$v" ]; then export PSP_B200_LIB=$PWD/build/variants/libpsp_$v.so; else unset PSP_B200_LIB; fi
  echo "== variant '${v:-default}'"
  timeout 300 python scripts/powerlaw_probe.py 21 22 2>&1 | tail -4
done > gpurun_out/p2_powerlaw_ab.log 2>&1
cat gpurun_out/p2_powerlaw_ab.log
unset PSP_B200_LIB
PSP_OPTS="spmv_variant=4" timeout 300 python scripts/powerlaw_probe.py 21 > gpurun_out/p2_powerlaw_v4.log 2>&1; echo "3 stages:"; tail -2 gpurun_out/p2_powerlaw_v4.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_spmv_tma -s 6 -c 1 -f -o gpurun_out/r2_spmv_powerlaw python scripts/powerlaw_probe.py 21 > gpurun_out/p2_ncu.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/p2_ncu.log
