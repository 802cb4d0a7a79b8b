#!/bin/bash
# compute-sanitizer passes over the hand-rolled synchronisation (ticket / flag all-reduce,
# TMA ring + mbarriers, cluster sweeps); logs go to gpurun_out/sanitizer/ (copy the summaries
# to profiles/).  Usage: scripts/sanitize.sh [tag]
tag=${1:-r2}
out=gpurun_out/sanitizer; mkdir -p $out
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # tool, name, kernels...
  tool=$1; name=$2; shift 2
  timeout 600 $CS --tool $tool --print-limit 20 --log-file $out/${tag}_${tool}_${name}.log python scripts/sanitize_target.py "$@" > $out/${tag}_${tool}_${name}.out 2>&1
  echo "$tool $name: exit $? ; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/${tag}_${tool}_${name}.log | tail -1)"
}
run memcheck all spmv pcg ssor phase
run racecheck spmv_pcg spmv pcg
run racecheck ssor ssor      # the point-to-point engine synchronises two warps through volatile shared-memory words on purpose: expect "hazards" there
run synccheck all spmv pcg ssor
