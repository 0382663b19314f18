#!/bin/bash
# targeted ncu --set full captures.  usage: bash tools/gpu_ncu.sh <tag> <kernel-regex> <skip> <count> [<kernel-regex> <skip> <count> ...]
set -u
TAG=$1; shift
mkdir -p gpurun_out/prof
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
n=0
while [ $# -ge 3 ]; do
  K=$1; S=$2; C=$3; shift 3; n=$((n+1))
  f=/tmp/prof_${TAG}_$n
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c $C -f -o $f $BENCH > gpurun_out/prof/ncu_${TAG}_$n.out 2>&1
  echo "capture $n ($K) rc=$?"
  ncu -i $f.ncu-rep --page raw --csv > gpurun_out/prof/raw_${TAG}_$n.csv 2>/dev/null
  ncu -i $f.ncu-rep --page source --csv > gpurun_out/prof/source_${TAG}_$n.csv 2>/dev/null
  ls -la $f.ncu-rep
  true
done
