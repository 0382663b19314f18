#!/bin/bash
# ncu evidence for one round: launch list of the bench command + full captures of the two hot kernels.
# usage: bash tools/gpu_profile.sh <round-tag>
set -u
TAG=${1:-r01}
mkdir -p gpurun_out/prof
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/prof/launches_$TAG.csv $BENCH > gpurun_out/prof/launches_$TAG.out 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:warpFrameKernel -s 20 -c 2 -f -o /tmp/prof_warp_$TAG $BENCH > gpurun_out/prof/ncu_warp_$TAG.out 2>&1
echo "warp capture rc=$?"
# one complete search ladder (44 launches at 4K): skip the first ladder(s) of the warm-up
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:sadPassKernel -s 88 -c 44 -f -o /tmp/prof_sad_$TAG $BENCH > gpurun_out/prof/ncu_sad_$TAG.out 2>&1
echo "sad capture rc=$?"
for k in warp sad; do
  f=/tmp/prof_${k}_$TAG.ncu-rep
  [ -f $f ] || continue
  ls -la $f
  ncu -i $f --page raw --csv > gpurun_out/prof/${k}_raw_$TAG.csv 2>/dev/null
  ncu -i $f --page details --csv > gpurun_out/prof/${k}_details_$TAG.csv 2>/dev/null
  sz=$(stat -c %s $f)
  if [ $sz -lt 25000000 ]; then cp $f gpurun_out/prof/; fi
done
ls -la gpurun_out/prof
