#!/bin/bash
# ncu evidence for one round: launch list of the bench command + `--set full` capture of every kernel of one steady-state
# step (ingest, the 22 search passes, blur, 6 warps at cfg3).  usage: bash tools/gpu_profile.sh <round-tag> [workload]
set -u
TAG=${1:-r01}
WL=${2:-cfg3}
mkdir -p gpurun_out/prof
BENCH="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline"
[ -n "${NO_LIST:-}" ] || timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/prof/launches_$TAG.csv $BENCH > gpurun_out/prof/launches_$TAG.out 2>&1
echo "launch list rc=$?"
K='regex:packPlanarKernel|sadTileKernel|sadCandKernel|sadPassKernel|blurFlow|warpKernel|copyFrameKernel'
# skip the priming uploads and seven whole steps (the first steps of a run still see the flow of the start-up frames,
# whose peak magnitude sends the warp kernel down its general path), then take a bit more than one step
BENCH2="python bench.py --workload $WL --steps 4 --warmup 6 --no-cpu-baseline"
timeout 1500 ncu --set full --clock-control none --import-source on -k "$K" -s 178 -c 30 -f -o /tmp/prof_step_$TAG $BENCH2 > gpurun_out/prof/ncu_step_$TAG.out 2>&1
echo "step capture rc=$?"
f=/tmp/prof_step_$TAG.ncu-rep
if [ -f $f ]; then
  ls -la $f
  ncu -i $f --page raw --csv > gpurun_out/prof/step_raw_$TAG.csv 2>/dev/null
  ncu -i $f --page source --csv > /tmp/step_source_$TAG.csv 2>/dev/null
  python tools/sass_hist.py /tmp/step_source_$TAG.csv 1 14 > gpurun_out/prof/sass_weighted_tile_$TAG.txt 2>&1
fi
ls -la gpurun_out/prof | tail -5
