#!/bin/bash
# per-launch device times of a few steady-state steps with WARM caches (ncu --cache-control none): shares, not bench values
# usage: bash tools/gpu_launches.sh <tag> [bench args]
set -u
TAG=${1:-x}; shift
mkdir -p gpurun_out/prof
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s ${SKIP:-130} -c ${COUNT:-110} --csv --log-file gpurun_out/prof/launches_$TAG.csv \
  python bench.py --steps 4 --warmup 6 --no-cpu-baseline "$@" > gpurun_out/prof/launches_$TAG.out 2>&1
echo "launch list rc=$?"
python tools/launch_table.py gpurun_out/prof/launches_$TAG.csv
