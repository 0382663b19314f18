#!/bin/bash
# quick visit: search-ladder parity tests + launch list of a short bench + (optionally) a full capture of selected kernels
set -u
TAG=${1:-q}
mkdir -p gpurun_out/prof
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "search or smoke or filter or device_resident or warp or levels or pipelined or cpp_shim or overlapped" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_quick.log
tail -5 gpurun_out/pytest_quick.log
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/prof/launches_$TAG.csv $BENCH > gpurun_out/prof/launches_$TAG.out 2>&1
echo "launch list rc=$?"
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.log
