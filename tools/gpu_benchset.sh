#!/bin/bash
# 1-GPU bench set for the record: default line, reference arm, the other workloads, several streams per GPU.
set -u
TAG=${1:-r01}
mkdir -p gpurun_out/benchset
O=gpurun_out/benchset
timeout 900 python bench.py > $O/bench_default_$TAG.json 2> $O/bench_default_$TAG.err; echo "default rc=$?"
HRB_REF_BUDGET_S=40 timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_$TAG.json 2> $O/bench_reference_$TAG.err; echo "reference rc=$?"
for wl in cfg1 cfg2; do
  timeout 600 python bench.py --workload $wl --steps 400 --warmup 5 --no-cpu-baseline > $O/bench_${wl}_$TAG.json 2> $O/bench_${wl}_$TAG.err; echo "$wl rc=$?"
done
for s in 8; do
  timeout 600 python bench.py --streams-per-gpu $s --steps 100 --warmup 5 > $O/bench_streams${s}_$TAG.json 2> $O/bench_streams${s}_$TAG.err; echo "streams $s rc=$?"
done
timeout 900 python bench.py --workload cfg4 --steps 20 --warmup 3 > $O/bench_cfg4_n1_$TAG.json 2> $O/bench_cfg4_n1_$TAG.err; echo "cfg4 rc=$?"
for f in $O/*_$TAG.json; do echo "== $f"; cut -c1-420 $f; done
tail -3 $O/*_$TAG.err | cut -c1-300
