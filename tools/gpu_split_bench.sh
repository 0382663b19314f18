#!/bin/bash
set -u
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 1200 python bench.py --workload cfg4 --steps 20 --warmup 3 > gpurun_out/split_bench_$N.log 2> gpurun_out/split_bench_$N.err
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29547 bench.py --workload cfg4 --gpus $N --steps 20 --warmup 3 > gpurun_out/split_bench_$N.log 2> gpurun_out/split_bench_$N.err
fi
echo "rc=$?"; tail -2 gpurun_out/split_bench_$N.log | cut -c1-1500; tail -3 gpurun_out/split_bench_$N.err
