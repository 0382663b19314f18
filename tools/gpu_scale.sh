#!/bin/bash
# multi-GPU leg: bench.py under torchrun at N = $1 (the launch line the driver uses)
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/scale_${N}.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 100 --warmup 5 >> gpurun_out/scale_${N}.log 2> gpurun_out/scale_${N}.err
echo "rc=$?"; tail -3 gpurun_out/scale_${N}.log | cut -c1-1200; tail -5 gpurun_out/scale_${N}.err
