#!/bin/bash
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stripes" > gpurun_out/pytest_stripes.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_stripes.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/gpu_split_test.py > gpurun_out/split_test_$N.log 2>&1; echo "split rc=$?"; tail -5 gpurun_out/split_test_$N.log
