#!/bin/bash
# multi-GPU evidence at N = $1: independent streams (the driver's launch line), cfg5-style several streams per GPU,
# the reference arm under torchrun (rank 0 works, the others exit), and the 8K split (cfg4).
set -u
N=${1:-2}
O=gpurun_out/multi
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus_$N.log 2>&1
timeout 600 $TR --master-port 29533 bench.py --gpus $N --steps 200 --warmup 5 --no-cpu-baseline > $O/scale_n$N.json 2> $O/scale_n$N.err; echo "scale rc=$?"
timeout 600 $TR --master-port 29535 bench.py --gpus $N --streams-per-gpu 8 --steps 50 --warmup 5 --no-cpu-baseline > $O/streams8_n$N.json 2> $O/streams8_n$N.err; echo "streams8 rc=$?"
HRB_REF_BUDGET_S=20 timeout 600 $TR --master-port 29537 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/reference_n$N.json 2> $O/reference_n$N.err; echo "reference rc=$?"
timeout 900 $TR --master-port 29547 bench.py --workload cfg4 --gpus $N --steps 20 --warmup 3 > $O/cfg4_n$N.json 2> $O/cfg4_n$N.err; echo "cfg4 rc=$?"
for f in $O/*_n$N.json; do echo "== $f"; grep '^{' $f | cut -c1-330; done
for f in $O/*_n$N.err; do echo "== $f"; tail -n 2 $f | cut -c1-300; done
