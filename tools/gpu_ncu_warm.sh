#!/bin/bash
# ncu with caches left warm (--cache-control none): steady-state L2 hit rates / DRAM traffic of selected kernels
set -u
TAG=$1; K=$2; S=$3; C=$4
mkdir -p gpurun_out/prof
timeout 1200 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__t_sector_hit_rate.pct -k regex:$K -s $S -c $C --csv --log-file gpurun_out/prof/warm_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof/warm_$TAG.out 2>&1
echo "rc=$?"
