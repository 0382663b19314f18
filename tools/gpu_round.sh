#!/bin/bash
# One GPU-box visit: environment probe, GPU parity tests, smoke, a short bench.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
{
  echo "== probe"; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv; nproc; free -g | head -2
  ls /etc/OpenCL/vendors 2>&1; ldconfig -p | grep -i -E 'opencl|pocl'; find / -name 'libnvidia-opencl*' 2>/dev/null | head
} > gpurun_out/probe.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -5 gpurun_out/probe.log; tail -30 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
