#!/bin/bash
# GPU-box visit for the reference-on-OpenCL legs: golden vectors + the three-way parity test.
set -u
mkdir -p gpurun_out
timeout 900 python tools/make_golden.py gpurun_out/golden > gpurun_out/golden.log 2>&1; echo "golden rc=$?" >> gpurun_out/golden.log
timeout 900 python -m pytest tests/test_ref_opencl.py -m gpu -q -x > gpurun_out/pytest_ref.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ref.log
tail -15 gpurun_out/golden.log; tail -30 gpurun_out/pytest_ref.log
