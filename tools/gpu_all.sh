#!/bin/bash
# Full GPU-box visit: reference-on-OpenCL legs (golden + three-way parity), GPU parity suite, smoke, bench, ncu.
set -u
TAG=${1:-r01}
bash tools/gpu_ref.sh
bash tools/gpu_round.sh
bash tools/gpu_profile.sh $TAG
