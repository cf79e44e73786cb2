#!/bin/bash
# compute-sanitizer passes over the small GPU parity tests (run on the GPU box through gpurun); logs -> gpurun_out/
# usage: scripts/sanitize.sh <tool> <tag> <pytest -k expression> [files...]
TOOL=$1; TAG=$2; KEXPR=$3; shift 3
FILES=${@:-tests/test_spline_gpu.py tests/test_det_gpu.py tests/test_vmc_gpu.py}
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool "$TOOL" --print-limit 20 --error-exitcode 0 \
  python -m pytest $FILES -m gpu -x -q -k "$KEXPR" -p no:cacheprovider > gpurun_out/r2_san_${TAG}.log 2>&1
echo "rc=$?" >> gpurun_out/r2_san_${TAG}.log
grep -E "ERROR SUMMARY|passed|failed|rc=" gpurun_out/r2_san_${TAG}.log | tail -5
