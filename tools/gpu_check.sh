#!/bin/bash
# Runs the GPU parity suites one file at a time (a sticky CUDA fault in one cannot poison the next), each under its
# own timeout, and keeps the logs under gpurun_out/ (merged back by gpurun).
#   gpurun --timeout 900 -- 'bash tools/gpu_check.sh [extra pytest args]'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_info.csv 2>&1
rc=0
for f in tests/test_gpu_stages.py tests/test_gpu_unet_eval.py tests/test_gpu_objective.py tests/test_gpu_train.py; do
  [ -f "$f" ] || continue
  name=$(basename "$f" .py)
  echo "=== $f"
  timeout 600 python -m pytest "$f" -q -m gpu -x --tb=short -s "$@" > "gpurun_out/$name.log" 2>&1
  r=$?
  tail -n 40 "gpurun_out/$name.log"
  echo "=== $f exit $r"
  [ $r -ne 0 ] && rc=$r
done
exit $rc
