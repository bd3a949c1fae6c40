#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3 / status=transient): tools/gpurun_retry.sh <timeout> '<command>'
t=$1; shift
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$t" -- "$@" 2>&1)
  echo "$out" | grep -q "status=transient" || { echo "$out"; exit 0; }
  sleep 45
done
echo "$out"; exit 3
