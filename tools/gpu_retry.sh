#!/bin/bash
# retry a gpurun call until it is not answered "transient":  tools/gpu_retry.sh '<command>' [extra gpurun flags]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cmd="$1"; shift
for n in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout 480 "$@" -- "$cmd" > gpurun_out/.try.log 2>&1
  if ! grep -q "status=transient" gpurun_out/.try.log; then break; fi
  sleep 45
done
tail -40 gpurun_out/.try.log
