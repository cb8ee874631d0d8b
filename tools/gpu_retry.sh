#!/bin/bash
# retry a gpurun call until it is not answered "transient"
cd /root/repo
for n in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout 480 -- "$1" > /tmp/gpurun_try.log 2>&1
  if ! grep -q "status=transient" /tmp/gpurun_try.log; then break; fi
  sleep 45
done
cat /tmp/gpurun_try.log | tail -40
