#!/bin/bash
# round 2, call 1: FFMA2 probe, walk variants, GPU test suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt 2>&1
./tools/probe/ffma2_probe > gpurun_out/c1_ffma2.txt 2>&1
for v in R A6 A8 B6 B8; do
  echo "== lib$v" >> gpurun_out/c1_walk.txt
  UMPA_LIB=$PWD/build/variants/lib$v.so timeout 300 python tools/diag_walk.py >> gpurun_out/c1_walk.txt 2>&1
done
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/c1_pytest.txt
tail -5 gpurun_out/c1_pytest.txt
cat gpurun_out/c1_ffma2.txt gpurun_out/c1_walk.txt
