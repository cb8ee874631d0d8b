#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/c6_*
for v in G7 G7b H7; do
  echo "== lib$v" >> gpurun_out/c6_walk.txt
  UMPA_LIB=$PWD/build/variants/lib$v.so timeout 300 python tools/diag_walk.py >> gpurun_out/c6_walk.txt 2>&1
done
cat gpurun_out/c6_walk.txt
UMPA_LIB=$PWD/build/variants/libH7.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/c6_pytest1.txt 2>&1
echo "pytest H7 parity exit $?" >> gpurun_out/c6_pytest1.txt
tail -3 gpurun_out/c6_pytest1.txt
UMPA_LIB=$PWD/build/variants/libH7.so timeout 900 python -m pytest tests/test_gpu_reference_fullsize.py -x -q > gpurun_out/c6_pytest2.txt 2>&1
echo "pytest H7 ref exit $?" >> gpurun_out/c6_pytest2.txt
tail -3 gpurun_out/c6_pytest2.txt
for cfg in cfg1 cfg4 cfg5; do
for v in G7b H7; do
  echo "== $cfg lib$v" >> gpurun_out/c6_cfgs.txt
  UMPA_LIB=$PWD/build/variants/lib$v.so timeout 600 python bench.py --config $cfg --steps 5 --no-e2e --no-cpu-baseline 2>> gpurun_out/c6_cfgs.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  px/s %.3e ms %.3f' % (d['value'], d['ms_per_step']), d['roofline']['stage_ms'])" >> gpurun_out/c6_cfgs.txt
done; done
cat gpurun_out/c6_cfgs.txt
