#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/c5_*
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/c5_pytest1.txt 2>&1
echo "pytest parity exit $?" >> gpurun_out/c5_pytest1.txt
tail -3 gpurun_out/c5_pytest1.txt
if ! grep -q "pytest parity exit 0" gpurun_out/c5_pytest1.txt; then exit 1; fi
for v in D7 F7 F6 G7 G7P; do
  echo "== lib$v" >> gpurun_out/c5_walk.txt
  UMPA_LIB=$PWD/build/variants/lib$v.so timeout 300 python tools/diag_walk.py >> gpurun_out/c5_walk.txt 2>&1
done
echo "== libG7 CTAS=2 (EH=8)" >> gpurun_out/c5_walk.txt
UMPA_TAB_CTAS=2 UMPA_LIB=$PWD/build/variants/libG7.so timeout 300 python tools/diag_walk.py >> gpurun_out/c5_walk.txt 2>&1
echo "== libG7 CTAS=2 EH=8 NSEG=4" >> gpurun_out/c5_walk.txt
UMPA_TAB_CTAS=2 UMPA_TAB_NSEG=4 UMPA_LIB=$PWD/build/variants/libG7.so timeout 300 python tools/diag_walk.py >> gpurun_out/c5_walk.txt 2>&1
echo "== libG7 FB=5" >> gpurun_out/c5_walk.txt
UMPA_TAB_FB=5 UMPA_LIB=$PWD/build/variants/libG7.so timeout 300 python tools/diag_walk.py >> gpurun_out/c5_walk.txt 2>&1
cat gpurun_out/c5_walk.txt
UMPA_TAB_CTAS=2 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/c5_pytest2.txt 2>&1
echo "pytest CTAS=2 parity exit $?" >> gpurun_out/c5_pytest2.txt
tail -3 gpurun_out/c5_pytest2.txt
