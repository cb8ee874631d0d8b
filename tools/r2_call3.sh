#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/c3_*
# correctness first (a wrong ring protocol would hang: bounded by timeout)
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/c3_pytest1.txt 2>&1
echo "pytest parity exit $?" >> gpurun_out/c3_pytest1.txt
tail -5 gpurun_out/c3_pytest1.txt
if ! grep -q "pytest parity exit 0" gpurun_out/c3_pytest1.txt; then exit 1; fi
for v in C6 D6 D7 D8; do
  echo "== lib$v" >> gpurun_out/c3_walk.txt
  UMPA_LIB=$PWD/build/variants/lib$v.so timeout 300 python tools/diag_walk.py >> gpurun_out/c3_walk.txt 2>&1
done
echo "== libD6 UMPA_TAB_STREAM=0" >> gpurun_out/c3_walk.txt
UMPA_TAB_STREAM=0 UMPA_LIB=$PWD/build/variants/libD6.so timeout 300 python tools/diag_walk.py >> gpurun_out/c3_walk.txt 2>&1
for ns in 1 4 8; do
echo "== libD6 UMPA_TAB_NSEG=$ns" >> gpurun_out/c3_walk.txt
UMPA_TAB_NSEG=$ns UMPA_LIB=$PWD/build/variants/libD6.so timeout 300 python tools/diag_walk.py >> gpurun_out/c3_walk.txt 2>&1
done
cat gpurun_out/c3_walk.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c3_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/c3_pytest.txt
tail -5 gpurun_out/c3_pytest.txt
