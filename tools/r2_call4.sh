#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/c4_*
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/c4_pytest1.txt 2>&1
echo "pytest parity exit $?" >> gpurun_out/c4_pytest1.txt
tail -3 gpurun_out/c4_pytest1.txt
if ! grep -q "pytest parity exit 0" gpurun_out/c4_pytest1.txt; then exit 1; fi
for v in D7 E7; do
  echo "== lib$v" >> gpurun_out/c4_walk.txt
  UMPA_LIB=$PWD/build/variants/lib$v.so timeout 300 python tools/diag_walk.py >> gpurun_out/c4_walk.txt 2>&1
done
cat gpurun_out/c4_walk.txt
timeout 1200 python -m pytest tests/test_gpu_reference_fullsize.py -x -q -s > gpurun_out/c4_pytest_ref.txt 2>&1
echo "pytest ref exit $?" >> gpurun_out/c4_pytest_ref.txt
tail -30 gpurun_out/c4_pytest_ref.txt | cut -c1-600
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_reference_fullsize.py > gpurun_out/c4_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/c4_pytest.txt
tail -5 gpurun_out/c4_pytest.txt
timeout 600 python bench.py --steps 10 > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
echo "bench exit $?"; tail -3 gpurun_out/c4_bench.err; cut -c1-1500 gpurun_out/c4_bench.json
python tools/prof_step.py cfg2 2 > gpurun_out/c4_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"table_walk|shift_table|moments" -s 4 -c 4 -o gpurun_out/c4_prof python tools/prof_step.py cfg2 2 > gpurun_out/c4_ncu.log 2>&1
tail -n 2 gpurun_out/c4_plain.log gpurun_out/c4_ncu.log
