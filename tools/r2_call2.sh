#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/c2_*
for v in B6 C6; do
  echo "== lib$v" >> gpurun_out/c2_walk.txt
  UMPA_LIB=$PWD/build/variants/lib$v.so timeout 300 python tools/diag_walk.py >> gpurun_out/c2_walk.txt 2>&1
done
cat gpurun_out/c2_walk.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c2_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/c2_pytest.txt
tail -5 gpurun_out/c2_pytest.txt
python tools/prof_step.py cfg2 2 > gpurun_out/c2_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"table_walk|shift_table|moments" -s 4 -c 4 -o gpurun_out/c2_prof python tools/prof_step.py cfg2 2 > gpurun_out/c2_ncu.log 2>&1
tail -3 gpurun_out/c2_plain.log gpurun_out/c2_ncu.log
