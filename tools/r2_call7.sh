#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/c7_*
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_fullsize.py -x -q > gpurun_out/c7_pytest1.txt 2>&1
echo "pytest exit $?" >> gpurun_out/c7_pytest1.txt
tail -3 gpurun_out/c7_pytest1.txt
if ! grep -q "pytest exit 0" gpurun_out/c7_pytest1.txt; then exit 1; fi
run() { echo "== $*" >> gpurun_out/c7_cfgs.txt; env "${@:2}" timeout 600 python bench.py --config $1 --steps 5 --no-e2e --no-cpu-baseline 2>> gpurun_out/c7_cfgs.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  px/s %.3e ms %.3f' % (d['value'], d['ms_per_step']), {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()})" >> gpurun_out/c7_cfgs.txt; }
run cfg4 A=1
run cfg4 UMPA_TAB_STREAM=0
run cfg4 UMPA_TAB_STREAM=2 UMPA_TAB_EH=16
run cfg4 UMPA_TAB_STREAM=2 UMPA_TAB_EH=24
run cfg5 A=1
run cfg5 UMPA_TAB_STREAM=2
run cfg1 A=1
run cfg2 A=1
run cfg2 UMPA_TAB_STREAM=2
cat gpurun_out/c7_cfgs.txt
python tools/diag_first_call.py 4 > gpurun_out/c7_first.txt 2>&1
UMPA_HOST_THREADS=0 python tools/diag_first_call.py 3 >> gpurun_out/c7_first.txt 2>&1
cat gpurun_out/c7_first.txt
