#!/bin/bash
# last verification of the round on the final tree: smoke, the GPU suite, the bench line; ncu capture of the masked walk
mkdir -p gpurun_out/verify
F=gpurun_out/verify
for f in $F/*; do [ -f "$f" ] && rm -f "$f"; done
python -c "import __graft_entry__ as g; g.smoke()" > $F/smoke.txt 2>&1
echo "smoke exit $?"; tail -2 $F/smoke.txt
timeout 2400 python -m pytest tests -m gpu -q > $F/pytest_gpu.txt 2>&1
echo "pytest exit $?" >> $F/pytest_gpu.txt
tail -4 $F/pytest_gpu.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $F/bench_n1.json 2> $F/bench_n1.err
echo "bench exit $?"
timeout 600 python tools/prof_masked.py cfg2 0.03 --no-lazy > $F/masked_plain.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"masked_walk" -s 1 -c 1 -o $F/masked_walk python tools/prof_masked.py cfg2 0.03 --no-lazy > $F/ncu_masked.log 2>&1
cat $F/masked_plain.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/verify/bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','steps','warmup')}, d['roofline']['stage_ms'])
print('e2e', {k:d['e2e'][k] for k in ('value','ms_per_step','first_call_ms')}, d['e2e']['float32_frames']['ms_per_step'])
print('parity', {k:d['parity'][k] for k in ('n_px','err_mismatch','ncalls_mismatch','walk_ties','walk_unexplained','exceptions')})
PY
