#!/bin/bash
# round 2 final measurements on one GPU: full test suite, bench line, other configs, masked config 2, ncu evidence
mkdir -p gpurun_out/final2
F=gpurun_out/final2
for f in $F/*; do [ -f "$f" ] && rm -f "$f"; done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $F/smi.txt
timeout 2400 python -m pytest tests -m gpu -x -q > $F/pytest_gpu.txt 2>&1
echo "pytest exit $?" >> $F/pytest_gpu.txt
tail -4 $F/pytest_gpu.txt
timeout 900 python bench.py > $F/bench_n1.json 2> $F/bench_n1.err
echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $F/bench_ref.json 2> $F/bench_ref.err
for c in cfg1 cfg3 cfg4 cfg5; do
  timeout 900 python bench.py --config $c --steps 5 --no-cpu-baseline > $F/bench_$c.json 2> $F/bench_$c.err
done
python tools/diag_first_call.py 12 > $F/first_call.txt 2>&1
timeout 600 python tools/prof_masked.py cfg2 0.03 > $F/masked_cfg2.txt 2>&1
timeout 600 python tools/prof_masked.py cfg2 0.01 --no-lazy >> $F/masked_cfg2.txt 2>&1
timeout 600 python tools/prof_masked.py cfg2 0.001 --no-lazy >> $F/masked_cfg2.txt 2>&1
# ncu: launch list of one short bench run, then the full capture of a config-2 step
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $F/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"table_walk|shift_table|moments|center_|lazy_|ktable" -c 400 --csv --log-file $F/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $F/ncu_launches.log 2>&1
python tools/prof_step.py cfg2 2 > $F/plain_prof.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"table_walk|shift_table|moments" -s 4 -c 4 -o $F/prof_r02 python tools/prof_step.py cfg2 2 > $F/ncu_prof.log 2>&1
ls -la $F
python - <<'PY'
import json
d=json.load(open('gpurun_out/final2/bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['stage_ms'])
print('e2e', {k:d['e2e'][k] for k in ('value','ms_per_step','first_call_ms')}, d['e2e']['float32_frames']['ms_per_step'])
print('parity', d['parity'])
PY
cat $F/masked_cfg2.txt
