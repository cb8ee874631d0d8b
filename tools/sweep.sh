#!/bin/bash
# usage: tools/sweep.sh "ENV1=.. ENV2=.." ...   -> one short bench per configuration
mkdir -p gpurun_out
for cfg in "$@"; do
  echo "=== $cfg"
  env $cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']['stage_ms']
print('  px/s %.3e  ms %.3f  moments %.3f cross %.3f mean %.3f walk %.3f  ok %.4f' % (d['value'], d['ms_per_step'], r['moments'], r['cross_table'], r['mean_table'], r['walk'], d['config']['err_ok_fraction']))"
done
