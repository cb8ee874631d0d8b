#!/bin/bash
mkdir -p gpurun_out/n8b
F=gpurun_out/n8b
for f in $F/*; do [ -f "$f" ] && rm -f "$f"; done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/diag_gather.py > $F/gather.txt 2>&1
grep world $F/gather.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 8 --steps 20 --warmup 5 > $F/bench_n8.json 2> $F/bench_n8.err
echo "cfg2 exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 8 --config cfg4 --no-e2e --steps 10 --warmup 3 > $F/bench_cfg4_n8.json 2> $F/bench_cfg4_n8.err
echo "cfg4 exit $?"
python - <<'PY'
import json
for f in ("bench_n8.json","bench_cfg4_n8.json"):
    d=[json.loads(l) for l in open("gpurun_out/n8b/"+f) if l.startswith("{")][-1]
    print(f, d["ms_per_step"], d["value"], d["roofline"]["stage_ms"], d["gather"]["gather_ms"], d["e2e"] and d["e2e"]["ms_per_step"])
PY
