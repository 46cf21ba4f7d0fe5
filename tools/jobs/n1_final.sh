#!/bin/bash
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['traffic_source'], d['e2e'], d['extra'], d['clocks'])
PY
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-300
timeout 300 python tools/sweep.py --sizes 10000000,38632323 --out gpurun_out/r2_sweep_small2.jsonl > gpurun_out/sweep_small2.log 2>&1; echo "sweep rc=$?"
grep -E "topk_select_total\"|topk_select_total_eager" gpurun_out/r2_sweep_small2.jsonl | cut -c1-140
