#!/bin/bash
# final 1-GPU job of the round: whole -m gpu suite, smoke, ncu launch list + set full of the step kernels (roofline.traffic),
# select launch list, then bench.py (both arms) and the small-vector sweep
set -o pipefail
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -1
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra"
$B > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/r2_launches_bench_n1.csv $B > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:fused_update_kernel|fisher_accum_kernel|ratio_mask_kernel|masked_sumsq_kernel" -s 8 -c 6 -o gpurun_out/r2_prof_step -f $B > gpurun_out/ncu2.log 2>&1
echo "set full rc=$?"
ncu -i gpurun_out/r2_prof_step.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_step_kernels_n675M_raw.csv 2>/dev/null; echo "raw rc=$?"
python tools/select_once.py 675129632 > gpurun_out/sel_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_select_n675M.csv python tools/select_once.py 675129632 > gpurun_out/ncu3.log 2>&1
echo "select list rc=$?"
cp gpurun_out/r2_ncu_full_step_kernels_n675M_raw.csv profiles/ && python tools/make_traffic.py profiles/r2_ncu_full_step_kernels_n675M_raw.csv && cp profiles/r2_traffic.json gpurun_out/
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['traffic_source'], d['e2e'], d['extra'], d['clocks'])
PY
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-300
timeout 300 python tools/sweep.py --sizes 10000000,38632323,100000000 --out gpurun_out/r2_sweep_small.jsonl > gpurun_out/sweep_small.log 2>&1; echo "sweep rc=$?"
grep -E "topk_select_total\"|topk_select_total_eager" gpurun_out/r2_sweep_small.jsonl | cut -c1-140
