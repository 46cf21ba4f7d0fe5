#!/bin/bash
# 1-GPU profiling job: launch lists and ncu --set full of the step kernels, small-vector sweep, 1-GPU end-to-end lines
set -o pipefail
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra"
$B > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/r2_launches_bench_n1.csv $B > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:fused_update_kernel|fisher_accum_kernel|ratio_mask_kernel|masked_sumsq_kernel" -s 8 -c 6 -o gpurun_out/r2_prof_step -f $B > gpurun_out/ncu2.log 2>&1
echo "set full rc=$?"
ncu -i gpurun_out/r2_prof_step.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_step_kernels_n675M_raw.csv 2>/dev/null; echo "raw rc=$?"
python tools/select_once.py 38632323 > gpurun_out/sel_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_select_n38M.csv python tools/select_once.py 38632323 > gpurun_out/ncu3.log 2>&1
echo "select list rc=$?"
timeout 600 python tools/sweep.py --sizes 10000000,38632323,100000000 --out gpurun_out/r2_sweep_small.jsonl > gpurun_out/sweep_small.log 2>&1; echo "sweep rc=$?"
grep -E "topk_select_total|clipped_forget|masked_sumsq|ratio_mask\"" gpurun_out/r2_sweep_small.jsonl | cut -c1-200
timeout 300 python tools/resnet_e2e.py --iters 150 --cuda-graph --out gpurun_out/r2_resnet18_e2e_1gpu.jsonl 2>&1 | tail -1 | cut -c1-400
timeout 300 python tools/dit_e2e.py --arm ours --dtype bf16 --cuda-graph --steps 20 --warmup 5 --out gpurun_out/r2_dit_e2e_1gpu.jsonl 2>&1 | tail -1 | cut -c1-400
