#!/bin/bash
# 1-GPU job after the pass-1 rewrite: whole -m gpu suite, kernels at N2/N3, bench with extras; then an A/B of the
# gradient / mask load hint (experimental build with default-cached loads swapped in on the box's scratch copy)
PKG=unified-unlearning-w-remain-geometry_b200
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_gpu.log | cut -c1-300
rm -f gpurun_out/r2_sweep_select.jsonl gpurun_out/sweep_plain.jsonl
timeout 400 python tools/sweep.py --sizes 38632323,675129632 --out gpurun_out/r2_sweep_select.jsonl > gpurun_out/sweep_sel.log 2>&1; echo "sweep rc=$?"
grep -E "\"n\": 675129632" gpurun_out/r2_sweep_select.jsonl | cut -c1-200
grep -E "topk" gpurun_out/r2_sweep_select.jsonl | grep 38632323 | cut -c1-200
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -c 1800 gpurun_out/bench_n1.json
cp tools/tune/libsfron_plain_grad_loads.bin $PKG/libsfron_b200.so
echo "---- default-cached gradient/mask loads"
timeout 400 python tools/sweep.py --sizes 675129632 --skip-select --out gpurun_out/sweep_plain.jsonl > gpurun_out/sweep_plain.log 2>&1; echo "sweep rc=$?"
cut -c1-200 gpurun_out/sweep_plain.jsonl
timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | cut -c1-900
