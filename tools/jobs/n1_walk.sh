#!/bin/bash
# 1-GPU job after the candidate-walk change: whole -m gpu suite, smoke, select timings at 1e7 / N2 / 1e8 / N3
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_gpu.log | cut -c1-300
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 400 python tools/sweep.py --sizes 10000000,38632323,100000000,675129632 --out gpurun_out/r2_sweep_walk.jsonl > gpurun_out/sweep_walk.log 2>&1; echo "sweep rc=$?"
grep -E "topk" gpurun_out/r2_sweep_walk.jsonl | cut -c1-150
