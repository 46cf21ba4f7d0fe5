#!/bin/bash
# 1-GPU DiT-XL/2 loop, bf16, CUDA graph: two runs back to back (box-to-box and run-to-run spread of the it/s figure)
for i in 1 2; do
timeout 300 python tools/dit_e2e.py --arm ours --dtype bf16 --cuda-graph --steps 30 --warmup 5 --out gpurun_out/r2_dit_e2e_1gpu_final.jsonl 2>&1 | tail -1 | cut -c1-400
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv,noheader
done
