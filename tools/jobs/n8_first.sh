#!/bin/bash
# 8-GPU job: exchange micro-bench at world 8, bench.py N=8 (with comparison lines), DiT-XL/2 loop with the peer exchange vs NCCL
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29512 tools/xchg_bench.py --iters 3 --out gpurun_out/xchg_n8.jsonl > gpurun_out/xchg_n8.log 2>&1; echo "xchg rc=$?"
grep -E '^\{' gpurun_out/xchg_n8.log | python -c '
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d["dtype"], d["transport"], d["op"], d["ms"], d["link_GBps"])
'
tail -3 gpurun_out/xchg_n8.log | cut -c1-300
timeout 400 $TR --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench8 rc=$?"
tail -c 600 gpurun_out/bench_n8.json; tail -3 gpurun_out/bench_n8.err | cut -c1-300
for X in peer reduce_scatter; do
  timeout 300 $TR --master-port 29514 tools/dit_e2e.py --arm ours --dtype bf16 --cuda-graph --dp-exchange $X --steps 20 --warmup 5 --out gpurun_out/dit_e2e_n8.jsonl > gpurun_out/dit_n8_$X.log 2>&1; echo "dit $X rc=$?"; tail -2 gpurun_out/dit_n8_$X.log | cut -c1-600
done
