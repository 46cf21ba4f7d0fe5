#!/bin/bash
# 2-GPU validation job (run under gpurun --gpus 2): peer/dist parity tests, exchange micro-bench, bench.py N=2
timeout 300 python -m pytest tests/test_gpu_peer.py tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/t_peer2.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/t_peer2.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/xchg_bench.py --out gpurun_out/xchg_n2b.jsonl > gpurun_out/xchg_n2b.log 2>&1; echo "xchg rc=$?"
grep -E '^\{' gpurun_out/xchg_n2b.log | python -c '
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d["dtype"], d["transport"], d["op"], d["ms"], d["link_GBps"])
'
tail -3 gpurun_out/xchg_n2b.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"
tail -c 3500 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
