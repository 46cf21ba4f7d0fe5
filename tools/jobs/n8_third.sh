#!/bin/bash
# 8-GPU job: DiT-XL/2 loop with the exchange pipelined over P pieces; 4-rank data points on the same box
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR8 --master-port 29514 tools/dit_e2e.py --arm ours --dtype bf16 --cuda-graph --dp-exchange peer --steps 20 --warmup 5 --out gpurun_out/dit_e2e_n8c.jsonl > gpurun_out/dit_n8c_peer.log 2>&1; echo "dit peer rc=$?"; tail -1 gpurun_out/dit_n8c_peer.log | cut -c150-500
for PC in "4 32" "4 48" "8 32" "3 32"; do
  set -- $PC
  timeout 300 $TR8 --master-port 29515 tools/dit_e2e.py --arm ours --dtype bf16 --cuda-graph --dp-exchange peer-overlap --overlap-parts $1 --overlap-ctas $2 --steps 20 --warmup 5 --out gpurun_out/dit_e2e_n8c.jsonl > gpurun_out/dit_n8c_ov$1_$2.log 2>&1; echo "dit overlap parts=$1 ctas=$2 rc=$?"; tail -1 gpurun_out/dit_n8c_ov$1_$2.log | cut -c150-500
done
timeout 200 $TR4 --master-port 29516 tools/xchg_bench.py --iters 3 --out gpurun_out/xchg_n4.jsonl > gpurun_out/xchg_n4.log 2>&1; echo "xchg4 rc=$?"
grep -E '^\{' gpurun_out/xchg_n4.log | python -c '
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if d["transport"] in ("tma", "nccl", "multimem"): print(d["dtype"], d["transport"], d["op"], d["ms"], d["link_GBps"])
'
timeout 300 $TR4 --master-port 29517 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "bench4 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n4.json').read().strip().splitlines()[-1])
print(d['exchange']['transport'], d['ms_per_step'], d['value'], {k:v['ms'] for k,v in d['kernels'].items()})
for k,v in d['extra'].items(): print(k, v.get('ms_per_step'), v.get('value'))
print(d['e2e'])
PY
