#!/bin/bash
# 8-GPU job: exchange micro-bench incl. the TMA transport, bench.py N=8, DiT-XL/2 loop (peer / peer-overlap), config-5 sweep
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29512 tools/xchg_bench.py --iters 3 --out gpurun_out/xchg_n8b.jsonl > gpurun_out/xchg_n8b.log 2>&1; echo "xchg rc=$?"
grep -E '^\{' gpurun_out/xchg_n8b.log | python -c '
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if d["transport"] in ("tma", "nccl", "flags", "multimem"): print(d["dtype"], d["transport"], d["op"], d["ms"], d["link_GBps"])
'
tail -2 gpurun_out/xchg_n8b.log | cut -c1-300
timeout 400 $TR --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench8 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n8.json').read().strip().splitlines()[-1])
print(d['exchange']['transport'], d['ms_per_step'], d['value'], {k:v['ms'] for k,v in d['kernels'].items()})
for k,v in d['extra'].items(): print(k, v.get('ms_per_step'), v.get('value'))
print(d['e2e'])
PY
grep -c "NCCL INFO" gpurun_out/bench_n8.err
timeout 300 $TR --master-port 29514 tools/dit_e2e.py --arm ours --dtype bf16 --cuda-graph --dp-exchange peer --steps 20 --warmup 5 --out gpurun_out/dit_e2e_n8b.jsonl > gpurun_out/dit_n8_peer.log 2>&1; echo "dit peer rc=$?"; tail -1 gpurun_out/dit_n8_peer.log | cut -c1-500
for C in 32 64; do
  timeout 300 $TR --master-port 29515 tools/dit_e2e.py --arm ours --dtype bf16 --cuda-graph --dp-exchange peer-overlap --overlap-ctas $C --steps 20 --warmup 5 --out gpurun_out/dit_e2e_n8b.jsonl > gpurun_out/dit_n8_ov$C.log 2>&1; echo "dit overlap $C rc=$?"; tail -1 gpurun_out/dit_n8_ov$C.log | cut -c1-500
done
timeout 300 $TR --master-port 29516 tools/sweep_dp.py --sizes 10000000,100000000,2000000000 --steps 5 --out gpurun_out/r2_sweep_dp_n8.jsonl > gpurun_out/sweep_dp_n8.log 2>&1; echo "sweep rc=$?"; grep -E '^\{' gpurun_out/sweep_dp_n8.log | cut -c1-300
