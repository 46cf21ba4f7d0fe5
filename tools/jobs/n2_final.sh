#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_peer.py -q -m gpu > gpurun_out/t_dist_final.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_dist_final.log | cut -c1-200
timeout 400 $TR --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print(d['exchange']['transport'], d['ms_per_step'], d['value'], {k:v['ms'] for k,v in d['kernels'].items()})
for k,v in d['extra'].items(): print(k, v.get('ms_per_step'), v.get('value'))
print(d['e2e'])
PY
grep -c "NCCL INFO" gpurun_out/bench_n2.err
timeout 300 $TR --master-port 29516 tools/sweep_dp.py --steps 5 --out gpurun_out/r2_sweep_dp_n2.jsonl > gpurun_out/sweep_dp_n2.log 2>&1; echo "sweep rc=$?"; grep -E '^\{' gpurun_out/sweep_dp_n2.log | cut -c1-250
for X in bucketed reduce_scatter peer; do
  timeout 200 $TR --master-port 29517 tools/dit_e2e.py --arm ours --dtype bf16 --dp-exchange $X --steps 10 --warmup 3 --out gpurun_out/dit_e2e_n2_eager.jsonl > gpurun_out/dit_n2e_$X.log 2>&1; echo "dit eager $X rc=$?"; tail -1 gpurun_out/dit_n2e_$X.log | cut -c150-450
done
