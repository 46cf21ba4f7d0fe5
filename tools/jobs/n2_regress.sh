#!/bin/bash
# 2-GPU regression after kernel-header changes: distributed + emulated-rank suites, bench.py at N=2 (main line only)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_peer.py -q -m gpu > gpurun_out/t_dist_final.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_dist_final.log | cut -c1-200
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print(d['exchange']['transport'], d['ms_per_step'], d['value'], d['roofline'])
for k,v in (d.get('extra') or {}).items(): print(k, v.get('ms_per_step'), v.get('value'))
print(d['e2e'])
PY
grep -c "NCCL INFO" gpurun_out/bench_n2.err
