#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 python -m pytest tests/test_gpu_dist.py -q -m gpu -k "overlapped or fused_peer" > gpurun_out/t_dist3.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/t_dist3.log | cut -c1-250
for C in 48 96; do
  timeout 300 $TR --master-port 29514 tools/dit_e2e.py --arm ours --dtype bf16 --cuda-graph --dp-exchange peer-overlap --overlap-ctas $C --steps 20 --warmup 5 --out gpurun_out/dit_e2e_n2.jsonl > gpurun_out/dit_n2_ov$C.log 2>&1; echo "dit overlap ctas=$C rc=$?"; tail -1 gpurun_out/dit_n2_ov$C.log | cut -c1-500
done
