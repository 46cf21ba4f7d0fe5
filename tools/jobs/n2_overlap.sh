#!/bin/bash
# 2-GPU job: distributed tests (incl. overlapped backward), DiT-XL/2 loop with the exchange variants
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 python -m pytest tests/test_gpu_dist.py -q -m gpu > gpurun_out/t_dist2.log 2>&1; echo "tests rc=$?"; tail -30 gpurun_out/t_dist2.log | cut -c1-250
for X in peer peer-overlap reduce_scatter; do
  timeout 300 $TR --master-port 29514 tools/dit_e2e.py --arm ours --dtype bf16 --cuda-graph --dp-exchange $X --steps 20 --warmup 5 --out gpurun_out/dit_e2e_n2.jsonl > gpurun_out/dit_n2_$X.log 2>&1; echo "dit $X rc=$?"; tail -2 gpurun_out/dit_n2_$X.log | cut -c1-500
done
timeout 300 $TR --master-port 29515 tools/dit_e2e.py --arm ours --dtype bf16 --dp-exchange peer-overlap --steps 10 --warmup 3 --out gpurun_out/dit_e2e_n2.jsonl > gpurun_out/dit_n2_overlap_eager.log 2>&1; echo "dit overlap eager rc=$?"; tail -2 gpurun_out/dit_n2_overlap_eager.log | cut -c1-500
