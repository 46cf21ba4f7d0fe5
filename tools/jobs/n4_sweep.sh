#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29516 tools/sweep_dp.py --steps 5 --out gpurun_out/r2_sweep_dp_n4.jsonl > gpurun_out/sweep_dp_n4.log 2>&1; echo "sweep rc=$?"; grep -E '^\{' gpurun_out/sweep_dp_n4.log | cut -c1-250
timeout 200 $TR --master-port 29517 tools/dit_e2e.py --arm ours --dtype bf16 --cuda-graph --dp-exchange peer-overlap --overlap-parts 4 --overlap-ctas 48 --steps 20 --warmup 5 --out gpurun_out/dit_e2e_n4.jsonl > gpurun_out/dit_n4_ov.log 2>&1; echo "dit rc=$?"; tail -1 gpurun_out/dit_n4_ov.log | cut -c150-450
timeout 200 $TR --master-port 29518 tools/dit_e2e.py --arm ours --dtype bf16 --cuda-graph --dp-exchange peer --steps 20 --warmup 5 --out gpurun_out/dit_e2e_n4.jsonl > gpurun_out/dit_n4_peer.log 2>&1; echo "dit rc=$?"; tail -1 gpurun_out/dit_n4_peer.log | cut -c150-450
