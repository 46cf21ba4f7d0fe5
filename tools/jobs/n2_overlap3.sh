#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 500 python -m pytest tests/test_gpu_dist.py -q -m gpu > gpurun_out/t_dist4.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/t_dist4.log | cut -c1-250
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --no-extra > gpurun_out/bench_n2c.json 2> gpurun_out/bench_n2c.err; echo "bench2 rc=$?"; grep -c "NCCL INFO" gpurun_out/bench_n2c.err; grep -m2 "nranks" gpurun_out/bench_n2c.err | cut -c1-200; wc -l gpurun_out/bench_n2c.json
