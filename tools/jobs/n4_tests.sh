#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_dist.py -q -m gpu -k "nccl_reduce_scatter or bucketed" > gpurun_out/t_dist_n4.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t_dist_n4.log | cut -c1-250
