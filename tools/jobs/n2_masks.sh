#!/bin/bash
# 2-GPU check of the sharded select (tie_base from the lower rank) after the apply change
timeout 300 python -m pytest tests/test_gpu_dist.py -q -m gpu -k "sharded_masks or peer_exchange" > gpurun_out/t_dist_masks.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_dist_masks.log | cut -c1-200
