#!/bin/bash
# 1-GPU job: the whole -m gpu suite and smoke()
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_gpu.log | cut -c1-300
python __graft_entry__.py smoke 2>&1 | tail -1
