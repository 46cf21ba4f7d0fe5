#!/bin/bash
# 1-GPU job: the whole -m gpu suite (incl. new peer/realsize/runner tests), then a short bench.py
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/t_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
