#!/bin/bash
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "topk or k2b" 2>&1 | tail -2
