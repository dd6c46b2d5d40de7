#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_group_topk.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/diag_group.py 1e9 2>&1 | grep '"G": 1000,' > gpurun_out/diag_group_wp.jsonl; cat gpurun_out/diag_group_wp.jsonl
