#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zonemap.py tests/test_gpu_warpdb.py tests/test_gpu_project.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python tools/diag_zonemap.py 1e9 > gpurun_out/diag_zonemap.jsonl 2>&1; cat gpurun_out/diag_zonemap.jsonl
python bench.py --workload group1k --steps 10 --warmup 3 2>/dev/null | cut -c1-150
python bench.py --workload group10m --steps 10 --warmup 3 2>/dev/null | cut -c1-150
