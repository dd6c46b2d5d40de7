#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
SWEEP_ONLY_VARIANT=99 timeout 300 python tools/sweep.py compact 2>&1 | grep "dense-untouched" | cut -c1-160
timeout 300 python tools/diag_zonemap.py 1e9 2>&1 | cut -c1-200
