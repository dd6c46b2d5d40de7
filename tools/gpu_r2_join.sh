#!/bin/bash
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_join.py -m gpu -q -x > gpurun_out/pytest_join.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_join.log; tail -40 gpurun_out/pytest_join.log
