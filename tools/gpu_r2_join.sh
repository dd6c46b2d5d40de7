#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_join.py -m gpu -q -x > gpurun_out/pytest_join.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_join.log; tail -15 gpurun_out/pytest_join.log
timeout 200 python tools/bench_join.py > gpurun_out/bench_join.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_join.log; tail -12 gpurun_out/bench_join.log
