#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_join.py -m gpu -q -x -k "not sql" > gpurun_out/pytest_join2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_join2.log; tail -5 gpurun_out/pytest_join2.log
timeout 40 python tools/bench_join.py 268435456 1048576 > gpurun_out/bench_join2.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_join2.log; tail -6 gpurun_out/bench_join2.log
