#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
rm -f gpurun_out/sweep_compact.jsonl
SWEEP_ONLY_VARIANT=2 timeout 900 python tools/sweep.py compact > gpurun_out/sweep_compact.log 2>&1; tail -2 gpurun_out/sweep_compact.log
python bench.py --workload topk5 --steps 10 --warmup 3 > gpurun_out/bench_topk5.json 2>gpurun_out/bench_topk5.err; cut -c1-200 gpurun_out/bench_topk5.json
