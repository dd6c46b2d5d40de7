#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then echo "tests failed; skipping sweeps"; exit 1; fi
python bench.py --steps 20 --warmup 5 --no-e2e --no-ref --no-cpu > gpurun_out/bench_projection_control.json 2>gpurun_out/bench_projection_control.err; cut -c1-160 gpurun_out/bench_projection_control.json
rm -f gpurun_out/sweep_compact.jsonl gpurun_out/sweep_topk.jsonl
timeout 900 python tools/sweep.py compact > gpurun_out/sweep_compact.log 2>&1; tail -2 gpurun_out/sweep_compact.log
timeout 600 python tools/sweep.py topk > gpurun_out/sweep_topk.log 2>&1; tail -2 gpurun_out/sweep_topk.log
timeout 600 python tools/diag_group.py 1e9 > gpurun_out/diag_group.jsonl 2>&1; cat gpurun_out/diag_group.jsonl
