#!/bin/bash
# first GPU pass: parity tests, smoke, variant sweeps, reference-JIT probe
mkdir -p gpurun_out
{ nvidia-smi -L; nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks.mem,power.draw,memory.total --format=csv; nproc; free -g | head -2; } > gpurun_out/box.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
rm -f gpurun_out/sweep_project.jsonl gpurun_out/sweep_compact.jsonl gpurun_out/probe_refjit.jsonl
timeout 900 python tools/sweep.py project > gpurun_out/sweep_project.log 2>&1
timeout 900 python tools/sweep.py compact > gpurun_out/sweep_compact.log 2>&1
timeout 120 python tools/sweep.py refjit 1 > gpurun_out/refjit1.log 2>&1
timeout 120 python tools/sweep.py refjit 0 > gpurun_out/refjit0.log 2>&1
tail -5 gpurun_out/pytest_gpu.log gpurun_out/smoke.log; tail -3 gpurun_out/sweep_project.log; tail -2 gpurun_out/refjit1.log gpurun_out/refjit0.log
