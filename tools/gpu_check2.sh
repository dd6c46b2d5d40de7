#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_projection.json 2> gpurun_out/bench_projection.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_projection.json'))
for k in ('metric','value','ms_per_step','steps','roofline','e2e','clocks','gpu_launches'): print(k, d.get(k))
PY
python tools/diag_numa.py 2>&1 | grep e2e_step
