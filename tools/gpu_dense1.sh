#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_group_topk.py tests/test_gpu_warpdb.py tests/test_gpu_fullsize.py -m gpu -x -q -k "group or warp or overflow or chunked or key_range or sql or direct" > gpurun_out/pytest_dense.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_dense.log
timeout 300 python bench.py --workload group10m --steps 10 --warmup 3 > gpurun_out/bench_group10m.json 2> gpurun_out/bench_group10m.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_group10m.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value']/1e9, d['roofline']['frac'], d['config'].get('result_checked'))"
