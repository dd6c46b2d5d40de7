#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 600 python tools/diag_group.py 1e9 > gpurun_out/diag_group.jsonl 2>&1; cat gpurun_out/diag_group.jsonl
for w in filter1 filter50 filter99 group1k group10m topk5; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
done
python - <<'PY'
import json
for w in ('filter1','filter50','filter99','group1k','group10m','topk5'):
    try:
        d=json.load(open(f'gpurun_out/bench_{w}.json'))
        print(w, round(d['ms_per_step'],3),'ms', round(d['value']/1e9,1),'Grows/s', round(d['roofline']['achieved']),'GB/s', round(d['roofline']['frac'],3), d['config'].get('result_checked'), d['gpu_launches'])
    except Exception as e: print(w, 'ERR', e)
PY
