#!/bin/bash
# round 2, first single-GPU run: parity tests (incl. the new reference-GPU and comm tests), smoke, the one-line bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/box.txt; nproc >> gpurun_out/box.txt; free -g >> gpurun_out/box.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print('projection', round(d['ms_per_step'],3),'ms', round(d['roofline']['frac'],3), d['clocks'], 'e2e', d.get('e2e',{}).get('value'))
    for w,r in d.get('workloads',{}).items():
        print(w, round(r['ms_per_step'],3),'ms', round(r['value']/1e9,1),'Grows/s kernel', round(r['roofline']['kernel_ms'],3), 'frac', round(r['roofline']['frac'],3), 'ok', r['result_checked'], 'launches', r['gpu_launches'], r['clocks']['samples'], r['clocks']['sm_mhz'])
except Exception as e: print('ERR', e)
PY
