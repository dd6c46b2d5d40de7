#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_group_topk.py tests/test_gpu_comm.py -m gpu -x -q 2>&1 | tail -3
rm -f gpurun_out/r02_sweep_group1k.jsonl
timeout 600 python tools/sweep_r2.py group1k 1e9 > gpurun_out/r02_sweep_group1k.log 2>&1; echo "group1k rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r02_sweep_group1k.jsonl'):
    r=json.loads(l)
    if 'error' in r: print(r); continue
    print(r['G'], r['agg'], {k.split('.')[1]:v for k,v in r['cfg'].items()}, round(r['ms_reset_plus_consume'],3), round(r['frac'],3), r['ok'])
PY
