#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_group_topk.py tests/test_gpu_comm.py tests/test_gpu_fullsize.py -m gpu -x -q -k "topk or order or limit or top" 2>&1 | tail -3
rm -f gpurun_out/r02_sweep_topk.jsonl
timeout 600 python tools/sweep_r2.py topk 8e9 > gpurun_out/r02_sweep_topk.log 2>&1; echo "topk rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r02_sweep_topk.jsonl'):
    r=json.loads(l); print(r['rows'], r['waves'], round(r['ms_incl_host_sync'],3), round(r['frac'],3), r['ok'])
PY
