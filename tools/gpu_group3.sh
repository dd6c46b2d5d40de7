#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_group_topk.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python tools/diag_group.py 1e9 1000,100000,10000000 > gpurun_out/diag_group.jsonl 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/diag_group.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['G'], d['cfg'], round(d['consume_ms'],2), 'ms', round(d['consume_grows_s'],1), 'Grows/s export', round(d['export_ms'],2), d['ok'])
PY
timeout 600 python tools/diag_e2e.py 1e9 > gpurun_out/diag_e2e.jsonl 2>&1; cat gpurun_out/diag_e2e.jsonl
