#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_project.py -m gpu -x -q 2>&1 | tail -3
rm -f gpurun_out/r02_sweep_compact.jsonl
timeout 900 python tools/sweep_r2.py compact 1e9 > gpurun_out/r02_sweep_compact.log 2>&1; echo "compact rc=$?"; grep -v BEST gpurun_out/r02_sweep_compact.log | python -c "
import sys, json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    if 'ms' in r: print(r['sel'], round(r['ms'],3), round(r['frac'],3), r['ok'], {k.split('.')[1]:v for k,v in r['cfg'].items()})
    else: print(r)
"
bash tools/gpu_r2_h.sh 2>&1 | grep -E "count_stage|gather|scan" | grep duration | head -8
