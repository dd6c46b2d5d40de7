#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r02_sweep_compact.jsonl
timeout 600 python tools/sweep_r2.py compact 2e9 > gpurun_out/r02_sweep_compact.log 2>&1; echo "compact rc=$?"; grep -v BEST gpurun_out/r02_sweep_compact.log | python -c "
import sys, json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    if 'ms' in r: print(r['sel'], round(r['ms'],3), round(r['frac'],3), r['ok'], {k.split('.')[1]:v for k,v in r['cfg'].items()})
    else: print(r)
"
python tools/prof_target_r2.py topk > gpurun_out/plain_prof2.log 2>&1; tail -1 gpurun_out/plain_prof2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^(wdb_topk_scan)$' -c 2 -o gpurun_out/prof_r02_topk -f python tools/prof_target_r2.py topk > gpurun_out/ncu_prof2.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_r02_topk.ncu-rep
