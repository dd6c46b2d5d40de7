#!/bin/bash
# round 2: kernel sweeps (compaction variant 4, wp MATCH mode, dense table geometry) on one GPU
mkdir -p gpurun_out
rm -f gpurun_out/r02_sweep_*.jsonl
timeout 600 python tools/sweep_r2.py compact 1e9 > gpurun_out/r02_sweep_compact.log 2>&1; echo "compact rc=$?"; grep BEST gpurun_out/r02_sweep_compact.log
timeout 600 python tools/sweep_r2.py group1k 1e9 > gpurun_out/r02_sweep_group1k.log 2>&1; echo "group1k rc=$?"
timeout 600 python tools/sweep_r2.py group10m 1e9 > gpurun_out/r02_sweep_group10m.log 2>&1; echo "group10m rc=$?"
python - <<'PY'
import json
def best(path, key, n=6):
    rows=[json.loads(l) for l in open(path) if l.strip()]
    groups={}
    for r in rows:
        if 'error' in r: print('ERR', r); continue
        groups.setdefault(key(r),[]).append(r)
    for k,v in sorted(groups.items(), key=lambda kv: str(kv[0])):
        v.sort(key=lambda r: r.get('ms', r.get('ms_reset_plus_consume')))
        for r in v[:n]: print(k, round(r.get('ms', r.get('ms_reset_plus_consume')),3), round(r['frac'],3), r['ok'], r['cfg'])
best('gpurun_out/r02_sweep_group1k.jsonl', lambda r:(r['G'],r['agg'],r['cfg']['group.wp_mode']), 3)
best('gpurun_out/r02_sweep_group10m.jsonl', lambda r:r['cfg']['group.dense_passes'], 3)
PY
