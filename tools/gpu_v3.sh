#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_project.py -m gpu -x -q -k "variants" 2>&1 | tail -3
rm -f gpurun_out/sweep_compact.jsonl
SWEEP_ONLY_VARIANT=3 timeout 900 python tools/sweep.py compact > gpurun_out/sweep_compact.log 2>&1
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/sweep_compact.jsonl')]
for sel in sorted(set(round(r['sel'],2) for r in rows)):
    cs=sorted([r for r in rows if round(r['sel'],2)==sel and r['name']=='wdb_compact'], key=lambda r:r['ms'])
    print('--- sel',sel)
    for r in cs[:6]+cs[-1:]: print(round(r['ms'],3), round(r['gbs']), round(r['frac_measured_peak'],3), r['cfg'])
PY
