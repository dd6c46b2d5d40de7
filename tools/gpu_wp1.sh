#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_group_topk.py tests/test_gpu_warpdb.py tests/test_gpu_fullsize.py -m gpu -x -q -k "group or warp or overflow or chunked or key_range or sql" > gpurun_out/pytest_wp.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_wp.log
timeout 900 python tools/diag_group_wp.py 1e9 > gpurun_out/diag_group_wp2.jsonl 2> gpurun_out/diag_group_wp2.err; echo "diag rc=$?"; tail -3 gpurun_out/diag_group_wp2.err
python - <<'PY'
import json
for l in open('gpurun_out/diag_group_wp2.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['G'], d['note'], d['cfg'], d['consume_ms'], 'ms', d['grows_s'], 'Grows/s', d['groups'], d['ok'], 'spilled', d['spilled'])
PY
timeout 300 python bench.py --workload group1k --steps 10 --warmup 3 > gpurun_out/bench_group1k.json 2> gpurun_out/bench_group1k.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_group1k.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value']/1e9, d['roofline'], d['config'].get('result_checked'))"
