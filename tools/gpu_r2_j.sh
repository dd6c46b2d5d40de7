#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print('projection', round(d['ms_per_step'],3),'ms', round(d['roofline']['frac'],3), 'e2e', d.get('e2e',{}).get('value'))
    for w,r in d.get('workloads',{}).items():
        print(w, round(r['ms_per_step'],3),'ms', round(r['value']/1e9,1),'Grows/s kernel', r['roofline']['kernel'], round(r['roofline']['kernel_ms'],3), 'frac', round(r['roofline']['frac'],3), 'ok', r['result_checked'], 'launches', r['gpu_launches'], r.get('optimizer'))
except Exception as e: print('ERR', e)
PY
python bench.py --steps 2 --warmup 3 --no-e2e --no-ref --no-cpu > gpurun_out/plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-ref --no-cpu > gpurun_out/ncu_bench.log 2>&1
echo "ncu launch list rc=$?"
python tools/prof_target_r2.py compact1 > gpurun_out/plain_prof3.log 2>&1; tail -1 gpurun_out/plain_prof3.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^(wdb_compact_l2|wdb_count_stage|wdb_gather_stage)$' -c 5 -o gpurun_out/prof_r02_filter1 -f python tools/prof_target_r2.py compact1 > gpurun_out/ncu_prof3.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_r02_filter1.ncu-rep
