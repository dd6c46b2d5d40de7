#!/bin/bash
# full single-GPU evidence run: tests, all bench workloads, ncu launch list + full capture
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_projection.json 2> gpurun_out/bench_projection.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench ref rc=$?"
for w in filter1 filter50 filter99 group1k group10m topk5; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
done
python - <<'PY'
import json
for w in ('projection','filter1','filter50','filter99','group1k','group10m','topk5'):
    try:
        d=json.load(open(f'gpurun_out/bench_{w}.json'))
        print(w, round(d['ms_per_step'],3),'ms', round(d['value']/1e9,1),'Grows/s', round(d['roofline']['achieved']),'GB/s', round(d['roofline']['frac'],3), d['config'].get('result_checked'), d['gpu_launches'], d.get('e2e',{}).get('value'))
    except Exception as e: print(w, 'ERR', e)
PY
python bench.py --steps 5 --warmup 3 --no-e2e --no-ref --no-cpu > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-ref --no-cpu > gpurun_out/ncu_bench.log 2>&1
echo "ncu launch list rc=$?"
python tools/prof_target.py 268435456 > gpurun_out/plain_prof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:wdb_ -c 24 -o gpurun_out/prof_r01b \
    python tools/prof_target.py 268435456 > gpurun_out/ncu_prof.log 2>&1
echo "ncu full rc=$?"
