#!/bin/bash
# full single-GPU evidence run (round 1, second half): tests, smoke, all bench workloads, GROUP BY
# sweeps, ncu launch list + full capture of every hot kernel
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/box.txt; nvidia-smi --query-gpu=clocks.current.sm,clocks.max.sm,clocks.current.memory,power.draw,memory.total --format=csv >> gpurun_out/box.txt; nproc >> gpurun_out/box.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
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
        d=json.loads(open(f'gpurun_out/bench_{w}.json').read().strip().splitlines()[-1])
        print(w, round(d['ms_per_step'],3),'ms', round(d['value']/1e9,1),'Grows/s', round(d['roofline']['achieved']),'GB/s', round(d['roofline']['frac'],3), d['config'].get('result_checked'), d['gpu_launches'], d.get('e2e',{}).get('value'))
    except Exception as e: print(w, 'ERR', e)
PY
if [ -z "$SKIP_DIAG" ]; then
timeout 600 python tools/diag_group_wp.py 1e9 > gpurun_out/diag_group_wp.jsonl 2> gpurun_out/diag_group_wp.err; echo "diag wp rc=$?"
timeout 600 python tools/diag_group_dense.py 1e9 > gpurun_out/diag_group_dense.jsonl 2> gpurun_out/diag_group_dense.err; echo "diag dense rc=$?"
fi
timeout 300 python tools/diag_sort.py > gpurun_out/diag_sort.jsonl 2>&1; echo "diag sort rc=$?"; cat gpurun_out/diag_sort.jsonl
python bench.py --steps 5 --warmup 3 --no-e2e --no-ref --no-cpu > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-ref --no-cpu > gpurun_out/ncu_bench.log 2>&1
echo "ncu launch list rc=$?"
python tools/prof_target.py 268435456 > gpurun_out/plain_prof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:wdb_ -c 40 -o gpurun_out/prof_r01c -f \
    python tools/prof_target.py 268435456 > gpurun_out/ncu_prof.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_prof.log
