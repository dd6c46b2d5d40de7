#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_projection.json 2> gpurun_out/bench_projection.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench ref rc=$?"
for w in filter1 filter50 filter99 group1k group10m topk5; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
done
cat gpurun_out/bench_*.json | cut -c1-600
# ncu: launch list of the bench command, then one full capture of the hot kernels
python bench.py --steps 5 --warmup 3 --no-e2e --no-ref --no-cpu > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-ref --no-cpu > gpurun_out/ncu_bench.log 2>&1
echo "ncu launch list rc=$?"
python tools/prof_target.py 268435456 > gpurun_out/plain_prof.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:wdb_ -c 16 -o gpurun_out/prof_r01 \
    python tools/prof_target.py 268435456 > gpurun_out/ncu_prof.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out
