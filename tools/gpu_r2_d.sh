#!/bin/bash
# round 2: new parity tests, then ncu evidence at the benchmarked sizes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zonemap.py tests/test_gpu_warpdb.py tests/test_gpu_comm.py tests/test_gpu_group_topk.py -m gpu -x -q > gpurun_out/pytest_d.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_d.log
python bench.py --steps 2 --warmup 3 --no-e2e --no-ref --no-cpu > gpurun_out/plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-ref --no-cpu > gpurun_out/ncu_bench.log 2>&1
echo "ncu launch list rc=$?"; wc -l gpurun_out/r02_launches_bench.csv
python tools/prof_target_r2.py > gpurun_out/plain_prof.log 2>&1; tail -1 gpurun_out/plain_prof.log
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'^(wdb_project|wdb_compact_l2|wdb_group_wp|wdb_group|wdb_topk_scan)$' --launch-skip 0 -c 14 -o gpurun_out/prof_r02 -f \
    python tools/prof_target_r2.py > gpurun_out/ncu_prof.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_prof.log; ls -la gpurun_out/prof_r02.ncu-rep
