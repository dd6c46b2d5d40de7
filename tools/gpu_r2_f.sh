#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
rm -f gpurun_out/r02_sweep_topk.jsonl
timeout 300 python tools/sweep_r2.py topk 8e9 > gpurun_out/r02_sweep_topk.log 2>&1; echo "topk rc=$?"; cat gpurun_out/r02_sweep_topk.jsonl
timeout 600 python tools/diag_csv_stream.py 4e6 > gpurun_out/r02_diag_csv_stream.jsonl 2> gpurun_out/diag_csv.err; echo "csv rc=$?"; cat gpurun_out/r02_diag_csv_stream.jsonl; tail -2 gpurun_out/diag_csv.err
timeout 900 ncu --clock-control none -k regex:'^(wdb_project|wdb_compact_l2|wdb_group_wp|wdb_group|wdb_topk_scan)$' -c 16 --csv --log-file gpurun_out/r02_ncu_sector_efficiency.csv \
  --metrics smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct,smsp__sass_average_data_bytes_per_sector_mem_global_op_st.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed_pipe_lsu.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum \
  python tools/prof_target_r2.py project,compact1,compact50,compact99,group1k,group10m,topk 0.25 > gpurun_out/ncu_sector.log 2>&1; echo "ncu sector rc=$?"; wc -l gpurun_out/r02_ncu_sector_efficiency.csv
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print('projection', round(d['ms_per_step'],3),'ms', round(d['roofline']['frac'],3), 'e2e', d.get('e2e',{}).get('value'), d.get('e2e',{}).get('pcie_yardstick'))
    for w,r in d.get('workloads',{}).items():
        print(w, round(r['ms_per_step'],3),'ms', round(r['value']/1e9,1),'Grows/s kernel', round(r['roofline']['kernel_ms'],3), 'frac', round(r['roofline']['frac'],3), 'ok', r['result_checked'], 'launches', r['gpu_launches'])
except Exception as e: print('ERR', e)
PY
