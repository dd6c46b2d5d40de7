#!/usr/bin/env python
"""Turn gpurun_out/ evidence (ncu report, launch list, bench JSON lines) into the tracked summaries under profiles/."""
import csv, json, os, subprocess, sys, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out"); PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
rep = sys.argv[2] if len(sys.argv) > 2 else "prof_r01b.ncu-rep"
rows_profiled = int(sys.argv[3]) if len(sys.argv) > 3 else 268435456
raw = subprocess.run(["ncu", "-i", os.path.join(OUT, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct', 'smsp__sass_average_data_bytes_per_sector_mem_global_op_st.pct']
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
cols = [w for w in want if w in hdr] + stall
with open(os.path.join(PROF, f"{tag}_ncu_full_summary.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([f"# ncu --set full --clock-control none --import-source on, tools/prof_target.py {rows_profiled} rows, B200; one row per launch"])
    w.writerow(cols); w.writerow([units[hdr.index(c)] for c in cols])
    for r in data:
        w.writerow([r[hdr.index(c)] for c in cols])
scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}
tr = {}
for r in data:
    k = r[hdr.index('Kernel Name')]
    if k in tr:
        continue
    rd = float(r[hdr.index('dram__bytes_read.sum')]) * scale[units[hdr.index('dram__bytes_read.sum')]]
    wr = float(r[hdr.index('dram__bytes_write.sum')]) * scale[units[hdr.index('dram__bytes_write.sum')]]
    tr[k] = {'rows': rows_profiled, 'dram_bytes_read': rd, 'dram_bytes_write': wr, 'dram_bytes_per_row': (rd + wr) / rows_profiled,
             'duration_ms': float(r[hdr.index('gpu__time_duration.sum')])}
json.dump({'source': f'profiles/{tag}_ncu_full_summary.csv (ncu --set full, {rows_profiled} rows)', 'kernels': tr}, open(os.path.join(PROF, 'traffic.json'), 'w'), indent=1)
for k, v in tr.items():
    print(k[:40], round(v['dram_bytes_per_row'], 3), 'B/row', v['duration_ms'], 'ms')
if os.path.exists(os.path.join(OUT, 'launches_bench.csv')):
    shutil.copy(os.path.join(OUT, 'launches_bench.csv'), os.path.join(PROF, f'{tag}_launches_bench_projection.csv'))
for w in ('projection', 'reference', 'filter1', 'filter50', 'filter99', 'group1k', 'group10m', 'topk5'):
    p = os.path.join(OUT, f'bench_{w}.json')
    if os.path.exists(p):
        shutil.copy(p, os.path.join(PROF, f'{tag}_bench_{w}.json'))
