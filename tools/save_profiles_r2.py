#!/usr/bin/env python
"""gpurun_out/prof_r02*.ncu-rep (ncu --set full over tools/prof_target_r2.py: one workload per group of
launches, at the sizes bench.py reports) -> profiles/r02_ncu_full_summary.csv and profiles/traffic.json
(DRAM bytes per row and launch, keyed by bench workload)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
reports = sys.argv[1:] or ["prof_r02.ncu-rep"]
# launch order of tools/prof_target_r2.py (two launches per workload; 10 M keys = two index slices per consume)
ORDER = {"wdb_project": [("projection", 1e9)] * 2,
         "wdb_compact_l2": [("filter1", 4e9)] * 1 + [("filter50", 4e9)] * 2 + [("filter99", 4e9)] * 2,
         "wdb_count_stage": [("filter1", 4e9)] * 2, "wdb_gather_stage": [("filter1", 4e9)] * 2,
         "wdb_group_wp": [("group1k", 2e9)] * 2,
         "wdb_group": [("group10m", 2e9)] * 4,
         "wdb_topk_scan": [("topk5", 8e9)] * 2}
WANT = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct', 'smsp__sass_average_data_bytes_per_sector_mem_global_op_st.pct',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum', 'lts__t_sectors_srcunit_tex_op_red.sum',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed_pipe_lsu.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
SCALE = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}
TSCALE = {'ms': 1.0, 'us': 1e-3, 'ns': 1e-6, 's': 1e3}
out_rows, traffic = [], {}
cols = None
for rep in reports:
    raw = subprocess.run(["ncu", "-i", os.path.join(OUT, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
    if cols is None:
        cols = [w for w in WANT if w in hdr] + stall
        out_rows.append(["workload", "rows"] + cols)
        out_rows.append(["", ""] + [units[hdr.index(c)] for c in cols])
    seen = {}
    for r in data:
        k = r[hdr.index('Kernel Name')].split('(')[0]
        i = seen.get(k, 0)
        seen[k] = i + 1
        wl, n = ORDER.get(k, [("?", 0)] * 99)[min(i, len(ORDER.get(k, [0])) - 1)] if k in ORDER else ("?", 0)
        out_rows.append([wl, int(n)] + [r[hdr.index(c)] if c in hdr else "" for c in cols])
        rd = float(r[hdr.index('dram__bytes_read.sum')]) * SCALE[units[hdr.index('dram__bytes_read.sum')]]
        wr = float(r[hdr.index('dram__bytes_write.sum')]) * SCALE[units[hdr.index('dram__bytes_write.sum')]]
        ms = float(r[hdr.index('gpu__time_duration.sum')]) * TSCALE[units[hdr.index('gpu__time_duration.sum')]]
        if n:
            traffic[f"{k}:{wl}"] = {"kernel": k, "workload": wl, "rows": int(n), "dram_bytes_read": rd, "dram_bytes_write": wr,
                                    "dram_bytes_per_row": (rd + wr) / n, "duration_ms_under_ncu": ms, "launches_per_step": 2 if wl == "group10m" else 1}
with open(os.path.join(PROF, "r02_ncu_full_summary.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["# ncu --set full --clock-control none --import-source on over tools/prof_target_r2.py (BASELINE sizes on one B200); one row per launch"])
    w.writerows(out_rows)
json.dump({"source": "profiles/r02_ncu_full_summary.csv (ncu --set full at the benchmarked sizes)", "kernels": traffic},
          open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
for k, v in traffic.items():
    print(k, round(v["dram_bytes_per_row"], 3), "B/row", round(v["duration_ms_under_ncu"], 3), "ms")
