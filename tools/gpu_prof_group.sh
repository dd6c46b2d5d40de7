#!/bin/bash
mkdir -p gpurun_out
python tools/prof_group.py > gpurun_out/plain_prof_group.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:wdb_group -c 4 -o gpurun_out/prof_group_r01 python tools/prof_group.py > gpurun_out/ncu_prof_group.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_prof_group.log
timeout 600 python tools/diag_group.py 1e9 2>&1 | grep '"G": 1000,' | head -3
