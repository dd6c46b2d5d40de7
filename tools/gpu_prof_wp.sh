#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_group_wp.py > gpurun_out/plain_prof_wp.log 2>&1; echo "plain rc=$?"; tail -3 gpurun_out/plain_prof_wp.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wdb_group_wp -o gpurun_out/prof_wp_r01 -f python tools/prof_group_wp.py > gpurun_out/ncu_prof_wp.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_prof_wp.log
ls -la gpurun_out/*.ncu-rep
