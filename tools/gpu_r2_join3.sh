#!/bin/bash
mkdir -p gpurun_out
timeout 45 ncu --set full --clock-control none --import-source on -k 'regex:join_(count|emit)' -c 2 -f -o gpurun_out/r02_join_ncu python tools/prof_join.py > gpurun_out/ncu_join.log 2>&1; echo "ncu rc=$?" >> gpurun_out/ncu_join.log; tail -5 gpurun_out/ncu_join.log; ls -la gpurun_out/*.ncu-rep
