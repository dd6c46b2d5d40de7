#!/bin/bash
mkdir -p gpurun_out
python tools/diag_numa.py > gpurun_out/diag_numa.log 2>&1; tail -30 gpurun_out/diag_numa.log
python tools/sanitize_target.py 300007 > gpurun_out/plain_sanitize.log 2>&1 && \
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_target.py 300007 > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?"; tail -15 gpurun_out/memcheck.log
