#!/bin/bash
# round 2, N GPUs of one box: NCCL parity tests through the C ABI, then the one-line bench under torchrun
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8 > gpurun_out/box_multi.txt; nvidia-smi topo -m >> gpurun_out/box_multi.txt 2>&1; free -g >> gpurun_out/box_multi.txt; nproc >> gpurun_out/box_multi.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_${N}gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${N}gpu.log; tail -30 gpurun_out/pytest_${N}gpu.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${N}gpu.json').read().strip().splitlines()[-1])
    print('projection', round(d['ms_per_step'],3),'ms', round(d['roofline']['frac'],3), d['clocks'], 'e2e', d.get('e2e',{}))
    for w,r in d.get('workloads',{}).items():
        print(w, round(r['ms_per_step'],3),'ms', round(r['value']/1e9,1),'Grows/s local', round(r.get('local_ms', r['roofline']['kernel_ms']),3), 'merge', round(r['merge_ms'],3), 'frac', round(r['roofline']['frac'],3), 'ok', r['result_checked'], r['collectives'])
except Exception as e: print('ERR', e)
PY
