#!/bin/bash
# round 2, 8 GPUs of one box: NCCL parity tests through the C ABI, the one-line bench under torchrun (the driver's command line)
N=8
mkdir -p gpurun_out
nvidia-smi -L | head -8 > gpurun_out/box_8gpu.txt; nvidia-smi topo -m >> gpurun_out/box_8gpu.txt 2>&1; free -g >> gpurun_out/box_8gpu.txt; nproc >> gpurun_out/box_8gpu.txt; numactl -H >> gpurun_out/box_8gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_8gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_8gpu.log; tail -5 gpurun_out/pytest_8gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_8gpu.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_8gpu.json').read().strip().splitlines()[-1])
    print('projection', round(d['ms_per_step'],3),'ms', round(d['roofline']['frac'],3), d['clocks'], 'e2e', d.get('e2e',{}))
    for w,r in d.get('workloads',{}).items():
        print(w, round(r['ms_per_step'],3),'ms', round(r['value']/1e9,1),'Grows/s local', round(r.get('local_ms', r['roofline']['kernel_ms']),3), 'merge', round(r['merge_ms'],3), 'frac', round(r['roofline']['frac'],3), 'ok', r['result_checked'])
except Exception as e: print('ERR', e)
PY
