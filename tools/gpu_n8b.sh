#!/bin/bash
# 8-GPU evidence (second half of round 1): the driver's own scaling command line, NCCL parity tests, every workload
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
t0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench${N}_projection.json 2> gpurun_out/bench${N}_projection.err; echo "bench$N projection rc=$? t=$(( $(date +%s) - t0 ))s"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi$N.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - t0 ))s" | tee -a gpurun_out/pytest_multi$N.log; tail -3 gpurun_out/pytest_multi$N.log
for w in filter50 topk5 group1k group10m; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload $w --steps 20 --warmup 3 > gpurun_out/bench${N}_$w.json 2> gpurun_out/bench${N}_$w.err; echo "bench$N $w rc=$? t=$(( $(date +%s) - t0 ))s"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/bench8_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'n', d['n_gpus'], round(d['ms_per_step'],3),'ms', round(d['value']/1e9,1),'Grows/s', d['config'].get('result_checked'), 'e2e', d.get('e2e',{}).get('value'))
    except Exception as e: print(f, 'ERR', e)
PY
