#!/bin/bash
# N-GPU evidence: NCCL parity tests + weak-scaling bench lines (N = number of visible GPUs)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi$N.log; tail -4 gpurun_out/pytest_multi$N.log
for n in $N 4 2; do
  if [ $n -gt $N ]; then continue; fi
  for w in projection filter50 topk5 group1k group10m; do
    extra=""; if [ $w == projection ]; then extra="--e2e-steps 2"; fi
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --workload $w --steps 20 --warmup 3 $extra > gpurun_out/bench${n}_$w.json 2> gpurun_out/bench${n}_$w.err; echo "bench$n $w rc=$?"
  done
  if [ $n -eq 2 ]; then break; fi
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/bench[248]_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'n', d['n_gpus'], round(d['ms_per_step'],3),'ms', round(d['value']/1e9,1),'Grows/s', d['config'].get('result_checked'), 'e2e', d.get('e2e',{}).get('value'))
    except Exception as e: print(f, 'ERR', e)
PY
