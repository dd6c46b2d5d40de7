#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_group_topk.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log; tail -15 gpurun_out/pytest_multi.log
for w in projection group1k group10m topk5 filter50; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $w --steps 10 --warmup 3 --e2e-steps 2 > gpurun_out/bench2_$w.json 2> gpurun_out/bench2_$w.err; echo "bench2 $w rc=$?"; tail -2 gpurun_out/bench2_$w.err
done
python - <<'PY'
import json
for w in ('projection','group1k','group10m','topk5','filter50'):
    try:
        d=json.loads(open(f'gpurun_out/bench2_{w}.json').read().strip().splitlines()[-1])
        print(w, 'n_gpus', d['n_gpus'], round(d['ms_per_step'],3),'ms', round(d['value']/1e9,1),'Grows/s', d['config'].get('result_checked'), d.get('e2e',{}).get('value'))
    except Exception as e: print(w, 'ERR', e)
PY
timeout 600 python tools/diag_group.py 1e9 > gpurun_out/diag_group.jsonl 2>&1; cut -c1-260 gpurun_out/diag_group.jsonl
