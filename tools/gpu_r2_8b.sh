#!/bin/bash
mkdir -p gpurun_out
for w in 8 1; do
WARPDB_OPT_group__dense_waves=$w timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 --workloads group10m --no-e2e --no-ref --no-cpu > gpurun_out/bench_8gpu_g10m_w$w.json 2> gpurun_out/bench_8gpu_g10m.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_8gpu_g10m_w$w.json').read().strip().splitlines()[-1])
r=d['workloads']['group10m']; print('waves $w', round(r['ms_per_step'],3), 'local', round(r['local_ms'],3), 'merge', round(r['merge_ms'],3), r['result_checked'])
PY
done
