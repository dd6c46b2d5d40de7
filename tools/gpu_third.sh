#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python tools/diag_group.py 1e9 > gpurun_out/diag_group.jsonl 2>&1; cat gpurun_out/diag_group.jsonl
rm -f gpurun_out/sweep_compact.jsonl
python tools/sweep.py compact > gpurun_out/sweep_compact.log 2>&1; tail -3 gpurun_out/sweep_compact.log
python bench.py --workload topk5 --steps 10 --warmup 3 > gpurun_out/bench_topk5.json 2>gpurun_out/bench_topk5.err; cut -c1-300 gpurun_out/bench_topk5.json
python - <<'PY'
import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, a, b in (("h2d", d, h), ("d2h", h, d)):
    a.copy_(b, non_blocking=True); torch.cuda.synchronize()
    t = time.perf_counter(); a.copy_(b, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(name, "GB/s", n / dt / 1e9)
PY
