#!/usr/bin/env python
import os, sys, subprocess, time, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for cmd in (["nvidia-smi", "topo", "-m"], ["numactl", "-H"], ["lscpu"]):
    try:
        print("$", " ".join(cmd)); print(subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout[:3000])
    except Exception as e:
        print("unavailable:", e)
print("affinity", sorted(os.sched_getaffinity(0)))
import torch
p = torch.cuda.get_device_properties(0)
bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
for f in ("local_cpulist", "numa_node"):
    try: print(f, open(f"/sys/bus/pci/devices/{bus}/{f}").read().strip())
    except Exception as e: print(f, "unavailable", e)
from warpdb_b200 import _core as wc, ops
wc.check(wc.lib().wdb_init(0))
n = 1_000_000_000
hp = torch.empty(n, dtype=torch.float32, pin_memory=True); hq = torch.empty(n, dtype=torch.int32, pin_memory=True); ho = torch.empty(n, dtype=torch.float32, pin_memory=True)
hp.fill_(1.5); hq.fill_(3)
cols, nc = wc.make_cols([("price", wc.FLOAT32, hp.data_ptr(), n), ("quantity", wc.INT32, hq.data_ptr(), n)])
cnt = C.c_int64(0); devs = (C.c_int * 1)(0)
d = torch.empty(n, dtype=torch.float32, device="cuda")
for i in range(4):
    t = time.perf_counter(); d.copy_(hp, non_blocking=True); torch.cuda.synchronize(); a = time.perf_counter() - t
    t = time.perf_counter(); ho.copy_(d, non_blocking=True); torch.cuda.synchronize(); b = time.perf_counter() - t
    print(json.dumps({"h2d_gbs": 4 * n / a / 1e9, "d2h_gbs": 4 * n / b / 1e9}), flush=True)
del d
for i in range(8):
    t = time.perf_counter()
    wc.check(wc.lib().wdb_multi_project_filter_host(1, devs, cols, nc, b"((price[idx] * quantity[idx]) * 1.08f)", b"", ho.data_ptr(), n, wc.DENSE_ZERO, C.byref(cnt)))
    print(json.dumps({"e2e_step_ms": (time.perf_counter() - t) * 1e3}), flush=True)
