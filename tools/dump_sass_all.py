#!/usr/bin/env python
"""SASS of the default instantiation of every hot kernel -> profiles/sass/<kernel>.sass (no GPU needed:
NVRTC + cuobjdump).  Each file holds the one function, preceded by a histogram of its memory / vote /
shuffle / atomic mnemonics so the claims of DESIGN.md (256-bit LDG/STG, L2 eviction hints, native
RED.ADD.F64, no shared-memory atomics in wdb_group_wp, ...) can be checked at a glance."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from warpdb_b200 import _core as wc  # noqa: E402

OUT = os.path.join(ROOT, "profiles", "sass")
os.makedirs(OUT, exist_ok=True)
SCHEMA = [("price", wc.FLOAT32, 0, 0), ("quantity", wc.INT32, 0, 0)]
UDF = "__device__ float discount(float price, float rate) {\n    return price * rate;\n}\n"
wc.set_udf_source(UDF)
JOBS = [
    # (file, kind, expr_a, expr_b, cond, mode, options, function)
    ("wdb_project", "project", "((price[idx] * quantity[idx]) * 1.08f)", None, None, wc.DENSE, {}, "wdb_project"),
    ("wdb_compact_l2", "compact", "(price[idx] * 0.9f)", None, "(price[idx] > 20.0f)", 0, {}, "wdb_compact_l2"),
    ("wdb_compact_sp", "compact", "(price[idx] * 0.9f)", None, "(price[idx] > 20.0f)", 0, {"compact.variant": 4}, "wdb_compact_sp"),
    ("wdb_count_stage", "compact", "(price[idx] * 0.9f)", None, "(price[idx] > 20.0f)", 0, {"compact.variant": 5}, "wdb_count_stage"),
    ("wdb_gather_stage", "compact", "(price[idx] * 0.9f)", None, "(price[idx] > 20.0f)", 0, {"compact.variant": 5}, "wdb_gather_stage"),
    ("wdb_group_wp_lane_private", "group", "price[idx]", "quantity[idx]", None, wc.SUM, {"group.debug_span": 10}, "wdb_group_wp"),
    ("wdb_group_wp", "group", "price[idx]", "quantity[idx]", None, wc.SUM, {"group.debug_span": 1000}, "wdb_group_wp"),
    ("wdb_group_dense", "group", "price[idx]", "quantity[idx]", None, wc.SUM, {"group.debug_span": 10_000_000}, "wdb_group"),
    ("wdb_group_hash", "group", "price[idx]", "quantity[idx]", None, wc.SUM, {}, "wdb_group"),
    ("wdb_topk_scan", "topk", "discount(price[idx], 0.9f)", "discount(price[idx], 0.9f)", None, 1, {}, "wdb_topk_scan"),
]
INTEREST = re.compile(r"\b(LDG|STG|LDS|STS|ATOMS|ATOMG|ATOM|RED|REDG|REDUX|VOTE|VOTEU|SHFL|MATCH|BAR|UBLKCP|CCTL|DADD|FMUL|FFMA|FSETP|POPC)[.\w]*")
for name, kind, a, b, cond, mode, opts, fn in JOBS:
    for k, v in opts.items():
        wc.set_option(k, v)
    try:
        _, cubin = wc.debug_compile(kind, SCHEMA, a, b, cond, mode)
    finally:
        for k in opts:
            wc.set_option(k, None)
    tmp = f"/tmp/{name}.cubin"
    open(tmp, "wb").write(cubin)
    sass = subprocess.run(["cuobjdump", "-sass", tmp], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", tmp], capture_output=True, text=True, check=True).stdout
    part = sass.split("Function : " + fn + "\n")[1].split("Function : ")[0]
    # drop the encoding columns: keep address + instruction
    lines = []
    for ln in part.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            lines.append(f"/*{m.group(1)}*/ {m.group(2).strip()} ;")
    hist = collections.Counter(m.group(0) for ln in lines for m in [INTEREST.search(ln)] if m)
    usage = [l.strip() for i, l in enumerate(res.splitlines()) if ("Function " + fn + ":") in l or (i and ("Function " + fn + ":") in res.splitlines()[i - 1])]
    with open(os.path.join(OUT, name + ".sass"), "w") as f:
        f.write(f"// {fn}: kind={kind} expr={a!r} key/expr2={b!r} cond={cond!r} options={opts}  (sm_100a, NVRTC 12.9, cuobjdump -sass)\n")
        f.write("// " + " | ".join(usage) + "\n")
        f.write("// mnemonic histogram: " + ", ".join(f"{k} x{v}" for k, v in sorted(hist.items())) + "\n")
        f.write("\n".join(lines) + "\n")
    print(name, len(lines), "instructions")

# ---- kernels compiled ahead of time into libwarpcore.so (nvcc): the join's probe kernels ------------
LIB = os.path.join(ROOT, "warpdb_b200", "libwarpcore.so")
STATIC = [("join_count_kernel_i32_i32", "join_count_kernelIiiEE"), ("join_emit_kernel_i32_i32", "join_emit_kernelIiiEE")]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout.splitlines()
for name, tag in STATIC:
    chunk = [c for c in sass.split("Function : ")[1:] if tag in c.split("\n", 1)[0]][0]
    mangled = chunk.split("\n", 1)[0].strip()
    lines = []
    for ln in chunk.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            lines.append(f"/*{m.group(1)}*/ {m.group(2).strip()} ;")
    hist = collections.Counter(m.group(0) for ln in lines for m in [INTEREST.search(ln)] if m)
    usage = [res[i + 1].strip() for i, l in enumerate(res) if mangled in l and i + 1 < len(res)]
    with open(os.path.join(OUT, name + ".sass"), "w") as f:
        f.write(f"// {mangled}  (warpdb_b200/csrc/ops_join.cu, nvcc 12.9 -gencode arch=compute_100a,code=sm_100a, cuobjdump -sass)\n")
        f.write("// " + " | ".join(usage) + "\n")
        f.write("// mnemonic histogram: " + ", ".join(f"{k} x{v}" for k, v in sorted(hist.items())) + "\n")
        f.write("\n".join(lines) + "\n")
