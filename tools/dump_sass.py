#!/usr/bin/env python
"""Generate + compile (NVRTC, no GPU needed) the kernel of a call and dump source / cubin for cuobjdump.
usage: dump_sass.py kind expr_a [expr_b] [cond] [mode] [opt=val ...]  -> /tmp/wdb_<kind>.cu, /tmp/wdb_<kind>.cubin"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from warpdb_b200 import _core as wc
args = [a for a in sys.argv[1:] if "=" not in a or a.startswith("(") or "[" in a]
opts = [a for a in sys.argv[1:] if a not in args]
for o in opts:
    k, v = o.split("=")
    wc.set_option(k, int(v))
kind = args[0]
a = args[1]
b = args[2] if len(args) > 2 and args[2] != "-" else None
cond = args[3] if len(args) > 3 and args[3] != "-" else None
mode = int(args[4]) if len(args) > 4 else 0
schema = [("price", wc.FLOAT32, 0, 0), ("quantity", wc.INT32, 0, 0)]
src, cubin = wc.debug_compile(kind, schema, a, b, cond, mode)
open(f"/tmp/wdb_{kind}.cu", "w").write(src)
open(f"/tmp/wdb_{kind}.cubin", "wb").write(cubin)
print(len(src), len(cubin))
