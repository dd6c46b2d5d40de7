"""CPU: the N>1 path (warpdb_b200/sharded.py: shard ranges, variable-length all_gather / all_to_all,
partial-aggregate merge, top-k candidate merge) under world_size-2 and -3 gloo groups."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return str(p)


@pytest.mark.parametrize("world,n", [(2, 20011), (3, 5000)])
def test_sharded_operators_under_gloo(world, n):
    port = free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "_gloo_worker.py"), str(r), str(world), port, str(n)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"rank {r} ok" in out, out[-3000:]


def test_shard_ranges_cover_rows_like_the_reference():
    from warpdb_b200.sharded import shard_range
    for n, w in [(10, 4), (4, 8), (1000003, 8), (0, 2), (7, 1)]:
        r = [shard_range(n, w, i) for i in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        chunk = (n + w - 1) // w
        assert all(e - s <= chunk for s, e in r)
