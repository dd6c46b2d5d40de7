"""Worker of tests/test_gpu_multi.py: one rank = one GPU, NCCL, the product's CUDA backend; results are
checked against the CPU oracle evaluated on the full (unsharded) table."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import pyoracle as orc  # noqa: E402
from warpdb_b200 import _core as wc  # noqa: E402
from warpdb_b200.sharded import ShardedDB, shard_range  # noqa: E402


def bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


def main():
    rank, world, port, n = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = port
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    wc.check(wc.lib().wdb_init(rank))
    wc.set_udf_source("__device__ float discount(float price, float rate) {\n    return price * rate;\n}\n")
    price = orc.synth_f32(n, 77, 0.0, 40.0)
    qty = orc.synth_i32(n, 78, -50, 1950)
    s, e = shard_range(n, world, rank)
    table = {"price": torch.from_numpy(price[s:e].copy()).cuda(rank), "quantity": torch.from_numpy(qty[s:e].copy()).cuda(rank)}
    full = {"price": price, "quantity": qty}
    db = ShardedDB(table, n, rank, world)

    ref, _ = orc.project_filter("price * quantity * 1.08", None, full)
    got = db.query("((price[idx] * quantity[idx]) * 1.08f)", None, gather=True).cpu().numpy()
    assert np.array_equal(bits(got), bits(ref))
    refc = orc.filter_compact("price * 0.9", "price > 20", full)
    loc, off, total = db.query_compact("(price[idx] * 0.9f)", "(price[idx] > 20.0f)")
    assert total == len(refc) and np.array_equal(bits(loc.cpu().numpy()), bits(refc[off:off + len(loc)]))
    for agg in (wc.SUM, wc.AVG, wc.COUNT, wc.MIN, wc.MAX):
        r = orc.group_agg("price", "quantity", "price > 20", full, agg=agg)
        for strat in ("core", "allgather", "exchange"):
            g = db.group_agg("price[idx]", "quantity[idx]", "(price[idx] > 20.0f)", agg=agg, expected_groups=2000, strategy=strat)
            assert np.array_equal(g["keys"].cpu().numpy(), r["keys"]), (agg, strat)
            if agg in (wc.COUNT, wc.MIN, wc.MAX):
                assert np.array_equal(bits(g["vals"].cpu().numpy()), bits(r["vals"])), (agg, strat)
            else:
                np.testing.assert_allclose(g["vals"].cpu().numpy(), r["vals"], rtol=1e-6)
    for desc in (True, False):
        want = orc.query_sql(f"SELECT price FROM t ORDER BY quantity {'DESC' if desc else 'ASC'} LIMIT 7 OFFSET 2", full)
        got = db.topk("quantity[idx]", "price[idx]", None, desc, 7, 2).cpu().numpy()
        assert np.array_equal(bits(got), bits(want)), (desc, got, want)
        want = orc.topk("discount(price, 0.9)", "price > 20", full, descending=desc, k=5)
        got = db.topk("discount(price[idx], 0.9f)", None, "(price[idx] > 20.0f)", desc, 5).cpu().numpy()
        assert np.array_equal(bits(got), bits(want))
    # ---- the merges inside the core, through the C ABI (wdb_multi_*): libwarpcore's own NCCL communicator
    from warpdb_b200 import ops
    comm = db._core()
    assert comm is not None and comm.world == world
    # sparse presence inside a wide range (keys are multiples of 3): the all-reduce must keep absent keys absent
    r = orc.group_agg("price", "quantity * 3", None, full, agg=orc.SUM)
    keys, vals = comm.group_agg(table, "price[idx]", "(quantity[idx] * 3)", None, wc.SUM, wc.ORDER_KEY_ASC, row_base=s)
    assert np.array_equal(keys.cpu().numpy(), r["keys"])
    np.testing.assert_allclose(vals.cpu().numpy(), r["vals"], rtol=1e-6)
    # a wide sparse range (direct-addressed table of ~2.6 M entries, 2000 of them present), descending
    r = orc.group_agg("price", "quantity * 1301", "price > 20", full, agg=orc.AVG, order=orc.ORDER_KEY_DESC)
    keys, vals = comm.group_agg(table, "price[idx]", "(quantity[idx] * 1301)", "(price[idx] > 20.0f)", wc.AVG, wc.ORDER_KEY_DESC, row_base=s)
    assert np.array_equal(keys.cpu().numpy(), r["keys"])
    np.testing.assert_allclose(vals.cpu().numpy(), r["vals"], rtol=1e-6)
    # the same with the table filled in three index slices: the all-reduce of a finished slice runs on a side stream
    # while the next slice's kernel runs (the 10 M-key configuration of the benchmark takes this path with two slices)
    wc.set_option("group.dense_passes", 3)
    for agg in (wc.AVG, wc.MAX):
        r = orc.group_agg("price", "quantity * 1301", "price > 20", full, agg=agg)
        keys, vals = comm.group_agg(table, "price[idx]", "(quantity[idx] * 1301)", "(price[idx] > 20.0f)", agg, wc.ORDER_KEY_ASC, row_base=s)
        assert np.array_equal(keys.cpu().numpy(), r["keys"]), agg
        np.testing.assert_allclose(vals.cpu().numpy(), r["vals"], rtol=1e-6)
    wc.set_option("group.dense_passes", None)
    for agg in (wc.MIN, wc.MAX, wc.COUNT):
        r = orc.group_agg("price", "quantity", "price > 20", full, agg=agg)
        keys, vals = comm.group_agg(table, "price[idx]", "quantity[idx]", "(price[idx] > 20.0f)", agg, wc.ORDER_KEY_ASC, row_base=s, key_range=(-50, 1949))
        assert np.array_equal(keys.cpu().numpy(), r["keys"]) and np.array_equal(bits(vals.cpu().numpy()), bits(r["vals"])), agg
    # first-appearance order (jit_group_sum's contract) takes the general path: ordered partials, all-gather, merge
    r = orc.group_agg("price", "quantity", None, full, agg=orc.SUM, order=orc.ORDER_FIRST)
    keys, vals = comm.group_agg(table, "price[idx]", "quantity[idx]", None, wc.SUM, wc.ORDER_FIRST, row_base=s, expected_groups=2000)
    assert np.array_equal(keys.cpu().numpy(), r["keys"])
    np.testing.assert_allclose(vals.cpu().numpy(), r["vals"], rtol=1e-6)
    # nothing survives anywhere
    keys, vals = comm.group_agg(table, "price[idx]", "quantity[idx]", "(price[idx] > 1000000.0f)", wc.SUM, wc.ORDER_KEY_ASC, row_base=s)
    assert keys.numel() == 0
    # large LIMIT (beyond the register top-k): threshold compaction per rank + one stable sort of the candidates
    for desc in (True, False):
        want = orc.query_sql(f"SELECT price FROM t WHERE price > 20 ORDER BY quantity {'DESC' if desc else 'ASC'} LIMIT 300 OFFSET 11", full)
        got = comm.topk(table, "quantity[idx]", "price[idx]", "(price[idx] > 20.0f)", desc, 300, 11, row_base=s).cpu().numpy()
        assert np.array_equal(bits(got), bits(want)), desc
    # compaction: per-rank survivors, global offset and total from one all-gather of the counts
    out, (cnt, off, total) = comm.project_filter(table, "(price[idx] * 0.9f)", "(price[idx] > 20.0f)", wc.COMPACT)
    assert total == len(refc) and np.array_equal(bits(out[:cnt].cpu().numpy()), bits(refc[off:off + cnt]))
    dist.barrier()
    # single-process multi-device host entry point (run_multi_gpu_jit_host replacement), rank 0 only
    if rank == 0:
        import ctypes as C
        hp = torch.from_numpy(price).pin_memory(); hq = torch.from_numpy(qty).pin_memory()
        ho = torch.empty(n, dtype=torch.float32).pin_memory()
        cols, nc = wc.make_cols([("price", wc.FLOAT32, hp.data_ptr(), n), ("quantity", wc.INT32, hq.data_ptr(), n)])
        cnt = C.c_int64(0)
        wc.check(wc.lib().wdb_multi_project_filter_host(world, None, cols, nc, b"((price[idx] * quantity[idx]) * 1.08f)", b"",
                                                        ho.data_ptr(), n, wc.DENSE_ZERO, C.byref(cnt)))
        assert np.array_equal(bits(ho.numpy()), bits(ref))
        wc.check(wc.lib().wdb_multi_project_filter_host(world, None, cols, nc, b"(price[idx] * 0.9f)", b"(price[idx] > 20.0f)",
                                                        ho.data_ptr(), n, wc.COMPACT, C.byref(cnt)))
        assert cnt.value == len(refc) and np.array_equal(bits(ho.numpy()[:cnt.value]), bits(refc))
        # ... and its aggregate / ORDER BY counterparts: one process, one thread per GPU, ncclCommInitAll, merges over NVLink
        r = orc.group_agg("price", "quantity", "price > 20", full, agg=orc.SUM)
        hk = np.empty(4096, np.int32); hv = np.empty(4096, np.float32); g = C.c_int64(0)
        wc.check(wc.lib().wdb_multi_group_agg_host(world, None, cols, nc, b"price[idx]", b"quantity[idx]", b"(price[idx] > 20.0f)", wc.SUM, wc.ORDER_KEY_ASC,
                                                   n, 0, hk.ctypes.data, hv.ctypes.data, 4096, C.byref(g)))
        assert g.value == len(r["keys"]) and np.array_equal(hk[:g.value], r["keys"])
        np.testing.assert_allclose(hv[:g.value], r["vals"], rtol=1e-6)
        want = orc.topk("discount(price, 0.9)", "price > 20", full, descending=True, k=5)
        hv5 = np.empty(5, np.float32); m = C.c_int64(0)
        wc.check(wc.lib().wdb_multi_topk_host(world, None, cols, nc, b"discount(price[idx], 0.9f)", None, b"(price[idx] > 20.0f)", 1, 5, 0, n,
                                              hv5.ctypes.data, C.byref(m)))
        assert m.value == 5 and np.array_equal(bits(hv5), bits(want))
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
