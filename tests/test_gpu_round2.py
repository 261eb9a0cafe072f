"""Round-2 GPU tests: compressed HOST staging, keyless partial merge, 64-bit aggregate normalisation and the peer-mapped
(NVLink) collectives — all through the C-ABI, all against the CPU oracle."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from lingodb_b200 import datagen

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tpch(ctx, t):
    from lingodb_b200 import runtime
    return runtime.Tpch(ctx, {k: ctx.table_from_host(v) for k, v in t.items()})


def test_compressed_staging_gives_the_oracles_rows(gpu_ctx, oracle):
    """HOST batches of >= 64 Ki rows go through pack (host) → copy → unpack kernel; ragged lengths, every query shape."""
    from lingodb_b200 import capi
    L = gpu_ctx.L
    t = datagen.tpch(0.1, seed=5, chunk_rows=200_003, with_parts=True)  # 600 K lineitem rows in ragged >64Ki batches
    oh = {k: oracle.table(v) for k, v in t.items()}
    h2d0 = int(L.ldb_gpu_context_h2d_bytes(gpu_ctx.h))
    g = _tpch(gpu_ctx, t)
    assert g.q1() == oracle.q1(oh["lineitem"])[0]
    staged = int(L.ldb_gpu_context_h2d_bytes(gpu_ctx.h)) - h2d0
    arrow = sum(v.nbytes if not isinstance(v, tuple) else v[0].nbytes + v[1].nbytes for tab in t.values() for ch in tab.chunks for v in ch.values())
    assert staged < 0.45 * arrow, (staged, arrow)  # the packed bytes that crossed the link, not the Arrow bytes
    assert g.q6() == oracle.q6(oh["lineitem"])[0]
    assert g.q3() == oracle.q3(oh["customer"], oh["orders"], oh["lineitem"])[0]
    assert g.q5() == oracle.q5(oh["customer"], oh["orders"], oh["lineitem"], oh["supplier"], oh["nation"], oh["region"])[0]
    assert g.q9() == oracle.q9(oh["part"], oh["supplier"], oh["lineitem"], oh["partsupp"], oh["orders"], oh["nation"])[0]
    # restaging after a clear reuses the staging buffers (generation handshake with the workers)
    li = g.tables["lineitem"]
    for _ in range(2):
        li.clear()
        for chunk, n in zip(t["lineitem"].chunks, t["lineitem"].chunk_rows):
            li.append_host(chunk, n)
        assert g.q1() == oracle.q1(oh["lineitem"])[0]


def test_compressed_staging_handles_negative_and_wide_values(gpu_ctx, oracle):
    """Blocks whose range needs 8 bytes per value, negative decimals, int32 extremes: unpack must reproduce every cell."""
    from lingodb_b200 import runtime
    rng = np.random.default_rng(3)
    n = 70_001
    specs = [c for c in datagen.LINEITEM_SCHEMA if c.name in ("l_quantity", "l_extendedprice", "l_discount", "l_shipdate")]
    qty = rng.integers(-2**40, 2**40, n, dtype=np.int64)
    ext = rng.integers(-10**7, 10**7, n, dtype=np.int64)
    disc = rng.integers(0, 11, n, dtype=np.int64)

    def dec(v):
        a = np.zeros((n, 2), np.int64)
        a[:, 0] = v
        a[:, 1] = v >> 63
        return a.view(np.uint8).reshape(n, 16)

    chunk = {"l_quantity": dec(qty), "l_extendedprice": dec(ext), "l_discount": dec(disc), "l_shipdate": rng.integers(-2**31, 2**31 - 1, n, dtype=np.int64).astype(np.int32)}
    td = datagen.TableData("lineitem", specs)
    td.chunks.append(chunk)
    td.chunk_rows.append(n)
    tab = gpu_ctx.table_from_host(td)
    from lingodb_b200 import capi
    s2, e = C.c_void_p(), capi.Error()
    capi.check(gpu_ctx.L.ldb_gpu_simple_state_create(gpu_ctx.h, 2, C.byref(s2), C.byref(e)), e)
    runtime.run_pipeline(gpu_ctx, "scan_reduce", tab, aggs=[("col", ["l_quantity"]), ("one", [])], sink=s2)
    out = (capi.I128 * 8)()
    capi.check(gpu_ctx.L.ldb_gpu_simple_state_read(s2, out, C.byref(e)), e)
    want = int(qty.sum())
    assert out[0].value() == want and out[0].hi == -1  # negative i64 sums come back sign-extended
    assert out[1].value() == n
    runtime.state_destroy(gpu_ctx, s2)
    s3 = C.c_void_p()
    capi.check(gpu_ctx.L.ldb_gpu_simple_state_create(gpu_ctx.h, 1, C.byref(s3), C.byref(e)), e)
    runtime.run_pipeline(gpu_ctx, "scan_reduce", tab, aggs=[("mul", ["l_extendedprice", "l_discount"])], sink=s3)
    capi.check(gpu_ctx.L.ldb_gpu_simple_state_read(s3, out, C.byref(e)), e)
    assert out[0].value() == int((ext.astype(object) * disc.astype(object)).sum())
    runtime.state_destroy(gpu_ctx, s3)


def test_keyless_partials_export_and_merge(gpu_ctx, oracle):
    """Two lineitem shards → two keyless Q6 partial states → export / merge_exported == the oracle over the whole table."""
    import torch
    from lingodb_b200 import capi, runtime
    t = datagen.tpch(0.05, seed=21, chunk_rows=50_000)
    whole = oracle.q6(oracle.table(t["lineitem"]))[0]
    li = t["lineitem"]
    half = len(li.chunks) // 2
    states = []
    for part in (range(0, half), range(half, len(li.chunks))):
        td = datagen.TableData("lineitem", li.columns)
        for i in part:
            td.chunks.append(li.chunks[i])
            td.chunk_rows.append(li.chunk_rows[i])
        tp = runtime.Tpch(gpu_ctx, {"lineitem": gpu_ctx.table_from_host(td)})
        s, e = C.c_void_p(), capi.Error()
        capi.check(gpu_ctx.L.ldb_tpch_q6_partial(gpu_ctx.h, C.byref(tp.t), b"1994-01-01", b"1995-01-01", b"0.05", b"0.07", 24, C.byref(s), C.byref(e)), e)
        states.append(s)
    L = gpu_ctx.L
    nbytes = int(L.ldb_gpu_groupby_export_bytes(states[0]))
    recv = torch.empty(2 * nbytes, dtype=torch.uint8, device=torch.device("cuda", gpu_ctx.device))
    e = capi.Error()
    for r in range(2):
        capi.check(L.ldb_gpu_groupby_export(states[r], C.c_void_p(recv.data_ptr() + r * nbytes), C.byref(e)), e)
    gpu_ctx.synchronize()
    capi.check(L.ldb_gpu_groupby_merge_exported(states[0], C.c_void_p(recv.data_ptr()), 2, 0, C.byref(e)), e)
    out = (capi.I128 * 8)()
    capi.check(L.ldb_gpu_simple_state_read(states[0], out, C.byref(e)), e)
    assert {"revenue": out[0].value()} == whole
    for s in states:
        runtime.state_destroy(gpu_ctx, s)


def _shard_tables(ctx, s, cols, r, world):
    from lingodb_b200 import devgen, parallel
    o_lo, o_hi, r_lo, r_hi = parallel.order_range(s, r, world)
    return {"lineitem": devgen.lineitem(ctx, s, cols, row_begin=r_lo, n_rows=r_hi - r_lo)}


def test_peer_allmerge_protocol_on_one_device(oracle):
    """The peer-mapped all-merge (csrc/peer.cu) with 3 'ranks' that are 3 contexts of this process on device 0: same kernels,
    flags and mailboxes as the multi-process NVLink path, peers wired by pointer.  Every rank must end with the whole result,
    twice in a row (epoch parity), for Q1 (4 groups) and the keyless Q6."""
    from lingodb_b200 import capi, parallel, runtime
    world = 3
    s = datagen.scale(0.05, seed=13)
    cols = ["l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_returnflag", "l_linestatus", "l_shipdate"]
    host = datagen.lineitem(s, cols)
    oh = oracle.table(host)
    want_q1, want_q6 = oracle.q1(oh)[0], oracle.q6(oh)[0]
    ctxs = [runtime.Context(0) for _ in range(world)]
    try:
        comms = parallel.Comm.local_group(ctxs)
        tps = [runtime.Tpch(c, _shard_tables(c, s, cols, r, world)) for r, c in enumerate(ctxs)]
        for _ in range(2):
            sts = [tp.q1_partial() for tp in tps]
            for cm, st in zip(comms, sts):  # enqueue on every rank before anyone synchronises: the kernels wait for each other
                cm.allmerge(st)
            for r in range(world):
                assert tps[r].q1_finish(sts[r]) == want_q1, f"rank {r}"
                runtime.state_destroy(ctxs[r], sts[r])
            sts = []
            for c, tp in zip(ctxs, tps):
                st, e = C.c_void_p(), capi.Error()
                capi.check(c.L.ldb_tpch_q6_partial(c.h, C.byref(tp.t), b"1994-01-01", b"1995-01-01", b"0.05", b"0.07", 24, C.byref(st), C.byref(e)), e)
                sts.append(st)
            for cm, st in zip(comms, sts):
                cm.allmerge(st)
            for r in range(world):
                out, e = (capi.I128 * 8)(), capi.Error()
                capi.check(ctxs[r].L.ldb_gpu_simple_state_read(sts[r], out, C.byref(e)), e)
                assert {"revenue": out[0].value()} == want_q6, f"rank {r}"
                runtime.state_destroy(ctxs[r], sts[r])
        for cm in comms:
            cm.barrier()
        for cm in comms:
            cm.check()
        for cm in comms:
            cm.close()
    finally:
        for c in ctxs:
            c.close()


def test_captured_query_replays_bit_identically(gpu_ctx, oracle):
    """ldb_gpu_graph_*: Q1 captured once (state init + pipeline) and replayed; every replay re-initialises the state and gives
    the oracle's rows; the per-kernel timers keep working through the graph's external event nodes."""
    from lingodb_b200 import devgen, runtime
    s = datagen.scale(0.05, seed=23)
    cols = ["l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_returnflag", "l_linestatus", "l_shipdate"]
    want = oracle.q1(oracle.table(datagen.lineitem(s, cols)))[0]
    tp = runtime.Tpch(gpu_ctx, {"lineitem": devgen.lineitem(gpu_ctx, s, cols)})
    assert tp.q1() == want  # eager first (warms the allocation pool)
    gpu_ctx.graph_begin()
    st = tp.q1_partial()
    g = gpu_ctx.graph_end()
    gpu_ctx.kernel_time_reset(True)
    for _ in range(3):
        g.launch()
        assert tp.q1_finish(st) == want
    ms, n = gpu_ctx.kernel_time("scan_groupby")
    gpu_ctx.kernel_time_reset(False)
    assert n == 3 and ms > 0
    g.destroy()
    runtime.state_destroy(gpu_ctx, st)
    # a HOST-staged table can be captured too: the capture waits for its staging on the host instead of on the stream
    tph = runtime.Tpch(gpu_ctx, {"lineitem": gpu_ctx.table_from_host(datagen.lineitem(s, cols, chunk_rows=70_000))})
    gpu_ctx.graph_begin()
    st = tph.q1_partial()
    g = gpu_ctx.graph_end()
    for _ in range(2):
        g.launch()
        assert tph.q1_finish(st) == want
    g.destroy()
    runtime.state_destroy(gpu_ctx, st)


def test_captured_query_with_peer_allmerge(oracle):
    """The all-merge kernel reads its epoch from device memory, so a captured (scan + all-merge) query replays correctly on
    every rank: 2 ranks = 2 contexts of this process on device 0."""
    from lingodb_b200 import parallel, runtime
    world = 2
    s = datagen.scale(0.05, seed=29)
    cols = ["l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_returnflag", "l_linestatus", "l_shipdate"]
    want = oracle.q1(oracle.table(datagen.lineitem(s, cols)))[0]
    ctxs = [runtime.Context(0) for _ in range(world)]
    try:
        comms = parallel.Comm.local_group(ctxs)
        tps = [runtime.Tpch(c, _shard_tables(c, s, cols, r, world)) for r, c in enumerate(ctxs)]
        graphs, states = [], []
        for c, cm, tp in zip(ctxs, comms, tps):
            tp.q1()  # warm the pools before capturing
            c.graph_begin()
            st = tp.q1_partial()
            cm.allmerge(st)
            graphs.append(c.graph_end())
            states.append(st)
        for _ in range(3):
            for g in graphs:
                g.launch()
            for tp, st in zip(tps, states):
                assert tp.q1_finish(st) == want
        for cm in comms:
            cm.barrier()  # eager collectives still line up with the replayed ones
        for cm in comms:
            cm.check()
        for g, c, st in zip(graphs, ctxs, states):
            g.destroy()
            runtime.state_destroy(c, st)
        for cm in comms:
            cm.close()
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("world", [1, 2, 3])
def test_q5_repartitioned_peer_stores_match_the_oracle(oracle, world):
    """ldb_tpch_q5_repartitioned: fused partition + peer store, device barriers, Bloom OR by peer loads, all-merge — `world` ranks
    as contexts of this process on device 0, one host thread per rank (the drivers block on their peers)."""
    import threading
    from lingodb_b200 import devgen, parallel, runtime
    s = datagen.scale(0.2, seed=31)
    host = datagen.tpch(0.2, seed=31)
    oh = {k: oracle.table(v) for k, v in host.items()}
    want = oracle.q5(oh["customer"], oh["orders"], oh["lineitem"], oh["supplier"], oh["nation"], oh["region"])[0]
    cols = ["l_orderkey", "l_suppkey", "l_extendedprice", "l_discount"]
    ctxs = [runtime.Context(0) for _ in range(world)]
    try:
        # (ranks of one process share the CUDA context: the world=1 case, which runs first, has loaded every kernel of the flow —
        #  with lazy loading one rank's first launch of a kernel would wait for another rank's spinning barrier kernel)
        heap = parallel.q5_heap_bytes(ctxs[0], s.n_orders, s.n_lineitem, world)
        comms = parallel.Comm.local_group(ctxs, user_bytes=heap)
        tps = []
        for r, c in enumerate(ctxs):
            o_lo, o_hi, r_lo, r_hi = parallel.order_range(s, r, world)
            # lineitem rows of OTHER orders than this rank's order shard: the join is not co-partitioned, the shuffle has to do it
            rr = (r + 1) % world
            _, _, l_lo, l_hi = parallel.order_range(s, rr, world)
            tps.append(runtime.Tpch(c, {"lineitem": devgen.lineitem(c, s, cols, row_begin=l_lo, n_rows=l_hi - l_lo), "orders": devgen.orders(c, s, row_begin=o_lo, n_rows=o_hi - o_lo),
                                        "customer": devgen.customer(c, s), "supplier": devgen.supplier(c, s), **devgen.small_tables(c)}))
        for it in range(2):
            res, errs = [None] * world, []

            def run(r):
                try:
                    res[r] = parallel.q5_repartitioned_peer(ctxs[r], tps[r], comms[r], s.n_orders, s.n_lineitem)
                except Exception as ex:  # noqa: BLE001
                    errs.append((r, str(ex)))
            ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
            [t.start() for t in ts]
            [t.join() for t in ts]
            assert not errs, "\n".join(f"rank {r}: {m}" for r, m in errs)
            for r in range(world):
                assert res[r][0] == want, f"rank {r} iteration {it}"
            assert sum(res[r][1]["orders_tuples_sent"] for r in range(world)) == sum(res[r][1]["orders_tuples_received"] for r in range(world)) > 0
            assert sum(res[r][1]["lineitem_tuples_sent"] for r in range(world)) == sum(res[r][1]["lineitem_tuples_received"] for r in range(world)) > 0
        for cm in comms:
            cm.close()
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("world", [1, 3])
def test_q9_repartitioned_orders_hash_partitioned_match_the_oracle(oracle, world):
    """ldb_tpch_q9_repartitioned: {o_orderkey, year} tuples to the owner of h64(o_orderkey) (K10 with the year expression), lineitem
    contributions {l_orderkey | nation, i128 amount} shipped after the partsupp / supplier probes (K11), probe of the received tuples
    against the orders partition, all-merge — ranks as contexts of this process, lineitem shards deliberately NOT co-partitioned."""
    import threading
    from lingodb_b200 import devgen, parallel, runtime
    s = datagen.scale(0.1, seed=37)
    cols = ["l_orderkey", "l_partkey", "l_suppkey", "l_quantity", "l_extendedprice", "l_discount"]
    host = datagen.tpch(0.1, seed=37, lineitem_columns=cols, with_parts=True)
    oh = {k: oracle.table(v) for k, v in host.items()}
    want = oracle.q9(oh["part"], oh["supplier"], oh["lineitem"], oh["partsupp"], oh["orders"], oh["nation"])[0]
    ctxs = [runtime.Context(0) for _ in range(world)]
    try:
        comms = parallel.Comm.local_group(ctxs, user_bytes=parallel.q9_heap_bytes(ctxs[0], s.n_orders, s.n_lineitem, world))
        tps = []
        for r, c in enumerate(ctxs):
            o_lo, o_hi, _, _ = parallel.order_range(s, r, world)
            _, _, l_lo, l_hi = parallel.order_range(s, (r + 1) % world, world)
            tps.append(runtime.Tpch(c, {"lineitem": devgen.lineitem(c, s, cols, row_begin=l_lo, n_rows=l_hi - l_lo), "orders": devgen.orders(c, s, row_begin=o_lo, n_rows=o_hi - o_lo),
                                        "supplier": devgen.supplier(c, s), "part": devgen.part(c, s), "partsupp": devgen.partsupp(c, s), **devgen.small_tables(c)}))
        for it in range(2):
            res, errs = [None] * world, []

            def run(r):
                try:
                    res[r] = parallel.q9_repartitioned_peer(ctxs[r], tps[r], comms[r], s.n_orders, s.n_lineitem)
                except Exception as ex:  # noqa: BLE001
                    errs.append((r, str(ex)))
            ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
            [t.start() for t in ts]
            [t.join() for t in ts]
            assert not errs, "\n".join(f"rank {r}: {m}" for r, m in errs)
            for r in range(world):
                assert res[r][0] == want, f"rank {r} iteration {it}"
            assert sum(res[r][1]["orders_tuples_sent"] for r in range(world)) == s.n_orders  # every order is shipped to its owner
            assert sum(res[r][1]["lineitem_tuples_sent"] for r in range(world)) == sum(res[r][1]["lineitem_tuples_received"] for r in range(world)) > 0
        for cm in comms:
            cm.close()
    finally:
        for c in ctxs:
            c.close()


def test_peer_collectives_across_processes_and_gpus():
    """world = 2 processes on 2 GPUs (CUDA IPC + NVLink P2P), when the box has them: tools/peer_selftest.py checks Q1/Q9
    against the oracle on every rank and exits non-zero on any mismatch."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29591",
                        os.path.join(ROOT, "tools", "peer_selftest.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "peer selftest ok" in r.stdout


def test_filter_shape_instantiations_agree_with_the_descriptor_driven_kernels(gpu_ctx, oracle):
    """The join pipelines run an instantiation compiled for their filter shape (none / one int32 compare / one int32 range —
    kernels.cu FilterShape) by default; the descriptor-driven form must give the same rows, and both the oracle's."""
    from lingodb_b200 import runtime
    t = datagen.tpch(0.05, seed=11, chunk_rows=40_003, with_parts=True)
    g = runtime.Tpch(gpu_ctx, {k: gpu_ctx.table_from_host(v) for k, v in t.items()})
    oh = {k: oracle.table(v) for k, v in t.items()}
    res = {}
    try:
        for on in (0, 1):
            gpu_ctx.L.ldb_gpu_set_filter_specialisation(on)
            res[on] = (g.q3(), g.q3("MACHINERY", "1996-01-01"), g.q5(), g.q5("EUROPE", "1996-01-01", "1997-01-01"), g.q9("green"))
    finally:
        gpu_ctx.L.ldb_gpu_set_filter_specialisation(1)
    assert res[0] == res[1]
    assert res[1][0] == oracle.q3(oh["customer"], oh["orders"], oh["lineitem"])[0]
    assert res[1][2] == oracle.q5(oh["customer"], oh["orders"], oh["lineitem"], oh["supplier"], oh["nation"], oh["region"])[0]
    assert res[1][4] == oracle.q9(oh["part"], oh["supplier"], oh["lineitem"], oh["partsupp"], oh["orders"], oh["nation"], "green")[0]
