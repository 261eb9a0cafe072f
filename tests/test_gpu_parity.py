"""GPU parity tests proper: the CUDA path through the C-ABI vs the CPU oracle, bit-exact."""
import numpy as np
import pytest

from lingodb_b200 import datagen

pytestmark = pytest.mark.gpu


def _oracle_tables(oracle, t):
    return {k: oracle.table(v) for k, v in t.items()}


def _gpu_tables(ctx, t):
    from lingodb_b200 import runtime
    return runtime.Tpch(ctx, {k: ctx.table_from_host(v) for k, v in t.items()})


@pytest.fixture(scope="module")
def small(oracle, gpu_ctx):
    t = datagen.tpch(0.05, seed=42, chunk_rows=8191)  # ragged Arrow-CSV-like batches
    return t, _oracle_tables(oracle, t), _gpu_tables(gpu_ctx, t)


def test_device_info(gpu_ctx):
    info = gpu_ctx.info()
    assert info["cc"][0] >= 10, info
    assert info["sm_count"] > 0


def test_hash_kat_on_device(gpu_ctx):
    # test/lit/DB/hash.mlir:27-34, test/unittests/storage/TestStorage.cpp:289
    h = gpu_ctx.hash_i64(np.array([10, -1, 1], dtype=np.int64))
    assert int(h[0]) == 9003023063795233148
    assert int(h[1]) == 14576801547736533962
    assert int(h[2]) == (-3797884931935089717) % 2**64
    # decimal<15,2> 100.01 = combine(h64(low=10001), h64(high=0))
    h2 = gpu_ctx.hash_i64(np.array([0], dtype=np.int64), np.array([10001], dtype=np.int64))
    assert int(h2[0]) == 5768746606534069840


def test_hash_matches_oracle_random(gpu_ctx, oracle):
    rng = np.random.default_rng(1)
    a = rng.integers(-2**63, 2**63 - 1, size=4096, dtype=np.int64)
    b = rng.integers(-2**63, 2**63 - 1, size=4096, dtype=np.int64)
    got = gpu_ctx.hash_i64(a, b)
    for i in range(0, 4096, 37):
        want = oracle.lib.oracle_hash_combine(oracle.lib.oracle_hash_i64(int(b[i])), oracle.lib.oracle_hash_i64(int(a[i])))
        assert int(got[i]) == want


def test_q6(small, oracle):
    t, oh, g = small
    want, _ = oracle.q6(oh["lineitem"])
    assert g.q6() == want
    # other constants, including an empty result
    want, _ = oracle.q6(oh["lineitem"], "1993-01-01", "1993-02-01", "0.00", "0.10", 51)
    assert g.q6("1993-01-01", "1993-02-01", "0.00", "0.10", 51) == want
    assert g.q6("2001-01-01", "2002-01-01") == {"revenue": 0}


def test_q1(small, oracle):
    t, oh, g = small
    want, _ = oracle.q1(oh["lineitem"])
    got = g.q1()
    assert got == want
    assert len(got) == 4
    for date in ("1995-06-17", "1992-01-02", "1991-01-01"):
        want, _ = oracle.q1(oh["lineitem"], date)
        assert g.q1(date) == want


def test_q3(small, oracle):
    t, oh, g = small
    want, _ = oracle.q3(oh["customer"], oh["orders"], oh["lineitem"])
    got = g.q3()
    assert got == want
    assert len(got) == 10
    want, _ = oracle.q3(oh["customer"], oh["orders"], oh["lineitem"], "MACHINERY", "1996-01-01")
    assert g.q3("MACHINERY", "1996-01-01") == want
    assert g.q3("NOSUCHSEG", "1995-03-15") == []


def test_q5(small, oracle):
    t, oh, g = small
    want, _ = oracle.q5(oh["customer"], oh["orders"], oh["lineitem"], oh["supplier"], oh["nation"], oh["region"])
    got = g.q5()
    assert got == want
    assert len(got) == 5
    want, _ = oracle.q5(oh["customer"], oh["orders"], oh["lineitem"], oh["supplier"], oh["nation"], oh["region"], "EUROPE", "1996-01-01", "1997-01-01")
    assert g.q5("EUROPE", "1996-01-01", "1997-01-01") == want


def test_device_generator_matches_host(gpu_ctx):
    from lingodb_b200 import devgen
    s = datagen.scale(0.02, seed=7)
    cols = [c.name for c in datagen.LINEITEM_SCHEMA]
    dev = devgen.to_host(devgen.lineitem(gpu_ctx, s, cols, batch_rows=50000))
    host = datagen.lineitem(s, chunk_rows=50000)
    assert dev.chunk_rows == host.chunk_rows
    for dc, hc in zip(dev.chunks, host.chunks):
        for name in cols:
            assert np.array_equal(dc[name], hc[name]), name
    do, ho = devgen.to_host(devgen.orders(gpu_ctx, s)), datagen.orders(s)
    for name in ("o_orderkey", "o_custkey", "o_orderdate", "o_shippriority"):
        assert np.array_equal(do.chunks[0][name], ho.chunks[0][name]), name
    dc, hc = devgen.to_host(devgen.customer(gpu_ctx, s)), datagen.customer(s)
    assert np.array_equal(dc.chunks[0]["c_custkey"], hc.chunks[0]["c_custkey"])
    assert np.array_equal(dc.chunks[0]["c_nationkey"], hc.chunks[0]["c_nationkey"])
    assert np.array_equal(dc.chunks[0]["c_mktsegment"][0], hc.chunks[0]["c_mktsegment"][0])
    assert np.array_equal(dc.chunks[0]["c_mktsegment"][1], hc.chunks[0]["c_mktsegment"][1])
    ds, hs = devgen.to_host(devgen.supplier(gpu_ctx, s)), datagen.supplier(s)
    assert np.array_equal(ds.chunks[0]["s_nationkey"], hs.chunks[0]["s_nationkey"])


def test_device_resident_tables_all_queries(gpu_ctx, oracle):
    """Tables generated in HBM (the bench path), one big batch, vs the oracle on the host twin."""
    from lingodb_b200 import devgen, runtime
    s = datagen.scale(0.1, seed=3)
    cols = ["l_orderkey", "l_suppkey", "l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_returnflag", "l_linestatus", "l_shipdate"]
    tabs = {"lineitem": devgen.lineitem(gpu_ctx, s, cols), "orders": devgen.orders(gpu_ctx, s), "customer": devgen.customer(gpu_ctx, s),
            "supplier": devgen.supplier(gpu_ctx, s), **devgen.small_tables(gpu_ctx)}
    g = runtime.Tpch(gpu_ctx, tabs)
    host = datagen.tpch(0.1, seed=3, lineitem_columns=cols)
    oh = {k: oracle.table(v) for k, v in host.items()}
    assert g.q6() == oracle.q6(oh["lineitem"])[0]
    assert g.q1() == oracle.q1(oh["lineitem"])[0]
    assert g.q3() == oracle.q3(oh["customer"], oh["orders"], oh["lineitem"])[0]
    assert g.q5() == oracle.q5(oh["customer"], oh["orders"], oh["lineitem"], oh["supplier"], oh["nation"], oh["region"])[0]


def test_empty_and_tiny_tables(gpu_ctx, oracle):
    from lingodb_b200 import runtime
    s = datagen.scale(0.001, seed=5)
    t = {"lineitem": datagen.lineitem(s, chunk_rows=1000, n_rows=0)}
    g = runtime.Tpch(gpu_ctx, {"lineitem": gpu_ctx.table_from_host(t["lineitem"])})
    assert g.q1() == []
    assert g.q6() == {"revenue": 0}
    t = datagen.tpch(0.001, seed=5, chunk_rows=333)
    g = _gpu_tables(gpu_ctx, t)
    oh = _oracle_tables(oracle, t)
    assert g.q1() == oracle.q1(oh["lineitem"])[0]
    assert g.q3() == oracle.q3(oh["customer"], oh["orders"], oh["lineitem"])[0]
    assert g.q5() == oracle.q5(oh["customer"], oh["orders"], oh["lineitem"], oh["supplier"], oh["nation"], oh["region"])[0]


def test_unsupported_and_invalid_descriptors(gpu_ctx):
    from lingodb_b200 import capi, runtime
    s = datagen.scale(0.001)
    g = runtime.Tpch(gpu_ctx, {"lineitem": gpu_ctx.table_from_host(datagen.lineitem(s))})
    with pytest.raises(capi.LdbRuntimeError) as ei:
        g.q6(date_ge="not-a-date")
    assert ei.value.code == capi.LDB_ERR_INVALID
    with pytest.raises(capi.LdbRuntimeError):
        g.q6(disc_ge="abc")


def test_radix_partition_and_tuple_insert(gpu_ctx, oracle):
    """K6: partition by the top bits of the reference hash; every tuple lands in exactly one contiguous block of
    its partition, payload columns travel with their key, and re-inserting the blocks builds a probe-able table."""
    import ctypes as C

    import torch
    from lingodb_b200 import capi
    L = gpu_ctx.L
    n, parts = 200000, 8
    rng = np.random.default_rng(5)
    keys = rng.choice(np.arange(1, 5_000_000, dtype=np.int32), size=n, replace=False)
    pay = (keys.astype(np.int64) * 7 + 1).astype(np.int32)
    wide = (keys.astype(np.int64) * 1000003)
    dev = torch.device("cuda", gpu_ctx.device)
    dk, dp, dw = (torch.from_numpy(a).to(dev) for a in (keys, pay, wide))
    ok, op, ow = torch.empty_like(dk), torch.empty_like(dp), torch.empty_like(dw)
    torch.cuda.synchronize()
    cols = (C.c_void_p * 2)(dp.data_ptr(), dw.data_ptr())
    outs = (C.c_void_p * 2)(op.data_ptr(), ow.data_ptr())
    widths = (C.c_int32 * 2)(4, 8)
    offs = (C.c_int64 * (parts + 1))()
    e = capi.Error()
    capi.check(L.ldb_gpu_partition_tuples(gpu_ctx.h, C.c_void_p(dk.data_ptr()), cols, widths, 2, n, parts, C.c_void_p(ok.data_ptr()), outs, offs, C.byref(e)), e)
    offs = list(offs)
    assert offs[0] == 0 and offs[-1] == n and all(a <= b for a, b in zip(offs, offs[1:]))
    hk, hp, hw = ok.cpu().numpy(), op.cpu().numpy(), ow.cpu().numpy()
    assert sorted(hk.tolist()) == sorted(keys.tolist())
    assert np.array_equal(hp, (hk.astype(np.int64) * 7 + 1).astype(np.int32)) and np.array_equal(hw, hk.astype(np.int64) * 1000003)
    for p in range(parts):  # destination = top bits of h64(key), as the oracle hashes
        for k in hk[offs[p]:offs[p + 1]][:50]:
            h = oracle.lib.oracle_hash_i64(int(k))
            assert ((h >> 32) * parts) >> 32 == p
    # the received block of one partition → join table → probe through a build pipeline is exercised by q5; here: count
    st = C.c_void_p()
    capi.check(L.ldb_gpu_join_table_create(gpu_ctx.h, n, 1, 0, 0, C.byref(st), C.byref(e)), e)
    capi.check(L.ldb_gpu_join_table_insert(gpu_ctx.h, st, C.c_void_p(ok.data_ptr()), C.c_void_p(op.data_ptr()), None, n, C.byref(e)), e)
    cnt = C.c_int64()
    capi.check(L.ldb_gpu_join_table_count(st, C.byref(cnt), C.byref(e)), e)
    assert cnt.value == n
    # duplicate keys in a table declared unique are reported, not silently dropped
    capi.check(L.ldb_gpu_join_table_insert(gpu_ctx.h, st, C.c_void_p(ok.data_ptr()), C.c_void_p(op.data_ptr()), None, 10, C.byref(e)), e)
    rc = L.ldb_gpu_join_table_count(st, C.byref(cnt), C.byref(e))
    assert rc == capi.LDB_ERR_INVALID
    L.ldb_gpu_state_destroy(st)


def test_q5_repartitioned_world1_matches_oracle(gpu_ctx, oracle):
    """The multi-GPU Q5 plan (K8 materialise → K6 partition → exchange → partition-local probes) run with world = 1."""
    from lingodb_b200 import parallel
    t = datagen.tpch(0.05, seed=21, chunk_rows=1 << 20)
    tabs = {k: gpu_ctx.table_from_host(v) for k, v in t.items()}
    oh = {k: oracle.table(v) for k, v in t.items()}
    want, _ = oracle.q5(oh["customer"], oh["orders"], oh["lineitem"], oh["supplier"], oh["nation"], oh["region"])
    got, stats = parallel.q5_repartitioned(gpu_ctx, tabs, 1, 0, t["orders"].num_rows)
    assert got == want
    assert 0 < stats["lineitem_tuples_sent"] < 0.1 * stats["lineitem_rows_scanned"]  # the Bloom semi-join did its job


def test_materialize_pipeline_exact_rows(gpu_ctx):
    import torch
    from lingodb_b200 import parallel
    s = datagen.scale(0.01, seed=2)
    host = datagen.orders(s)
    tab = gpu_ctx.table_from_host(host)
    dev = torch.device("cuda", gpu_ctx.device)
    (k, d), n = parallel._materialize(gpu_ctx, tab, ["o_orderkey", "o_orderdate"], [4, 4], 16, dev, filters=[("o_orderdate", "<", "1993-01-01")])
    c = host.chunks[0]
    m = c["o_orderdate"] < 8401  # 1993-01-01
    assert n == int(m.sum())  # the first attempt overflowed its 16-row buffer and was regrown
    got = sorted(zip(k[:n].cpu().tolist(), d[:n].cpu().tolist()))
    assert got == sorted(zip(c["o_orderkey"][m].tolist(), c["o_orderdate"][m].tolist()))
