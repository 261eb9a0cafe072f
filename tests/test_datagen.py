"""Deterministic generator: distribution invariants the plans rely on, random access == sequential."""
import numpy as np

from lingodb_b200 import datagen


def test_scale_and_line_prefix():
    s = datagen.scale(1.0)
    assert (s.n_orders, s.n_customer, s.n_supplier, s.n_part) == (1500000, 150000, 10000, 200000)
    assert s.n_lineitem == 6000004  # 28 lines per 7 orders + the partial last block
    L = datagen.lib()
    import ctypes as C
    assert L.ldbgen_order_first_line(C.byref(s), 0) == 0
    assert L.ldbgen_order_first_line(C.byref(s), 7) == 28
    assert L.ldbgen_order_first_line(C.byref(s), s.n_orders) == s.n_lineitem


def test_keys_and_distributions():
    s = datagen.scale(0.05, seed=3)
    o = datagen.orders(s).chunks[0]
    li = datagen.lineitem(s).chunks[0]
    assert np.all(np.diff(o["o_orderkey"]) > 0)  # sparse, increasing, unique
    assert np.all((o["o_orderkey"] & 0x18) == 0)  # 8 of every 32 key values are used
    assert np.all(o["o_custkey"] % 3 != 0) and o["o_custkey"].min() >= 1 and o["o_custkey"].max() <= s.n_customer
    # every lineitem key exists in orders, 1..7 lines per order, exactly 28 per 7 orders
    keys, counts = np.unique(li["l_orderkey"], return_counts=True)
    assert np.array_equal(keys, o["o_orderkey"])
    assert counts.min() == 1 and counts.max() == 7
    assert counts[: 7 * (len(counts) // 7)].reshape(-1, 7).sum(axis=1).tolist() == [28] * (len(counts) // 7)
    qty = li["l_quantity"][:, :8].copy().view(np.int64).reshape(-1)
    disc = li["l_discount"][:, :8].copy().view(np.int64).reshape(-1)
    assert qty.min() == 100 and qty.max() == 5000 and disc.min() == 0 and disc.max() == 10
    assert np.all(li["l_quantity"][:, 8:] == 0)  # positive decimals: sign-extension bytes are zero
    assert set(np.unique(li["l_returnflag"])) == {ord("A"), ord("N"), ord("R")}
    assert set(np.unique(li["l_linestatus"])) == {ord("F"), ord("O")}
    # l_linestatus / l_returnflag rules of the spec → exactly 4 (flag, status) combinations
    assert len(set(zip(li["l_returnflag"].tolist(), li["l_linestatus"].tolist()))) == 4
    assert 1 <= li["l_suppkey"].min() and li["l_suppkey"].max() <= s.n_supplier


def test_random_access_equals_sequential():
    s = datagen.scale(0.01, seed=5)
    whole = datagen.lineitem(s, chunk_rows=1 << 20).chunks[0]
    part = datagen.lineitem(s, chunk_rows=1 << 20, row_begin=12345, n_rows=1000).chunks[0]
    for k, v in part.items():
        assert np.array_equal(v, whole[k][12345:13345]), k
    ragged = datagen.lineitem(s, chunk_rows=999)
    assert sum(ragged.chunk_rows) == s.n_lineitem
    assert np.array_equal(np.concatenate([c["l_shipdate"] for c in ragged.chunks]), whole["l_shipdate"])
    assert datagen.lineitem(datagen.scale(0.01, seed=6), n_rows=100).chunks[0]["l_partkey"].tolist() != whole["l_partkey"][:100].tolist()


def test_utf8_layout():
    s = datagen.scale(0.01)
    c = datagen.customer(s).chunks[0]
    offs, data = c["c_mktsegment"]
    assert offs[0] == 0 and offs[-1] == len(data) and np.all(np.diff(offs) >= 8)
    segs = {bytes(data[offs[i]:offs[i + 1]]).decode() for i in range(200)}
    assert segs <= {"AUTOMOBILE", "BUILDING", "FURNITURE", "MACHINERY", "HOUSEHOLD"}
    n = datagen.nation().chunks[0]
    assert bytes(n["n_name"][1][n["n_name"][0][8]:n["n_name"][0][9]]) == b"INDIA"


def test_compiled_dbgen_twin_equals_numpy():
    """csrc/dbgen_gen.h (host build; the device generator compiles the same functions) == lingodb_b200/dbgen.py (numpy), which
    tests/test_reference_answers_sf1.py validates against the reference's SF1 answers."""
    from lingodb_b200 import dbgen
    a, b = dbgen.tpch(0.05, chunk_rows=1 << 30), dbgen.tpch_compiled(0.05, chunk_rows=1 << 30)
    assert a["lineitem"].num_rows == b["lineitem"].num_rows > 290000
    for name in a:
        for c in a[name].columns:
            x, y = a[name].chunks[0][c.name], b[name].chunks[0][c.name]
            if isinstance(x, tuple):
                assert np.array_equal(x[0], y[0]) and np.array_equal(x[1][: x[0][-1]], y[1][: y[0][-1]]), (name, c.name)
            else:
                assert np.array_equal(x, y), (name, c.name)
    # random access: an order range generated on its own equals the slice of the whole
    import ctypes as C
    from lingodb_b200 import datagen
    L = datagen.lib()
    s = dbgen.scale_compiled(0.05)
    counts = np.zeros(1000, np.int32)
    L.ldbgen_dbgen_line_counts_host(C.byref(s), 40000, 1000, datagen._ptr(counts))
    _, per_order = np.unique(a["lineitem"].chunks[0]["l_orderkey"], return_counts=True)  # order keys ascend with the order index
    assert counts.tolist() == per_order[40000:41000].tolist()
