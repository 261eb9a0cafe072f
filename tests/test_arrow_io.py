"""Arrow IPC files (the reference's on-disk table format, LingoDBTable.cpp:27-54) → scan layout, zero-copy."""
import numpy as np
import pyarrow as pa
import pytest

from lingodb_b200 import arrow_io, datagen


def _same(a, b):
    if isinstance(a, tuple):
        n = len(a[0]) - 1
        return all(bytes(a[1][a[0][i]:a[0][i + 1]]) == bytes(b[1][b[0][i]:b[0][i + 1]]) for i in range(n)) and len(b[0]) - 1 == n
    return np.array_equal(a, b)


def test_ipc_round_trip_is_bit_exact_and_zero_copy(tmp_path):
    t = datagen.tpch(0.01, seed=5, chunk_rows=3000, with_parts=True)
    for name, td in t.items():
        path = str(tmp_path / f"{name}.arrow")
        arrow_io.write_ipc(path, td)
        back = arrow_io.read_ipc(path)
        assert back.name == name and [(c.name, c.phys, c.precision, c.scale) for c in back.columns] == [(c.name, c.phys, c.precision, c.scale) for c in td.columns]
        assert back.chunk_rows == td.chunk_rows
        for ca, cb in zip(td.chunks, back.chunks):
            for c in td.columns:
                assert _same(ca[c.name], cb[c.name]), (name, c.name)
    # the Arrow types are the reference's physical types (LingoDBTable.cpp:122-195)
    schema, _ = arrow_io.to_arrow_batches(t["lineitem"])
    assert schema.field("l_quantity").type == pa.decimal128(12, 2) and schema.field("l_returnflag").type == pa.binary(4)
    assert schema.field("l_shipdate").type == pa.date32() and schema.field("l_orderkey").type == pa.int32()
    # views borrow the mapped file: no copy of the column data
    back = arrow_io.read_ipc(str(tmp_path / "lineitem.arrow"))
    assert not back.chunks[0]["l_extendedprice"].flags["OWNDATA"]


def test_sliced_batches_offsets_and_column_subset(tmp_path):
    td = datagen.customer(datagen.scale(0.01, seed=2), chunk_rows=1 << 20)
    schema, batches = arrow_io.to_arrow_batches(td)
    sliced = [batches[0].slice(17, 500), batches[0].slice(517, 83)]  # non-zero array offsets, utf8 offsets not starting at 0
    back = arrow_io.tabledata_from_batches("customer", schema, sliced, columns=["c_custkey", "c_mktsegment"], chunk_rows=256)
    assert back.chunk_rows == [256, 244, 83] and [c.name for c in back.columns] == ["c_custkey", "c_mktsegment"]
    keys = np.concatenate([c["c_custkey"] for c in back.chunks])
    assert np.array_equal(keys, td.chunks[0]["c_custkey"][17:600])
    offs, data = td.chunks[0]["c_mktsegment"]
    want = [bytes(data[offs[i]:offs[i + 1]]) for i in range(17, 600)]
    got = []
    for c in back.chunks:
        o, d = c["c_mktsegment"]
        got += [bytes(d[o[i]:o[i + 1]]) for i in range(len(o) - 1)]
    assert got == want


def test_rejects_what_the_gpu_path_does_not_take():
    with pytest.raises(TypeError):
        arrow_io.tabledata_from_batches("t", pa.schema([("x", pa.float64())]), [])
    b = pa.RecordBatch.from_arrays([pa.array([1, None, 3], pa.int32())], names=["x"])
    with pytest.raises(ValueError):
        arrow_io.tabledata_from_batches("t", b.schema, [b])


def test_oracle_scans_arrow_file_in_place(tmp_path):
    from oracle import oracle as O
    t = datagen.tpch(0.02, seed=8, chunk_rows=5000)
    o = O.Oracle("auto", workers=2)
    want = o.q1(o.table(t["lineitem"]))[0]
    arrow_io.write_ipc(str(tmp_path / "lineitem.arrow"), t["lineitem"])
    back = arrow_io.read_ipc(str(tmp_path / "lineitem.arrow"), columns=["l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_returnflag", "l_linestatus", "l_shipdate"])
    assert o.q1(o.table(back))[0] == want


@pytest.mark.gpu
def test_gpu_queries_on_arrow_files(tmp_path, gpu_ctx, oracle):
    from lingodb_b200 import runtime
    t = datagen.tpch(0.03, seed=6, chunk_rows=7000, with_parts=True)
    loaded = {}
    for name, td in t.items():
        arrow_io.write_ipc(str(tmp_path / f"{name}.arrow"), td)
        loaded[name] = arrow_io.read_ipc(str(tmp_path / f"{name}.arrow"))
    g = runtime.Tpch(gpu_ctx, {k: gpu_ctx.table_from_host(v) for k, v in loaded.items()})
    oh = {k: oracle.table(v) for k, v in t.items()}
    assert g.q1() == oracle.q1(oh["lineitem"])[0]
    assert g.q3() == oracle.q3(oh["customer"], oh["orders"], oh["lineitem"])[0]
    assert g.q9() == oracle.q9(oh["part"], oh["supplier"], oh["lineitem"], oh["partsupp"], oh["orders"], oh["nation"])[0]


def test_cpp_reader_opens_the_references_file_format(tmp_path):
    """libldb_arrow_io.so (Arrow C++ RecordBatchFileReader over a memory map, like LingoDBTable::loadTable): schema and batch count
    without a device; nullable and float columns are accepted (the program pipeline reads them)."""
    td = datagen.lineitem(datagen.scale(0.01, seed=3), chunk_rows=20_000)
    path = str(tmp_path / "lineitem.arrow")
    arrow_io.write_ipc(path, td)
    f = arrow_io.ArrowFile(path, ["l_orderkey", "l_quantity", "l_shipdate", "l_returnflag"])
    assert [(c.name, c.phys, c.precision, c.scale) for c in f.schema()] == [("l_orderkey", "int32", 0, 0), ("l_quantity", "decimal128", 12, 2), ("l_returnflag", "fsb4", 0, 0), ("l_shipdate", "date32", 0, 0)]
    assert f.num_batches == len(td.chunks)
    f.close()
    b = pa.RecordBatch.from_arrays([pa.array([1, None, 3], pa.int32()), pa.array([0.5, 1.5, None], pa.float64()), pa.array(["a", None, "c"])], names=["x", "y", "s"])
    p2 = str(tmp_path / "n.arrow")
    with pa.OSFile(p2, "wb") as sink, pa.ipc.new_file(sink, b.schema) as w:
        w.write_batch(b)
    f = arrow_io.ArrowFile(p2)
    assert [(c.name, c.phys) for c in f.schema()] == [("x", "int32"), ("y", "float64"), ("s", "utf8")]
    f.close()
    from lingodb_b200 import capi
    with pytest.raises(capi.LdbRuntimeError):
        arrow_io.ArrowFile(str(tmp_path / "missing.arrow"))


@pytest.mark.gpu
def test_cpp_reader_stages_files_and_the_table_is_the_column_cache(tmp_path, gpu_ctx, oracle):
    """File → ldb_arrow_file_load → compressed staging → Q1/Q6 == oracle; a second query reads the staged table (no new H2D);
    a file with NULLs goes through the program pipeline."""
    from lingodb_b200 import program as P, runtime
    t = datagen.tpch(0.05, seed=8, chunk_rows=100_000)
    path = str(tmp_path / "lineitem.arrow")
    arrow_io.write_ipc(path, t["lineitem"])
    f = arrow_io.ArrowFile(path)
    tab = f.load(gpu_ctx, "lineitem")
    oh = oracle.table(t["lineitem"])
    tp = runtime.Tpch(gpu_ctx, {"lineitem": tab})
    assert tp.q1() == oracle.q1(oh)[0]
    h2d = int(gpu_ctx.L.ldb_gpu_context_h2d_bytes(gpu_ctx.h))
    assert tp.q6() == oracle.q6(oh)[0]
    assert int(gpu_ctx.L.ldb_gpu_context_h2d_bytes(gpu_ctx.h)) == h2d  # column-cache hit: nothing crossed the link again
    tab.clear()
    f.close()
    x = np.arange(100_000, dtype=np.int32)
    arr = pa.array(x, mask=(x % 7 == 0))
    b = pa.RecordBatch.from_arrays([arr, pa.array((x % 5).astype(np.int32))], names=["x", "g"])
    p2 = str(tmp_path / "n.arrow")
    with pa.OSFile(p2, "wb") as sink, pa.ipc.new_file(sink, b.schema) as w:
        w.write_batch(b.slice(0, 60_001))
        w.write_batch(b.slice(60_001))
    f = arrow_io.ArrowFile(p2)
    tn = f.load(gpu_ctx, "n")
    st = P.group_by(gpu_ctx, tn, [("col", "g")], [("sum", ("col", "x")), ("count", ("col", "x")), ("count_star", None)], expected_groups=16)
    got = P.decode_groups(P.read_groups(gpu_ctx, st, 16), 1, 3)
    for g in range(5):
        m = (x % 5 == g) & (x % 7 != 0)
        assert got[(g,)] == [int(x[m].sum()), int(m.sum()), int((x % 5 == g).sum())]
    gpu_ctx.L.ldb_gpu_state_destroy(st)
    tn.clear()
    f.close()
