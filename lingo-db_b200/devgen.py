"""Device-side table generation: fills torch CUDA tensors (plain HBM buffers) in the Arrow physical
layout with the device twin of the deterministic generator (csrc/datagen.cu) and registers them
as borrowed DEVICE batches.  torch is only the allocator here."""
import ctypes as C
from typing import Dict, List

import torch

from . import capi, datagen
from .capi import Error, check
from .runtime import Context, Table


def _alloc(spec: datagen.ColumnSpec, n: int, dev):
    if spec.phys == "decimal128":
        return torch.empty((n, 16), dtype=torch.uint8, device=dev)
    return torch.empty(n, dtype=torch.int32, device=dev)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def lineitem(ctx: Context, s: datagen.GenScale, columns: List[str], row_begin: int = 0, n_rows: int = None, batch_rows: int = None) -> Table:
    dev = torch.device("cuda", ctx.device)
    n_rows = s.n_lineitem - row_begin if n_rows is None else n_rows
    batch_rows = n_rows if not batch_rows else batch_rows
    specs = [c for c in datagen.LINEITEM_SCHEMA if c.name in columns]
    tab = Table(ctx, "lineitem", specs)
    b = 0
    while b < n_rows:
        n = min(batch_rows, n_rows - b)
        tens = {c.name: _alloc(c, n, dev) for c in specs}
        torch.cuda.synchronize(dev)
        cols = datagen.LineitemCols(**{k: _ptr(v) for k, v in tens.items()})
        e = Error()
        check(ctx.L.ldb_gpu_datagen_lineitem(ctx.h, C.byref(s), row_begin + b, n, C.byref(cols), C.byref(e)), e)
        tab.append_device(tens, n)
        b += n
    ctx.synchronize()
    return tab


def orders(ctx: Context, s: datagen.GenScale, row_begin: int = 0, n_rows: int = None) -> Table:
    dev = torch.device("cuda", ctx.device)
    n = s.n_orders - row_begin if n_rows is None else n_rows
    tens = {c.name: _alloc(c, n, dev) for c in datagen.ORDERS_SCHEMA}
    torch.cuda.synchronize(dev)
    cols = datagen.OrdersCols(**{k: _ptr(v) for k, v in tens.items()})
    e = Error()
    check(ctx.L.ldb_gpu_datagen_orders(ctx.h, C.byref(s), row_begin, n, C.byref(cols), C.byref(e)), e)
    tab = Table(ctx, "orders", datagen.ORDERS_SCHEMA)
    tab.append_device(tens, n)
    ctx.synchronize()
    return tab


def customer(ctx: Context, s: datagen.GenScale) -> Table:
    dev = torch.device("cuda", ctx.device)
    n = s.n_customer
    ck = torch.empty(n, dtype=torch.int32, device=dev)
    cn = torch.empty(n, dtype=torch.int32, device=dev)
    lens = torch.empty(n, dtype=torch.int32, device=dev)
    torch.cuda.synchronize(dev)
    e = Error()
    cols = datagen.CustomerCols(_ptr(ck), _ptr(cn), None, None)
    check(ctx.L.ldb_gpu_datagen_customer_fixed(ctx.h, C.byref(s), 0, n, C.byref(cols), _ptr(lens), C.byref(e)), e)
    ctx.synchronize()
    offs = torch.zeros(n + 1, dtype=torch.int32, device=dev)
    offs[1:] = torch.cumsum(lens, 0, dtype=torch.int64).to(torch.int32)
    data = torch.empty(int(offs[-1].item()), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize(dev)
    check(ctx.L.ldb_gpu_datagen_customer_bytes(ctx.h, C.byref(s), 0, n, _ptr(offs), _ptr(data), C.byref(e)), e)
    tab = Table(ctx, "customer", datagen.CUSTOMER_SCHEMA)
    tab.append_device({"c_custkey": ck, "c_nationkey": cn, "c_mktsegment": (offs, data)}, n)
    ctx.synchronize()
    return tab


def supplier(ctx: Context, s: datagen.GenScale) -> Table:
    dev = torch.device("cuda", ctx.device)
    n = s.n_supplier
    tens = {c.name: _alloc(c, n, dev) for c in datagen.SUPPLIER_SCHEMA}
    torch.cuda.synchronize(dev)
    cols = datagen.SupplierCols(**{k: _ptr(v) for k, v in tens.items()})
    e = Error()
    check(ctx.L.ldb_gpu_datagen_supplier(ctx.h, C.byref(s), 0, n, C.byref(cols), C.byref(e)), e)
    tab = Table(ctx, "supplier", datagen.SUPPLIER_SCHEMA)
    tab.append_device(tens, n)
    ctx.synchronize()
    return tab


def part(ctx: Context, s: datagen.GenScale, batch_rows: int = 16 << 20) -> Table:
    """utf8 offsets are int32 per batch: 60 M names at SF300 are > 2 GiB, so the table is made of several batches."""
    dev = torch.device("cuda", ctx.device)
    tab = Table(ctx, "part", datagen.PART_SCHEMA)
    e = Error()
    b = 0
    while b < s.n_part:
        n = min(batch_rows, s.n_part - b)
        pk = torch.empty(n, dtype=torch.int32, device=dev)
        lens = torch.empty(n, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)
        cols = datagen.PartCols(_ptr(pk), None, None)
        check(ctx.L.ldb_gpu_datagen_part_fixed(ctx.h, C.byref(s), b, n, C.byref(cols), _ptr(lens), C.byref(e)), e)
        ctx.synchronize()
        offs = torch.zeros(n + 1, dtype=torch.int32, device=dev)
        offs[1:] = torch.cumsum(lens, 0, dtype=torch.int64).to(torch.int32)
        data = torch.empty(int(offs[-1].item()), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize(dev)
        check(ctx.L.ldb_gpu_datagen_part_bytes(ctx.h, C.byref(s), b, n, _ptr(offs), _ptr(data), C.byref(e)), e)
        tab.append_device({"p_partkey": pk, "p_name": (offs, data)}, n)
        b += n
    ctx.synchronize()
    return tab


def partsupp(ctx: Context, s: datagen.GenScale) -> Table:
    dev = torch.device("cuda", ctx.device)
    n = 4 * s.n_part
    tens = {c.name: _alloc(c, n, dev) for c in datagen.PARTSUPP_SCHEMA}
    torch.cuda.synchronize(dev)
    cols = datagen.PartsuppCols(**{k: _ptr(v) for k, v in tens.items()})
    e = Error()
    check(ctx.L.ldb_gpu_datagen_partsupp(ctx.h, C.byref(s), 0, n, C.byref(cols), C.byref(e)), e)
    tab = Table(ctx, "partsupp", datagen.PARTSUPP_SCHEMA)
    tab.append_device(tens, n)
    ctx.synchronize()
    return tab


def dbgen_tables(ctx: Context, sf: float, lineitem_columns: List[str] = None, order_begin: int = 0, n_orders: int = None, part_batch_rows: int = 16 << 20) -> Dict[str, Table]:
    """dbgen-faithful tables generated in HBM (csrc/dbgen_gen.h; host twin: dbgen.tpch_compiled).  lineitem/orders may be an
    order range [order_begin, order_begin + n_orders) — the shard of one GPU; the small tables are always whole.
    The device kernels share every value function with the host twin (tests/test_datagen.py) and are compared with it bit
    for bit by tests/test_gpu_parity.py::test_device_dbgen_twin_matches_host_twin."""
    from . import dbgen
    dev = torch.device("cuda", ctx.device)
    s = dbgen.scale_compiled(sf, count_lines=False)
    n_orders = s.n_orders - order_begin if n_orders is None else n_orders
    e = Error()
    counts = torch.empty(n_orders, dtype=torch.int32, device=dev)
    torch.cuda.synchronize(dev)
    check(ctx.L.ldb_gpu_dbgen_line_counts(ctx.h, C.byref(s), order_begin, n_orders, _ptr(counts), C.byref(e)), e)
    ctx.synchronize()
    first = torch.zeros(n_orders + 1, dtype=torch.int64, device=dev)
    first[1:] = torch.cumsum(counts, 0, dtype=torch.int64)
    n_lines = int(first[-1].item())
    specs = [c for c in datagen.LINEITEM_SCHEMA if lineitem_columns is None or c.name in lineitem_columns]
    tens = {c.name: _alloc(c, n_lines, dev) for c in specs}
    torch.cuda.synchronize(dev)
    cols = datagen.LineitemCols(**{k: _ptr(v) for k, v in tens.items()})
    check(ctx.L.ldb_gpu_dbgen_lineitem(ctx.h, C.byref(s), order_begin, n_orders, _ptr(first), C.byref(cols), C.byref(e)), e)
    lineitem_t = Table(ctx, "lineitem", specs)
    lineitem_t.append_device(tens, n_lines)
    ot = {c.name: _alloc(c, n_orders, dev) for c in datagen.ORDERS_SCHEMA}
    torch.cuda.synchronize(dev)
    check(ctx.L.ldb_gpu_dbgen_orders(ctx.h, C.byref(s), order_begin, n_orders, C.byref(datagen.OrdersCols(**{k: _ptr(v) for k, v in ot.items()})), C.byref(e)), e)
    orders_t = Table(ctx, "orders", datagen.ORDERS_SCHEMA)
    orders_t.append_device(ot, n_orders)

    def utf8_table(name, schema, table_id, n_rows, key_name, second_name, text_name, batch_rows):
        tab = Table(ctx, name, schema)
        b = 0
        while b < n_rows:
            n = min(batch_rows, n_rows - b)
            key = torch.empty(n, dtype=torch.int32, device=dev)
            second = torch.empty(n, dtype=torch.int32, device=dev) if second_name else None
            lens = torch.empty(n, dtype=torch.int32, device=dev)
            torch.cuda.synchronize(dev)
            check(ctx.L.ldb_gpu_dbgen_small_fixed(ctx.h, C.byref(s), table_id, b, n, _ptr(key), _ptr(second), None, _ptr(lens), C.byref(e)), e)
            ctx.synchronize()
            offs = torch.zeros(n + 1, dtype=torch.int32, device=dev)
            offs[1:] = torch.cumsum(lens, 0, dtype=torch.int64).to(torch.int32)
            data = torch.empty(int(offs[-1].item()), dtype=torch.uint8, device=dev)
            torch.cuda.synchronize(dev)
            check(ctx.L.ldb_gpu_dbgen_bytes(ctx.h, C.byref(s), table_id, b, n, _ptr(offs), _ptr(data), C.byref(e)), e)
            chunk = {key_name: key, text_name: (offs, data)}
            if second_name:
                chunk[second_name] = second
            tab.append_device(chunk, n)
            b += n
        return tab

    customer_t = utf8_table("customer", datagen.CUSTOMER_SCHEMA, 0, s.n_customer, "c_custkey", "c_nationkey", "c_mktsegment", 1 << 30)
    part_t = utf8_table("part", datagen.PART_SCHEMA, 2, s.n_part, "p_partkey", None, "p_name", part_batch_rows)
    sk, sn = torch.empty(s.n_supplier, dtype=torch.int32, device=dev), torch.empty(s.n_supplier, dtype=torch.int32, device=dev)
    n_ps = 4 * s.n_part
    pk, psk, cost = torch.empty(n_ps, dtype=torch.int32, device=dev), torch.empty(n_ps, dtype=torch.int32, device=dev), torch.empty((n_ps, 16), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize(dev)
    check(ctx.L.ldb_gpu_dbgen_small_fixed(ctx.h, C.byref(s), 1, 0, s.n_supplier, _ptr(sk), _ptr(sn), None, None, C.byref(e)), e)
    check(ctx.L.ldb_gpu_dbgen_small_fixed(ctx.h, C.byref(s), 3, 0, n_ps, _ptr(pk), _ptr(psk), _ptr(cost), None, C.byref(e)), e)
    supplier_t = Table(ctx, "supplier", datagen.SUPPLIER_SCHEMA)
    supplier_t.append_device({"s_suppkey": sk, "s_nationkey": sn}, s.n_supplier)
    partsupp_t = Table(ctx, "partsupp", datagen.PARTSUPP_SCHEMA)
    partsupp_t.append_device({"ps_partkey": pk, "ps_suppkey": psk, "ps_supplycost": cost}, n_ps)
    ctx.synchronize()
    return {"lineitem": lineitem_t, "orders": orders_t, "customer": customer_t, "supplier": supplier_t, "part": part_t, "partsupp": partsupp_t, **small_tables(ctx)}


def small_tables(ctx: Context) -> Dict[str, Table]:
    return {"nation": ctx.table_from_host(datagen.nation()), "region": ctx.table_from_host(datagen.region())}


def to_host(tab: Table) -> datagen.TableData:
    """Copy a device-generated table back into host TableData (tests: device twin == host generator)."""
    t = datagen.TableData(tab.name, tab.columns)
    for item in tab._keep:
        if not isinstance(item, dict):
            continue
        chunk, n = {}, None
        for c in tab.columns:
            v = item[c.name]
            if c.phys == "utf8":
                chunk[c.name] = (v[0].cpu().numpy(), v[1].cpu().numpy())
                n = v[0].numel() - 1
            else:
                chunk[c.name] = v.cpu().numpy()
                n = v.shape[0]
        t.chunks.append(chunk)
        t.chunk_rows.append(n)
    return t
