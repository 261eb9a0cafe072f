"""Host-side deterministic TPC-H-shaped tables in the reference's Arrow physical layout.

Thin ctypes front of libldb_datagen_host.so (csrc/datagen_host.cpp, csrc/tpch_gen.h).  The arrays
are plain numpy buffers laid out exactly like the Arrow buffers LingoDB scans
(src/runtime/storage/LingoDBTable.cpp:122-195): int32 / date32 / fixed_size_binary(4) → int32,
decimal128 → 16 bytes per value (uint8[n,16]), utf8 → (int32 offsets[n+1], uint8 data).
The same tables feed the CPU oracle and the host-buffer path of the GPU C-ABI.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

from . import build as _build

PHYS = {"int32": 0, "int64": 1, "date32": 2, "decimal128": 3, "fsb4": 4, "utf8": 5}


class GenScale(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_orders", C.c_int64), ("n_customer", C.c_int64),
                ("n_supplier", C.c_int64), ("n_part", C.c_int64), ("n_lineitem", C.c_int64)]


class LineitemCols(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "l_orderkey", "l_partkey", "l_suppkey", "l_quantity", "l_extendedprice", "l_discount", "l_tax",
        "l_returnflag", "l_linestatus", "l_shipdate", "l_commitdate", "l_receiptdate")]


class OrdersCols(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("o_orderkey", "o_custkey", "o_orderdate", "o_shippriority")]


class CustomerCols(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("c_custkey", "c_nationkey", "c_mktsegment_offsets", "c_mktsegment_data")]


class SupplierCols(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("s_suppkey", "s_nationkey")]


class PartCols(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("p_partkey", "p_name_offsets", "p_name_data")]


class PartsuppCols(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("ps_partkey", "ps_suppkey", "ps_supplycost")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build_datagen_host())
        _lib.ldbgen_scale.argtypes = [C.c_double, C.c_uint64, C.POINTER(GenScale)]
        _lib.ldbgen_order_first_line.restype = C.c_int64
        _lib.ldbgen_order_first_line.argtypes = [C.POINTER(GenScale), C.c_int64]
        _lib.ldbgen_lineitem_host.argtypes = [C.POINTER(GenScale), C.c_int64, C.c_int64, C.POINTER(LineitemCols)]
        _lib.ldbgen_orders_host.argtypes = [C.POINTER(GenScale), C.c_int64, C.c_int64, C.POINTER(OrdersCols)]
        _lib.ldbgen_customer_host.restype = C.c_int64
        _lib.ldbgen_customer_host.argtypes = [C.POINTER(GenScale), C.c_int64, C.c_int64, C.POINTER(CustomerCols)]
        _lib.ldbgen_supplier_host.argtypes = [C.POINTER(GenScale), C.c_int64, C.c_int64, C.POINTER(SupplierCols)]
        _lib.ldbgen_part_host.restype = C.c_int64
        _lib.ldbgen_part_host.argtypes = [C.POINTER(GenScale), C.c_int64, C.c_int64, C.POINTER(PartCols)]
        _lib.ldbgen_partsupp_host.argtypes = [C.POINTER(GenScale), C.c_int64, C.c_int64, C.POINTER(PartsuppCols)]
    return _lib


def scale(sf: float, seed: int = 42) -> GenScale:
    s = GenScale()
    lib().ldbgen_scale(float(sf), int(seed), C.byref(s))
    return s


@dataclass
class ColumnSpec:
    name: str
    phys: str  # key of PHYS
    precision: int = 0
    scale: int = 0


@dataclass
class TableData:
    """A table as a list of record batches ("chunks"); each chunk maps column name → buffers."""
    name: str
    columns: List[ColumnSpec]
    chunks: List[Dict[str, object]] = field(default_factory=list)  # np.ndarray | (offsets, data)
    chunk_rows: List[int] = field(default_factory=list)

    @property
    def num_rows(self) -> int:
        return sum(self.chunk_rows)

    def spec(self, name: str) -> ColumnSpec:
        return next(c for c in self.columns if c.name == name)


LINEITEM_SCHEMA = [
    ColumnSpec("l_orderkey", "int32"), ColumnSpec("l_partkey", "int32"), ColumnSpec("l_suppkey", "int32"),
    ColumnSpec("l_quantity", "decimal128", 12, 2), ColumnSpec("l_extendedprice", "decimal128", 12, 2),
    ColumnSpec("l_discount", "decimal128", 12, 2), ColumnSpec("l_tax", "decimal128", 12, 2),
    ColumnSpec("l_returnflag", "fsb4"), ColumnSpec("l_linestatus", "fsb4"),
    ColumnSpec("l_shipdate", "date32"), ColumnSpec("l_commitdate", "date32"), ColumnSpec("l_receiptdate", "date32"),
]
ORDERS_SCHEMA = [ColumnSpec("o_orderkey", "int32"), ColumnSpec("o_custkey", "int32"),
                 ColumnSpec("o_orderdate", "date32"), ColumnSpec("o_shippriority", "int32")]
CUSTOMER_SCHEMA = [ColumnSpec("c_custkey", "int32"), ColumnSpec("c_nationkey", "int32"), ColumnSpec("c_mktsegment", "utf8")]
SUPPLIER_SCHEMA = [ColumnSpec("s_suppkey", "int32"), ColumnSpec("s_nationkey", "int32")]
PART_SCHEMA = [ColumnSpec("p_partkey", "int32"), ColumnSpec("p_name", "utf8")]
PARTSUPP_SCHEMA = [ColumnSpec("ps_partkey", "int32"), ColumnSpec("ps_suppkey", "int32"), ColumnSpec("ps_supplycost", "decimal128", 12, 2)]
NATION_SCHEMA = [ColumnSpec("n_nationkey", "int32"), ColumnSpec("n_name", "utf8"), ColumnSpec("n_regionkey", "int32")]
REGION_SCHEMA = [ColumnSpec("r_regionkey", "int32"), ColumnSpec("r_name", "utf8")]

# TPC-H spec 4.2.3 fixed tables
NATIONS: List[Tuple[str, int]] = [
    ("ALGERIA", 0), ("ARGENTINA", 1), ("BRAZIL", 1), ("CANADA", 1), ("EGYPT", 4), ("ETHIOPIA", 0), ("FRANCE", 3),
    ("GERMANY", 3), ("INDIA", 2), ("INDONESIA", 2), ("IRAN", 4), ("IRAQ", 4), ("JAPAN", 2), ("JORDAN", 4), ("KENYA", 0),
    ("MOROCCO", 0), ("MOZAMBIQUE", 0), ("PERU", 1), ("CHINA", 2), ("ROMANIA", 3), ("SAUDI ARABIA", 4), ("VIETNAM", 2),
    ("RUSSIA", 3), ("UNITED KINGDOM", 3), ("UNITED STATES", 1)]
REGIONS = ["AFRICA", "AMERICA", "ASIA", "EUROPE", "MIDDLE EAST"]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _alloc(spec: ColumnSpec, n: int):
    if spec.phys == "decimal128":
        return np.zeros((n, 16), dtype=np.uint8)
    return np.zeros(n, dtype=np.int32)


def _chunks(n: int, chunk_rows: int):
    b = 0
    while b < n:
        yield b, min(chunk_rows, n - b)
        b += chunk_rows


def utf8_column(strings: List[str]):
    data = "".join(strings).encode()
    offs = np.zeros(len(strings) + 1, dtype=np.int32)
    np.cumsum([len(s.encode()) for s in strings], out=offs[1:])
    return offs, np.frombuffer(data, dtype=np.uint8).copy()


def lineitem(s: GenScale, columns=None, chunk_rows: int = 1 << 20, row_begin: int = 0, n_rows: int = None) -> TableData:
    cols = [c for c in LINEITEM_SCHEMA if columns is None or c.name in columns]
    n_rows = s.n_lineitem - row_begin if n_rows is None else n_rows
    t = TableData("lineitem", cols)
    for b, n in _chunks(n_rows, chunk_rows):
        arrs = {c.name: _alloc(c, n) for c in cols}
        lc = LineitemCols(**{k: _ptr(v) for k, v in arrs.items()})
        lib().ldbgen_lineitem_host(C.byref(s), row_begin + b, n, C.byref(lc))
        t.chunks.append(arrs)
        t.chunk_rows.append(n)
    return t


def orders(s: GenScale, chunk_rows: int = 1 << 20, row_begin: int = 0, n_rows: int = None) -> TableData:
    n_rows = s.n_orders - row_begin if n_rows is None else n_rows
    t = TableData("orders", ORDERS_SCHEMA)
    for b, n in _chunks(n_rows, chunk_rows):
        arrs = {c.name: _alloc(c, n) for c in ORDERS_SCHEMA}
        oc = OrdersCols(**{k: _ptr(v) for k, v in arrs.items()})
        lib().ldbgen_orders_host(C.byref(s), row_begin + b, n, C.byref(oc))
        t.chunks.append(arrs)
        t.chunk_rows.append(n)
    return t


def customer(s: GenScale, chunk_rows: int = 1 << 20) -> TableData:
    t = TableData("customer", CUSTOMER_SCHEMA)
    for b, n in _chunks(s.n_customer, chunk_rows):
        ck, cn = np.zeros(n, np.int32), np.zeros(n, np.int32)
        offs = np.zeros(n + 1, np.int32)
        cc = CustomerCols(_ptr(ck), _ptr(cn), _ptr(offs), None)
        nbytes = lib().ldbgen_customer_host(C.byref(s), b, n, C.byref(cc))
        data = np.zeros(max(1, nbytes), np.uint8)
        cc = CustomerCols(None, None, None, _ptr(data))
        lib().ldbgen_customer_host(C.byref(s), b, n, C.byref(cc))
        t.chunks.append({"c_custkey": ck, "c_nationkey": cn, "c_mktsegment": (offs, data)})
        t.chunk_rows.append(n)
    return t


def supplier(s: GenScale, chunk_rows: int = 1 << 20) -> TableData:
    t = TableData("supplier", SUPPLIER_SCHEMA)
    for b, n in _chunks(s.n_supplier, chunk_rows):
        sk, sn = np.zeros(n, np.int32), np.zeros(n, np.int32)
        sc = SupplierCols(_ptr(sk), _ptr(sn))
        lib().ldbgen_supplier_host(C.byref(s), b, n, C.byref(sc))
        t.chunks.append({"s_suppkey": sk, "s_nationkey": sn})
        t.chunk_rows.append(n)
    return t


def part(s: GenScale, chunk_rows: int = 1 << 20) -> TableData:
    t = TableData("part", PART_SCHEMA)
    for b, n in _chunks(s.n_part, chunk_rows):
        pk = np.zeros(n, np.int32)
        offs = np.zeros(n + 1, np.int32)
        pc = PartCols(_ptr(pk), _ptr(offs), None)
        nbytes = lib().ldbgen_part_host(C.byref(s), b, n, C.byref(pc))
        data = np.zeros(max(1, nbytes), np.uint8)
        pc = PartCols(None, None, _ptr(data))
        lib().ldbgen_part_host(C.byref(s), b, n, C.byref(pc))
        t.chunks.append({"p_partkey": pk, "p_name": (offs, data)})
        t.chunk_rows.append(n)
    return t


def partsupp(s: GenScale, chunk_rows: int = 1 << 20) -> TableData:
    t = TableData("partsupp", PARTSUPP_SCHEMA)
    for b, n in _chunks(4 * s.n_part, chunk_rows):
        arrs = {c.name: _alloc(c, n) for c in PARTSUPP_SCHEMA}
        pc = PartsuppCols(**{k: _ptr(v) for k, v in arrs.items()})
        lib().ldbgen_partsupp_host(C.byref(s), b, n, C.byref(pc))
        t.chunks.append(arrs)
        t.chunk_rows.append(n)
    return t


def nation() -> TableData:
    t = TableData("nation", NATION_SCHEMA)
    t.chunks.append({"n_nationkey": np.arange(25, dtype=np.int32), "n_name": utf8_column([n for n, _ in NATIONS]),
                     "n_regionkey": np.array([r for _, r in NATIONS], dtype=np.int32)})
    t.chunk_rows.append(25)
    return t


def region() -> TableData:
    t = TableData("region", REGION_SCHEMA)
    t.chunks.append({"r_regionkey": np.arange(5, dtype=np.int32), "r_name": utf8_column(REGIONS)})
    t.chunk_rows.append(5)
    return t


def tpch(sf: float, seed: int = 42, chunk_rows: int = 1 << 20, lineitem_columns=None, with_parts: bool = False) -> Dict[str, TableData]:
    s = scale(sf, seed)
    t = {"lineitem": lineitem(s, lineitem_columns, chunk_rows), "orders": orders(s, chunk_rows),
         "customer": customer(s, chunk_rows), "supplier": supplier(s, chunk_rows), "nation": nation(), "region": region()}
    if with_parts:  # Q9
        t["part"] = part(s, chunk_rows)
        t["partsupp"] = partsupp(s, chunk_rows)
    return t


def dec128_to_int(a: np.ndarray) -> np.ndarray:
    """uint8[n,16] little-endian two's complement → python ints (object array); test helper."""
    lo = a[:, :8].copy().view(np.uint64).reshape(-1)
    hi = a[:, 8:].copy().view(np.int64).reshape(-1)
    return np.array([(int(h) << 64) | int(l) for l, h in zip(lo, hi)], dtype=object)
