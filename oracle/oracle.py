"""oracle.py — TEST INFRASTRUCTURE ONLY: ctypes front of the CPU oracle.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this module.  Parity status: pinned (reference KATs + the reference's own tpchSf1.test answers on dbgen-faithful
tables, tests/test_reference_answers_sf1.py).  It loads oracle/_ref/liboracle_ref.so (the reference's own runtime objects
compiled from /root/reference, kind "reference") when that library exists and loads, else
oracle/liboracle_port.so (self-contained restatement, kind "port").
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_LIB = os.path.join(HERE, "liboracle_port.so")
REF_LIB = os.path.join(HERE, "_ref", "liboracle_ref.so")
PHYS = {"int32": 0, "int64": 1, "date32": 2, "decimal128": 3, "fsb4": 4, "utf8": 5}


class Q1Row(C.Structure):
    _fields_ = [("returnflag", C.c_int32), ("linestatus", C.c_int32), ("sum_qty", C.c_int64), ("sum_base_price", C.c_int64),
                ("sum_disc_price", C.c_int64 * 2), ("sum_charge", C.c_int64 * 2), ("avg_qty", C.c_int64 * 2),
                ("avg_price", C.c_int64 * 2), ("avg_disc", C.c_int64 * 2), ("count", C.c_int64)]


class Q3Row(C.Structure):
    _fields_ = [("orderkey", C.c_int32), ("orderdate", C.c_int32), ("shippriority", C.c_int32), ("pad", C.c_int32),
                ("revenue", C.c_int64 * 2)]


class Q5Row(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("revenue", C.c_int64 * 2)]


class CountRow(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("a", C.c_int64), ("b", C.c_int64)]


class Q18Row(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("custkey", C.c_int32), ("orderkey", C.c_int32), ("orderdate", C.c_int32), ("pad", C.c_int32),
                ("totalprice", C.c_int64), ("sum_quantity", C.c_int64)]


class Q9Row(C.Structure):
    _fields_ = [("nation", C.c_char * 32), ("year", C.c_int64), ("sum_profit", C.c_int64 * 2)]


def i128(pair) -> int:
    """{lo, hi} int64 pair → python int (two's complement 128-bit)."""
    return (int(pair[1]) << 64) | (int(pair[0]) & 0xFFFFFFFFFFFFFFFF)


def build(ref: bool = True, verbose: bool = False):
    """Compile the restatement (always) and, when /root/reference is present, oracle/_ref."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.run(["make", "-C", HERE, "port"], check=True, stdout=out)
    if ref and os.path.isdir("/root/reference/src/runtime"):
        subprocess.run(["make", "-C", HERE, "-j8", "ref"], check=True, stdout=out)


class Oracle:
    def __init__(self, kind: str = "auto", workers: int = 0):
        path = None
        if kind in ("auto", "reference") and os.path.exists(REF_LIB):
            try:
                self.lib = C.CDLL(REF_LIB)
                path = REF_LIB
            except OSError:
                if kind == "reference":
                    raise
        if path is None:
            if kind == "reference":
                raise FileNotFoundError(REF_LIB)
            if not os.path.exists(PORT_LIB):
                build(ref=False)
            self.lib = C.CDLL(PORT_LIB)
            path = PORT_LIB
        L = self.lib
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_runtime_kind.restype = C.c_char_p
        L.oracle_table_create.restype = C.c_void_p
        L.oracle_table_create.argtypes = [C.c_char_p]
        L.oracle_table_free.argtypes = [C.c_void_p]
        L.oracle_table_add_column.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
        L.oracle_table_add_chunk.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]
        for f in ("oracle_hash_i64", "oracle_hash_bool", "oracle_hash_i128", "oracle_hash_date_days", "oracle_hash_string",
                  "oracle_hash_combine", "oracle_xxh64"):
            getattr(L, f).restype = C.c_uint64
        L.oracle_hash_i64.argtypes = [C.c_int64]
        L.oracle_hash_bool.argtypes = [C.c_int]
        L.oracle_hash_i128.argtypes = [C.c_int64, C.c_int64]
        L.oracle_hash_date_days.argtypes = [C.c_int32]
        L.oracle_hash_string.argtypes = [C.c_char_p, C.c_uint32]
        L.oracle_hash_combine.argtypes = [C.c_uint64, C.c_uint64]
        L.oracle_xxh64.argtypes = [C.c_char_p, C.c_uint64]
        L.oracle_parse_date.restype = C.c_int32
        L.oracle_parse_date.argtypes = [C.c_char_p]
        L.oracle_parse_decimal.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.oracle_avg_dec12_2.argtypes = [C.c_int64, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.oracle_mul_i128.argtypes = [C.c_int64] * 4 + [C.POINTER(C.c_int64)] * 2
        L.oracle_q6.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int64,
                                C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_double)]
        L.oracle_q1.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(Q1Row), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.oracle_q3.argtypes = [C.c_void_p] * 3 + [C.c_char_p] * 2 + [C.POINTER(Q3Row), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.oracle_q5.argtypes = [C.c_void_p] * 6 + [C.c_char_p] * 3 + [C.POINTER(Q5Row), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.oracle_q9.argtypes = [C.c_void_p] * 6 + [C.c_char_p, C.POINTER(Q9Row), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.oracle_q4.argtypes = [C.c_void_p] * 2 + [C.c_char_p] * 2 + [C.POINTER(CountRow), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.oracle_q12.argtypes = [C.c_void_p] * 2 + [C.c_char_p] * 4 + [C.POINTER(CountRow), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.oracle_q18.argtypes = [C.c_void_p] * 3 + [C.c_int64, C.POINTER(Q18Row), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.oracle_scan_count_sum.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_char_p), C.POINTER(C.c_int64),
                                            C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.oracle_extract_year.restype = C.c_int64
        L.oracle_extract_year.argtypes = [C.c_int64]
        L.oracle_const_like_contains.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p]
        self.path = path
        self.kind = L.oracle_runtime_kind().decode()
        L.oracle_set_workers(int(workers))
        self.workers = L.oracle_num_workers()
        self._keep = []

    def set_workers(self, n: int):
        self.lib.oracle_set_workers(int(n))
        self.workers = self.lib.oracle_num_workers()

    # ---- tables: borrow the numpy buffers of a lingodb_b200.datagen.TableData
    def table(self, t):
        L = self.lib
        h = L.oracle_table_create(t.name.encode())
        for c in t.columns:
            L.oracle_table_add_column(h, c.name.encode(), PHYS[c.phys], c.precision, c.scale)
        for chunk, n in zip(t.chunks, t.chunk_rows):
            ptrs = (C.c_void_p * (3 * len(t.columns)))()
            for i, c in enumerate(t.columns):
                v = chunk[c.name]
                if c.phys == "utf8":
                    offs, data = v
                    ptrs[3 * i + 1] = offs.ctypes.data
                    ptrs[3 * i + 2] = data.ctypes.data
                else:
                    assert v.flags["C_CONTIGUOUS"]
                    ptrs[3 * i + 1] = v.ctypes.data
                if c.name + "$valid" in chunk:  # Arrow validity bitmap of a nullable column
                    ptrs[3 * i] = chunk[c.name + "$valid"].ctypes.data
            L.oracle_table_add_chunk(h, n, ptrs)
            self._keep.append((chunk, ptrs))
        return h

    def free(self, h):
        self.lib.oracle_table_free(h)

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.oracle_last_error().decode())

    # ---- queries: python-int results + pipeline seconds
    OPS = {"=": 0, "!=": 1, "<": 2, "<=": 3, ">": 4, ">=": 5, "notnull": 6}

    def scan_count_sum(self, table, filters, sum_column=None):
        """count(*) and sum(sum_column) of the rows the reference's pushed-down filters keep; filters: (column, op, value | None);
        nullable columns carry "<col>$valid" bitmaps in their chunks (NOTNULL consults them, Restrictions.cpp:67-162)."""
        n = len(filters)
        cols = (C.c_char_p * max(1, n))(*[f[0].encode() for f in filters])
        ops = (C.c_int * max(1, n))(*[self.OPS[f[1]] for f in filters])
        is_int = (C.c_int * max(1, n))(*[int(isinstance(f[2], int)) for f in filters])
        strs = (C.c_char_p * max(1, n))(*[(f[2].encode() if isinstance(f[2], str) else None) for f in filters])
        ints = (C.c_int64 * max(1, n))(*[(f[2] if isinstance(f[2], int) else 0) for f in filters])
        cnt, lo, hi = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self.lib.oracle_scan_count_sum(table, n, cols, ops, is_int, strs, ints, sum_column.encode() if sum_column else None, C.byref(cnt), C.byref(lo), C.byref(hi)))
        return cnt.value, i128((lo.value, hi.value))

    def q6(self, lineitem, date_ge="1994-01-01", date_lt="1995-01-01", disc_ge="0.05", disc_le="0.07", qty_lt=24):
        lo, hi, sec = C.c_int64(), C.c_int64(), C.c_double()
        self._check(self.lib.oracle_q6(lineitem, date_ge.encode(), date_lt.encode(), disc_ge.encode(), disc_le.encode(), qty_lt,
                                       C.byref(lo), C.byref(hi), C.byref(sec)))
        return {"revenue": i128((lo.value, hi.value))}, sec.value

    def q1(self, lineitem, date_le="1998-09-02"):
        rows, n, sec = (Q1Row * 64)(), C.c_int(), C.c_double()
        self._check(self.lib.oracle_q1(lineitem, date_le.encode(), rows, 64, C.byref(n), C.byref(sec)))
        out = []
        for r in rows[: n.value]:
            out.append({"l_returnflag": r.returnflag, "l_linestatus": r.linestatus, "sum_qty": r.sum_qty,
                        "sum_base_price": r.sum_base_price, "sum_disc_price": i128(r.sum_disc_price),
                        "sum_charge": i128(r.sum_charge), "avg_qty": i128(r.avg_qty), "avg_price": i128(r.avg_price),
                        "avg_disc": i128(r.avg_disc), "count_order": r.count})
        return out, sec.value

    def q3(self, customer, orders, lineitem, segment="BUILDING", date="1995-03-15"):
        rows, n, sec = (Q3Row * 16)(), C.c_int(), C.c_double()
        self._check(self.lib.oracle_q3(customer, orders, lineitem, segment.encode(), date.encode(), rows, 16, C.byref(n), C.byref(sec)))
        return [{"l_orderkey": r.orderkey, "revenue": i128(r.revenue), "o_orderdate": r.orderdate, "o_shippriority": r.shippriority}
                for r in rows[: n.value]], sec.value

    def q5(self, customer, orders, lineitem, supplier, nation, region, region_name="ASIA", date_ge="1994-01-01", date_lt="1995-01-01"):
        rows, n, sec = (Q5Row * 32)(), C.c_int(), C.c_double()
        self._check(self.lib.oracle_q5(customer, orders, lineitem, supplier, nation, region, region_name.encode(), date_ge.encode(),
                                       date_lt.encode(), rows, 32, C.byref(n), C.byref(sec)))
        return [{"n_name": r.name.decode(), "revenue": i128(r.revenue)} for r in rows[: n.value]], sec.value

    def q9(self, part, supplier, lineitem, partsupp, orders, nation, needle="green"):
        rows, n, sec = (Q9Row * 256)(), C.c_int(), C.c_double()
        self._check(self.lib.oracle_q9(part, supplier, lineitem, partsupp, orders, nation, needle.encode(), rows, 256, C.byref(n), C.byref(sec)))
        return [{"nation": r.nation.decode(), "o_year": r.year, "sum_profit": i128(r.sum_profit)} for r in rows[: n.value]], sec.value

    # oracle twins prepared for the next widening step (no GPU operator yet)
    def q4(self, orders, lineitem, date_ge="1993-07-01", date_lt="1993-10-01"):
        rows, n, sec = (CountRow * 16)(), C.c_int(), C.c_double()
        self._check(self.lib.oracle_q4(orders, lineitem, date_ge.encode(), date_lt.encode(), rows, 16, C.byref(n), C.byref(sec)))
        return [{"o_orderpriority": r.name.decode(), "order_count": r.a} for r in rows[: n.value]], sec.value

    def q12(self, orders, lineitem, mode1="MAIL", mode2="SHIP", date_ge="1994-01-01", date_lt="1995-01-01"):
        rows, n, sec = (CountRow * 16)(), C.c_int(), C.c_double()
        self._check(self.lib.oracle_q12(orders, lineitem, mode1.encode(), mode2.encode(), date_ge.encode(), date_lt.encode(), rows, 16, C.byref(n), C.byref(sec)))
        return [{"l_shipmode": r.name.decode(), "high_line_count": r.a, "low_line_count": r.b} for r in rows[: n.value]], sec.value

    def q18(self, customer, orders, lineitem, quantity_gt=300):
        rows, n, sec = (Q18Row * 128)(), C.c_int(), C.c_double()
        self._check(self.lib.oracle_q18(customer, orders, lineitem, quantity_gt, rows, 128, C.byref(n), C.byref(sec)))
        return [{"c_name": r.name.decode(), "c_custkey": r.custkey, "o_orderkey": r.orderkey, "o_orderdate": r.orderdate, "o_totalprice": r.totalprice,
                 "sum_quantity": r.sum_quantity} for r in rows[: n.value]], sec.value
