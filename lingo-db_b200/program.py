"""Expression trees → register programs of the generic pipeline (include/ldb_gpu.h "program pipelines", csrc/program.cu).

An expression is a nested tuple:
  ("col", name) ("const", int) ("f64", float)
  ("add"|"sub"|"mul"|"div", a, b) ("neg", a)                 exact i128 arithmetic; decimal scales are the writer's job
  ("cmp", "<"|"<="|"="|"!="|">"|">=", a, b) ("between", x, lo, hi)
  ("and", a, b) ("or", a, b) ("not", a) ("isnull", a)       SQL three-valued logic
  ("case", cond, a, b)                                       cond is true ? a : b
  ("i2f", a) ("fadd"|"fsub"|"fmul"|"fdiv", a, b) ("fcmp", op, a, b)
  ("strcmp", op, column, "constant") ("like", "prefix"|"suffix"|"contains", column, "text") ("strkey8", column)
  ("year", a)                                                extract(year from date32)
  ("probe", join_table_state, key)                           payload, or NULL when the key is absent (semi / anti / mark / outer joins)
This is test/bench plumbing over the C-ABI, like runtime.py; in a LingoDB build the sub-operator lowering would emit LdbInstr lists."""
import ctypes as C
import struct
from typing import Dict, List, Optional

from . import capi
from .capi import Error, check

OPS = dict(load=1, const=2, add=3, sub=4, mul=5, div=6, neg=7, cmp=8, **{"and": 9, "or": 10, "not": 11}, isnull=12, select=13, i2f=14, fadd=15, fsub=16, fmul=17,
           fdiv=18, fcmp=19, strcmp=20, strlike=21, year=22, probe=23, strkey8=24)
CMP = {"=": 0, "!=": 1, "<": 2, "<=": 3, ">": 4, ">=": 5}
AGG = dict(sum=1, sum_f64=2, count=3, count_star=4, min=5, max=6, min_f64=7, max_f64=8, any=9)
LIKE = dict(prefix=0, suffix=1, contains=2)
SINK_HASHAGG, SINK_JOIN_BUILD, SINK_MATERIALIZE = 1, 2, 3


class Builder:
    def __init__(self):
        self.instr, self.columns, self.consts, self.strings, self.tables = [], [], [], [], []
        self._cache, self._next = {}, 0

    def _reg(self):
        r = self._next
        self._next += 1
        if r >= 48:
            raise ValueError("program needs more than 48 registers")
        return r

    def _col(self, name):
        if name not in self.columns:
            self.columns.append(name)
        return self.columns.index(name)

    def _emit(self, op, a=0, b=0, arg=0):
        r = self._reg()
        self.instr.append((OPS[op], r, a, b, arg))
        return r

    def _const(self, v: int):
        v &= (1 << 128) - 1
        if v not in self.consts:
            self.consts.append(v)
        return self.consts.index(v)

    def _string(self, s: str):
        if s not in self.strings:
            self.strings.append(s)
        return self.strings.index(s)

    def expr(self, e) -> int:
        key = repr(e) if not (isinstance(e, tuple) and e and e[0] == "probe") else None
        if key is not None and key in self._cache:
            return self._cache[key]
        k = e[0]
        if k == "col":
            r = self._emit("load", arg=self._col(e[1]))
        elif k == "const":
            r = self._emit("const", arg=self._const(int(e[1])))
        elif k == "f64":
            r = self._emit("const", arg=self._const(struct.unpack("<q", struct.pack("<d", float(e[1])))[0] & 0xFFFFFFFFFFFFFFFF))
        elif k in ("add", "sub", "mul", "div", "and", "or", "fadd", "fsub", "fmul", "fdiv"):
            a, b = self.expr(e[1]), self.expr(e[2])
            r = self._emit(k, a, b)
        elif k in ("neg", "not", "isnull", "i2f", "year"):
            r = self._emit(k, self.expr(e[1]))
        elif k in ("cmp", "fcmp"):
            a, b = self.expr(e[2]), self.expr(e[3])
            r = self._emit(k, a, b, CMP[e[1]])
        elif k == "between":
            return self.expr(("and", ("cmp", ">=", e[1], e[2]), ("cmp", "<=", e[1], e[3])))
        elif k == "case":
            c, a, b = self.expr(e[1]), self.expr(e[2]), self.expr(e[3])
            r = self._emit("select", a, b, c)
        elif k == "strcmp":
            r = self._emit("strcmp", self._col(e[2]), CMP[e[1]], self._string(e[3]))
        elif k == "like":
            r = self._emit("strlike", self._col(e[2]), LIKE[e[1]], self._string(e[3]))
        elif k == "strkey8":
            r = self._emit("strkey8", self._col(e[1]))
        elif k == "probe":
            if e[1] not in self.tables:
                self.tables.append(e[1])
            r = self._emit("probe", self.expr(e[2]), 0, self.tables.index(e[1]))
        else:
            raise ValueError(f"unknown expression {k}")
        if key is not None:
            self._cache[key] = r
        return r


def _desc(ctx, table, b: Builder, filter_reg: int):
    keep = []
    d = capi.ProgramDesc()
    d.source = table.h
    cols = [c.encode() for c in b.columns]
    arr = (C.c_char_p * max(1, len(cols)))(*cols)
    d.n_columns, d.columns = len(cols), arr
    ins = (capi.Instr * max(1, len(b.instr)))(*[capi.Instr(*i) for i in b.instr])
    d.n_instr, d.instr = len(b.instr), ins
    cs = (capi.I128 * max(1, len(b.consts)))(*[capi.I128(v & 0xFFFFFFFFFFFFFFFF, (v >> 64) - (1 << 64 if v >> 127 else 0)) for v in b.consts])
    d.n_consts, d.consts = len(b.consts), cs
    ss = [s.encode() for s in b.strings]
    sarr = (C.c_char_p * max(1, len(ss)))(*ss)
    d.n_strings, d.strings = len(ss), sarr
    tarr = (C.c_void_p * max(1, len(b.tables)))(*[t.value if isinstance(t, C.c_void_p) else t for t in b.tables])
    d.n_tables, d.tables = len(b.tables), tarr
    d.filter_reg = filter_reg
    keep += [cols, arr, ins, cs, ss, sarr, tarr]
    return d, keep


def hashagg_state(ctx, n_keys: int, agg_kinds: List[str], expected_groups: int) -> C.c_void_p:
    aggs = (capi.ProgAgg * max(1, len(agg_kinds)))(*[capi.ProgAgg(AGG[k], 0) for k in agg_kinds])
    s, e = C.c_void_p(), Error()
    check(ctx.L.ldb_gpu_hashagg_create(ctx.h, n_keys, len(agg_kinds), aggs, int(expected_groups), C.byref(s), C.byref(e)), e)
    return s


def group_by(ctx, table, keys: list, aggs: list, where=None, expected_groups: int = 1024, state=None) -> C.c_void_p:
    """aggs: [(kind, expr | None)].  Returns the hash-aggregation state (pass `state` to accumulate further tables into it)."""
    b = Builder()
    f = b.expr(where) if where is not None else -1
    kregs = [b.expr(k) for k in keys]
    aregs = [b.expr(x) if x is not None else 0 for _, x in aggs]
    st = state or hashagg_state(ctx, len(keys), [k for k, _ in aggs], expected_groups)
    d, keep = _desc(ctx, table, b, f)
    d.sink_kind, d.sink = SINK_HASHAGG, st
    d.n_keys = len(keys)
    for i, r in enumerate(kregs):
        d.key_regs[i] = r
    d.n_aggs = len(aggs)
    for i, ((kind, _), r) in enumerate(zip(aggs, aregs)):
        d.aggs[i] = capi.ProgAgg(AGG[kind], r)
    e = Error()
    check(ctx.L.ldb_gpu_run_program(ctx.h, C.byref(d), C.byref(e)), e)
    return st


def read_groups(ctx, state, max_rows: int = 1 << 22, f64_aggs=()) -> list:
    """[(keys tuple with None for NULL, aggs list with None for NULL)] — order unspecified."""
    rows = (capi.HashAggRow * max_rows)()
    n, e = C.c_int64(), Error()
    check(ctx.L.ldb_gpu_hashagg_read(state, rows, max_rows, C.byref(n), C.byref(e)), e)
    if n.value > max_rows:
        raise ValueError(f"{n.value} groups, buffer holds {max_rows}")
    nk = None
    out = []
    for r in rows[: n.value]:
        out.append((r.keys[:], r.key_null_mask, r.agg_valid_mask, [(a.lo, a.hi) for a in r.aggs]))
    return out


def decode_groups(raw, n_keys: int, n_aggs: int, f64_aggs=()):
    res = {}
    for keys, knull, avalid, aggs in raw:
        k = tuple(None if (knull >> i) & 1 else int(keys[i]) for i in range(n_keys))
        vals = []
        for a in range(n_aggs):
            if not (avalid >> a) & 1:
                vals.append(None)
            elif a in f64_aggs:
                vals.append(struct.unpack("<d", struct.pack("<Q", aggs[a][0]))[0])
            else:
                vals.append((int(aggs[a][1]) << 64) | int(aggs[a][0]))
        res[k] = vals
    return res


def build_join(ctx, table, join_state, key, payload=None, where=None):
    b = Builder()
    f = b.expr(where) if where is not None else -1
    kr = b.expr(key)
    pr = b.expr(payload) if payload is not None else -1
    d, keep = _desc(ctx, table, b, f)
    d.sink_kind, d.sink = SINK_JOIN_BUILD, join_state
    d.build_key_reg, d.build_payload_reg = kr, pr
    e = Error()
    check(ctx.L.ldb_gpu_run_program(ctx.h, C.byref(d), C.byref(e)), e)


def materialize(ctx, table, outs: list, where=None) -> C.c_void_p:
    """Returns a DEVICE table handle with columns c0..cN (raw i128 cells + validity bytes)."""
    b = Builder()
    f = b.expr(where) if where is not None else -1
    regs = [b.expr(x) for x in outs]
    d, keep = _desc(ctx, table, b, f)
    d.sink_kind = SINK_MATERIALIZE
    d.n_out = len(regs)
    for i, r in enumerate(regs):
        d.out_regs[i] = r
    out = C.c_void_p()
    d.out_table = C.pointer(out)
    e = Error()
    check(ctx.L.ldb_gpu_run_program(ctx.h, C.byref(d), C.byref(e)), e)
    return out


class RawTable:
    """Handle-only wrapper (tables created by the library: exported groups, materialised rows)."""

    def __init__(self, ctx, h):
        self.ctx, self.h = ctx, h

    @property
    def num_rows(self):
        return int(self.ctx.L.ldb_gpu_table_num_rows(self.h))

    def order_by(self, column: str, descending=False, limit=-1):
        n = self.num_rows
        ids = (C.c_int64 * max(1, n))()
        m, e = C.c_int64(), Error()
        check(self.ctx.L.ldb_gpu_table_order_by(self.h, column.encode(), int(descending), limit, ids, C.byref(m), C.byref(e)), e)
        return list(ids[: m.value])

    def gather(self, column: str, row_ids: list, cell_bytes=16):
        n = len(row_ids)
        ids = (C.c_int64 * max(1, n))(*row_ids)
        buf = (C.c_uint8 * max(1, n * cell_bytes))()
        valid = (C.c_uint8 * max(1, n))()
        e = Error()
        check(self.ctx.L.ldb_gpu_table_gather(self.h, column.encode(), ids, n, buf, valid, C.byref(e)), e)
        raw = bytes(buf)
        out = []
        for i in range(n):
            if not valid[i]:
                out.append(None)
            else:
                out.append(int.from_bytes(raw[i * cell_bytes:(i + 1) * cell_bytes], "little", signed=True))
        return out

    def destroy(self):
        if self.h:
            self.ctx.L.ldb_gpu_table_destroy(self.h)
            self.h = C.c_void_p()


def groups_table(ctx, state, name="groups") -> RawTable:
    t, e = C.c_void_p(), Error()
    check(ctx.L.ldb_gpu_hashagg_to_table(state, name.encode(), C.byref(t), C.byref(e)), e)
    return RawTable(ctx, t)
