"""Arrow IPC files <-> the table layout the backend scans (storage → HBM staging).

LingoDB persists a table as one Arrow IPC *file* per table (`<dbDir>/<table>.arrow`): `storeTable` / `loadTable` write and read
record batches with arrow::ipc::MakeFileWriter / RecordBatchFileReader (src/runtime/storage/LingoDBTable.cpp:27-54), and every
batch becomes a TableChunk whose ArrayViews point INTO the Arrow buffers (:200-225).  This module does the same on this side of
the C-ABI: a record batch is turned into `datagen.TableData` chunks that are zero-copy numpy views of the Arrow buffers
(array offsets respected), which `Context.table_from_host` stages to HBM batch by batch and the oracle scans in place.

Physical types accepted are the ones LingoDBTable.cpp:122-195 produces for the hot path: int32, int64, date32,
decimal128(p, s), fixed_size_binary(4) (char(1)), utf8 with int32 offsets.  Batches with nulls are rejected, like the C-ABI does.
"""
from typing import Iterable, List, Optional

import numpy as np
import pyarrow as pa
import pyarrow.ipc

from .datagen import ColumnSpec, TableData

MAX_CHUNK_ROWS = 1 << 20  # ArrayView::maxNullCount (ArrowView.h:9): one shared all-valid bitmap covers 2^20 rows


def _spec(field: pa.Field) -> ColumnSpec:
    t = field.type
    if pa.types.is_int32(t):
        return ColumnSpec(field.name, "int32")
    if pa.types.is_int64(t):
        return ColumnSpec(field.name, "int64")
    if pa.types.is_date32(t):
        return ColumnSpec(field.name, "date32")
    if pa.types.is_decimal128(t):
        return ColumnSpec(field.name, "decimal128", t.precision, t.scale)
    if pa.types.is_fixed_size_binary(t) and t.byte_width == 4:
        return ColumnSpec(field.name, "fsb4")
    if pa.types.is_string(t):
        return ColumnSpec(field.name, "utf8")
    raise TypeError(f"column {field.name}: Arrow type {t} is not on the GPU hot path")


def _view(arr: pa.Array, spec: ColumnSpec):
    """numpy views over the buffers of one Arrow array (no copy)."""
    if arr.null_count:
        raise ValueError(f"column {spec.name}: batches with nulls are not supported on the GPU path")
    bufs, off, n = arr.buffers(), arr.offset, len(arr)
    if spec.phys == "utf8":
        offs = np.frombuffer(bufs[1], dtype=np.int32, count=off + n + 1)[off:]
        data = np.frombuffer(bufs[2], dtype=np.uint8) if bufs[2] is not None and bufs[2].size else np.zeros(1, np.uint8)
        return offs, data
    if spec.phys == "decimal128":
        return np.frombuffer(bufs[1], dtype=np.uint8, count=16 * (off + n)).reshape(-1, 16)[off:]
    dt = np.int64 if spec.phys == "int64" else np.int32  # int32 / date32 / fixed_size_binary(4) are 4-byte cells
    return np.frombuffer(bufs[1], dtype=dt, count=off + n)[off:]


def tabledata_from_batches(name: str, schema: pa.Schema, batches: Iterable[pa.RecordBatch], columns: Optional[List[str]] = None,
                           chunk_rows: int = MAX_CHUNK_ROWS) -> TableData:
    fields = [f for f in schema if columns is None or f.name in columns]
    specs = [_spec(f) for f in fields]
    t = TableData(name, specs)
    chunk_rows = min(chunk_rows, MAX_CHUNK_ROWS)
    for batch in batches:
        for b in range(0, batch.num_rows, chunk_rows):
            piece = batch.slice(b, min(chunk_rows, batch.num_rows - b))  # zero-copy: only the array offset moves
            t.chunks.append({s.name: _view(piece.column(piece.schema.get_field_index(s.name)), s) for s in specs})
            t.chunk_rows.append(piece.num_rows)
    t._arrow_keepalive = batches  # the numpy views borrow the Arrow buffers
    return t


def read_ipc(path: str, name: Optional[str] = None, columns: Optional[List[str]] = None, chunk_rows: int = MAX_CHUNK_ROWS) -> TableData:
    """`<table>.arrow` (Arrow IPC file, memory-mapped) → TableData; twin of loadTable (LingoDBTable.cpp:27-38)."""
    src = pa.memory_map(path, "r")
    reader = pa.ipc.open_file(src)
    batches = [reader.get_batch(i) for i in range(reader.num_record_batches)]
    import os
    t = tabledata_from_batches(name or os.path.splitext(os.path.basename(path))[0], reader.schema, batches, columns, chunk_rows)
    t._arrow_keepalive = (src, reader, batches)
    return t


# ---- the other direction (tests, and exporting generator tables in the reference's on-disk format)
def _arrow_array(spec: ColumnSpec, v, n: int) -> pa.Array:
    if spec.phys == "utf8":
        offs, data = v
        return pa.Array.from_buffers(pa.string(), n, [None, pa.py_buffer(np.ascontiguousarray(offs[: n + 1])), pa.py_buffer(np.ascontiguousarray(data))])
    typ = {"int32": pa.int32(), "int64": pa.int64(), "date32": pa.date32(), "fsb4": pa.binary(4)}.get(spec.phys) or pa.decimal128(spec.precision, spec.scale)
    return pa.Array.from_buffers(typ, n, [None, pa.py_buffer(np.ascontiguousarray(v))])


def to_arrow_batches(t: TableData):
    schema = pa.schema([pa.field(s.name, _arrow_array(s, (np.zeros(1, np.int32), np.zeros(1, np.uint8)) if s.phys == "utf8" else
                                                     np.zeros((0, 16), np.uint8) if s.phys == "decimal128" else np.zeros(0, np.int64 if s.phys == "int64" else np.int32), 0).type)
                        for s in t.columns])
    batches = [pa.RecordBatch.from_arrays([_arrow_array(s, chunk[s.name], n) for s in t.columns], schema=schema) for chunk, n in zip(t.chunks, t.chunk_rows)]
    return schema, batches


def write_ipc(path: str, t: TableData):
    """TableData → Arrow IPC file; twin of storeTable (LingoDBTable.cpp:39-54)."""
    schema, batches = to_arrow_batches(t)
    with pa.OSFile(path, "wb") as sink:
        with pa.ipc.new_file(sink, schema) as w:
            for b in batches:
                w.write_batch(b)


# ---- the C++ path (libldb_arrow_io.so, include/ldb_arrow_io.h): file → backend table without Python touching the buffers
_arrow_lib = None


def _lib():
    import ctypes as C

    from . import build, capi
    global _arrow_lib
    if _arrow_lib is None:
        L = C.CDLL(build.build_arrow_io())
        P, E = C.c_void_p, C.POINTER(capi.Error)
        L.ldb_arrow_file_open.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int32, C.POINTER(P), E]
        L.ldb_arrow_file_num_columns.argtypes = [P]
        L.ldb_arrow_file_schema.restype = C.POINTER(capi.ColumnSchema)
        L.ldb_arrow_file_schema.argtypes = [P]
        L.ldb_arrow_file_num_batches.argtypes = [P]
        L.ldb_arrow_file_load.argtypes = [P, P, P, C.c_int64, C.POINTER(C.c_int64), E]
        L.ldb_arrow_file_close.argtypes = [P]
        _arrow_lib = L
    return _arrow_lib


class ArrowFile:
    """An open `<table>.arrow` (memory-mapped by Arrow C++).  schema() works without a GPU; load() stages the file into a backend
    table — the column cache: later pipelines read it from HBM, the file only has to stay open until the table is cleared."""

    def __init__(self, path: str, columns: Optional[List[str]] = None):
        import ctypes as C

        from . import capi
        self.L, self.h = _lib(), C.c_void_p()
        e = capi.Error()
        arr = (C.c_char_p * max(1, len(columns or [])))(*[c.encode() for c in (columns or [])])
        capi.check(self.L.ldb_arrow_file_open(path.encode(), arr, len(columns or []), C.byref(self.h), C.byref(e)), e)

    def schema(self) -> List[ColumnSpec]:
        from . import capi
        inv = {v: k for k, v in capi.PHYS.items()}
        s = self.L.ldb_arrow_file_schema(self.h)
        return [ColumnSpec(s[i].name.decode(), inv[s[i].type], s[i].precision, s[i].scale) for i in range(self.L.ldb_arrow_file_num_columns(self.h))]

    @property
    def num_batches(self) -> int:
        return int(self.L.ldb_arrow_file_num_batches(self.h))

    def load(self, ctx, name: str, max_rows_per_batch: int = 1 << 24):
        """→ runtime.Table staged from the mapped file (ldb_arrow_file_load → ldb_gpu_table_append_batch)."""
        import ctypes as C

        from . import capi, runtime
        tab = runtime.Table(ctx, name, self.schema())
        n, e = C.c_int64(), capi.Error()
        fn = C.cast(ctx.L.ldb_gpu_table_append_batch, C.c_void_p)
        capi.check(self.L.ldb_arrow_file_load(self.h, tab.h, fn, max_rows_per_batch, C.byref(n), C.byref(e)), e)
        tab._keep.append(self)  # the views point into the mapping
        return tab

    def close(self):
        if self.h:
            self.L.ldb_arrow_file_close(self.h)
            self.h = None
