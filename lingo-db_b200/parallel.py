"""Multi-GPU plumbing (one process per GPU, torch.distributed): how relations are split across ranks
and how small partial aggregates are merged.  The reference has no counterpart (single process;
SURVEY §2 "Parallelism strategies"): scans shard by row range with no data-path collective, only
the 4-group partial tables (≈9 KB) are all-gathered and folded by a merge kernel (K7).

The table image exchanged is exactly what ldb_gpu_groupby_export writes:
    int32 state[cap] | int32 keys[cap][2] | uint64 acc[cap][8][2]      (cap = table capacity)
"""
import ctypes as C
from typing import List, Tuple

import numpy as np

from . import datagen

MAX_KEYS, MAX_AGGS = 2, 8


def order_range(s: datagen.GenScale, rank: int, world: int) -> Tuple[int, int, int, int]:
    """Orders [o_lo, o_hi) and their lineitem rows [r_lo, r_hi) owned by `rank` (contiguous, exhaustive, disjoint)."""
    o_lo, o_hi = s.n_orders * rank // world, s.n_orders * (rank + 1) // world
    L = datagen.lib()
    return o_lo, o_hi, int(L.ldbgen_order_first_line(C.byref(s), o_lo)), int(L.ldbgen_order_first_line(C.byref(s), o_hi))


def row_range(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    return n_rows * rank // world, n_rows * (rank + 1) // world


def image_bytes(cap: int) -> int:
    return cap * 4 + cap * MAX_KEYS * 4 + cap * MAX_AGGS * 2 * 8


def pack_image(cap: int, groups: List[Tuple[Tuple[int, int], List[int]]]) -> np.ndarray:
    """groups: [((k0, k1), [agg0, agg1, … as python ints (128-bit two's complement)])] → image bytes."""
    state = np.zeros(cap, np.int32)
    keys = np.zeros((cap, MAX_KEYS), np.int32)
    acc = np.zeros((cap, MAX_AGGS, 2), np.uint64)
    for i, ((k0, k1), aggs) in enumerate(groups):
        state[i] = 2
        keys[i] = (k0, k1)
        for a, v in enumerate(aggs):
            v &= (1 << 128) - 1
            acc[i, a, 0] = v & 0xFFFFFFFFFFFFFFFF
            acc[i, a, 1] = v >> 64
    return np.concatenate([state.view(np.uint8), keys.reshape(-1).view(np.uint8), acc.reshape(-1).view(np.uint8)])


def unpack_image(img: np.ndarray, cap: int):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    state = img[: cap * 4].view(np.int32)
    keys = img[cap * 4: cap * 4 + cap * MAX_KEYS * 4].view(np.int32).reshape(cap, MAX_KEYS)
    acc = img[cap * 4 + cap * MAX_KEYS * 4:].view(np.uint64).reshape(cap, MAX_AGGS, 2)
    out = []
    for i in range(cap):
        if state[i] == 2:
            out.append(((int(keys[i, 0]), int(keys[i, 1])), [(int(acc[i, a, 1]) << 64) | int(acc[i, a, 0]) for a in range(MAX_AGGS)]))
    return out


def merge_images_host(images: List[np.ndarray], cap: int):
    """Reference semantics of the K7 merge (groupMergeImagesKernel): sums mod 2^128 per key."""
    total = {}
    for img in images:
        for key, aggs in unpack_image(img, cap):
            cur = total.setdefault(key, [0] * MAX_AGGS)
            for a in range(MAX_AGGS):
                cur[a] = (cur[a] + aggs[a]) & ((1 << 128) - 1)
    return total


def _signed(v: int, bits: int) -> int:
    v &= (1 << bits) - 1
    return v - (1 << bits) if v >> (bits - 1) else v


def q1_rows_from_groups(total) -> list:
    """Host finish of Q1 from merged groups (mirror of ldb_tpch_q1_finish): avg = (sum * 10^19) sdiv count."""
    rows = []
    for (k0, k1), a in sorted(total.items()):
        sum_qty, sum_base, sum_disc, cnt = _signed(a[0], 64), _signed(a[1], 64), _signed(a[4], 64), _signed(a[5], 64)

        def avg(x):
            q = abs(x * 10**19) // cnt
            return q if x >= 0 else -q

        rows.append({"l_returnflag": k0, "l_linestatus": k1, "sum_qty": sum_qty, "sum_base_price": sum_base, "sum_disc_price": _signed(a[2], 128),
                     "sum_charge": _signed(a[3], 128), "avg_qty": avg(sum_qty), "avg_price": avg(sum_base), "avg_disc": avg(sum_disc), "count_order": cnt})
    return rows


class Comm:
    """Peer-mapped exchange over NVLink (include/ldb_gpu.h "multi-GPU", csrc/peer.cu): every rank's symmetric heap is mapped
    into its peers through CUDA IPC; collectives are kernels that store into peer HBM and publish a flag.  torch.distributed
    only carries the 64-byte handles at start-up (`exchange` may be any callable bytes → [bytes per rank])."""

    def __init__(self, ctx, rank: int, world: int, user_bytes: int = 0, exchange=None, connect: bool = True):
        from . import capi
        self.ctx, self.rank, self.world, self.L = ctx, rank, world, ctx.L
        self.h = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        e = capi.Error()
        capi.check(self.L.ldb_gpu_comm_create(ctx.h, rank, world, int(user_bytes), C.byref(self.h), handle, C.byref(e)), e)
        if world > 1 and connect:
            if exchange is None:
                import torch.distributed as dist

                def exchange(b):
                    out = [None] * world
                    dist.all_gather_object(out, b)
                    return out
            handles = exchange(bytes(handle))
            blob = (C.c_uint8 * (64 * world)).from_buffer_copy(b"".join(handles))
            capi.check(self.L.ldb_gpu_comm_connect(self.h, blob, C.byref(e)), e)

    @classmethod
    def local_group(cls, ctxs, user_bytes: int = 0):
        """All ranks inside ONE process (tests; contexts may share a device): peers are wired by pointer, no IPC."""
        from . import capi
        comms = [cls(c, r, len(ctxs), user_bytes, connect=False) for r, c in enumerate(ctxs)]
        arr = (C.c_void_p * len(comms))(*[c.h for c in comms])
        e = capi.Error()
        capi.check(comms[0].L.ldb_gpu_comm_connect_local(arr, len(comms), C.byref(e)), e)
        return comms

    def close(self):
        if self.h:
            self.L.ldb_gpu_comm_destroy(self.h)
            self.h = C.c_void_p()

    def barrier(self):
        from . import capi
        e = capi.Error()
        capi.check(self.L.ldb_gpu_comm_barrier(self.h, C.byref(e)), e)

    def allmerge(self, state):
        """K7 over NVLink: afterwards every rank's group state holds the merged groups of all ranks."""
        from . import capi
        e = capi.Error()
        capi.check(self.L.ldb_gpu_groupby_allmerge(state, self.h, C.byref(e)), e)

    def check(self):
        from . import capi
        e = capi.Error()
        capi.check(self.L.ldb_gpu_comm_check(self.h, C.byref(e)), e)

    def heap(self):
        n = C.c_int64()
        p = self.L.ldb_gpu_comm_heap(self.h, C.byref(n))
        return int(p or 0), int(n.value)


def allgather_merge_state(ctx, state, world: int, rank: int, bufs: dict):
    """NCCL path (kept as the comparison arm of bench.py --merge nccl): export → all_gather_into_tensor → merge kernel."""
    import torch
    import torch.distributed as dist

    from . import capi
    L = ctx.L
    nbytes = int(L.ldb_gpu_groupby_export_bytes(state))
    if "send" not in bufs or bufs["send"].numel() != nbytes:
        dev = torch.device("cuda", ctx.device)
        bufs["send"] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        bufs["recv"] = torch.empty(nbytes * world, dtype=torch.uint8, device=dev)
    if "stream" not in bufs:
        bufs["stream"] = torch.cuda.ExternalStream(int(L.ldb_gpu_context_stream(ctx.h)), device=torch.device("cuda", ctx.device))
    e = capi.Error()
    capi.check(L.ldb_gpu_groupby_export(state, C.c_void_p(bufs["send"].data_ptr()), C.byref(e)), e)
    # torch sees the context's compute stream as its current stream: NCCL orders itself after the export and the merge
    # kernel after NCCL with stream events only — no host synchronisation inside the step
    with torch.cuda.stream(bufs["stream"]):
        dist.all_gather_into_tensor(bufs["recv"], bufs["send"])
    capi.check(L.ldb_gpu_groupby_merge_exported(state, C.c_void_p(bufs["recv"].data_ptr()), world, rank, C.byref(e)), e)


def q9_sharded(ctx, tpch, world: int, rank: int, bufs: dict, name_contains: str = "green", comm: "Comm" = None):
    """Q9 with lineitem ⋈ orders co-partitioned by order range (each rank's `tpch` holds its lineitem/orders shard and
    replicas of part, partsupp, supplier, nation): per-rank pipelines → merge of the group tables across ranks (peer-mapped
    all-merge kernel when a Comm is given, else NCCL all-gather + K7)."""
    from . import runtime
    st = tpch.q9_partial(name_contains)
    if world > 1:
        if comm is not None:
            comm.allmerge(st)
        else:
            allgather_merge_state(ctx, st, world, rank, bufs)
    rows = tpch.q9_finish(st)
    runtime.state_destroy(ctx, st)
    return rows


# ------------------------------------------------------------------------------------------------ repartitioned joins (C++ drivers)
# (round 1 orchestrated this plan from Python with NCCL all-to-alls and host synchronisations between the phases; it is now
#  csrc/tpch_plans.cpp + csrc/peer.cu: fused partition → NVLink peer stores, device-side barriers, no host round trip)
def q5_heap_bytes(ctx, n_orders_total: int, n_lineitem_total: int, world: int) -> int:
    return int(ctx.L.ldb_tpch_q5_repartitioned_heap_bytes(int(n_orders_total), int(n_lineitem_total), world))


def q5_repartitioned_peer(ctx, tpch, comm: "Comm", n_orders_total: int, n_lineitem_total: int, region_name="ASIA", date_ge="1994-01-01", date_lt="1995-01-01"):
    """The C++ driver (include/ldb_tpch.h ldb_tpch_q5_repartitioned): fused partition + NVLink peer stores, device-side barriers,
    Bloom OR through peer loads, peer all-merge — no NCCL, no host synchronisation inside the data path.  Returns (rows, stats)."""
    from . import capi
    rows, n, st, e = (capi.Q5Row * 25)(), C.c_int32(), capi.Q5ShuffleStats(), capi.Error()
    capi.check(ctx.L.ldb_tpch_q5_repartitioned(ctx.h, C.byref(tpch.t), comm.h, region_name.encode(), date_ge.encode(), date_lt.encode(), int(n_orders_total), int(n_lineitem_total),
                                               rows, C.byref(n), C.byref(st), C.byref(e)), e)
    out = [{"n_name": tpch.nation_names[r.n_nationkey], "revenue": r.revenue.value()} for r in rows[: n.value]]
    out.sort(key=lambda r: (-r["revenue"], r["n_name"]))
    return out, {k: int(getattr(st, k)) for k, _ in capi.Q5ShuffleStats._fields_}


def q9_heap_bytes(ctx, n_orders_total: int, n_lineitem_total: int, world: int) -> int:
    return int(ctx.L.ldb_tpch_q9_repartitioned_heap_bytes(int(n_orders_total), int(n_lineitem_total), world))


def q9_repartitioned_peer(ctx, tpch, comm: "Comm", n_orders_total: int, n_lineitem_total: int, name_contains: str = "green"):
    """ldb_tpch_q9_repartitioned: orders hash-partitioned across the ranks, lineitem contributions shipped to the owner of their order
    (K10 + K11 peer stores, device barriers, peer all-merge).  Returns (rows, stats)."""
    from . import capi
    rows, n, st, e = (capi.Q9Row * 1024)(), C.c_int32(), capi.Q5ShuffleStats(), capi.Error()
    capi.check(ctx.L.ldb_tpch_q9_repartitioned(ctx.h, C.byref(tpch.t), comm.h, name_contains.encode(), int(n_orders_total), int(n_lineitem_total), rows, 1024, C.byref(n),
                                               C.byref(st), C.byref(e)), e)
    return tpch._q9_rows(rows, n.value), {k: int(getattr(st, k)) for k, _ in capi.Q5ShuffleStats._fields_}


def _materialize(ctx, table, out_columns, widths, capacity, dev, **kw):
    """Run a K8 (scan → filters → [probe] → compacted columns) pipeline into fresh device buffers; regrow if the estimate was too small."""
    import torch

    from . import runtime
    while True:
        bufs = [torch.empty((capacity, 16), dtype=torch.uint8, device=dev) if w == 16 else torch.empty(capacity, dtype=torch.int32, device=dev) for w in widths]
        count = torch.zeros(1, dtype=torch.int64, device=dev)
        torch.cuda.synchronize(dev)
        runtime.run_pipeline(ctx, "scan_materialize", table, out_columns=out_columns, out_buffers=[t.data_ptr() for t in bufs], out_capacity=capacity,
                             out_count=count.data_ptr(), **kw)
        ctx.synchronize()
        n = int(count.item())
        if n <= capacity:
            return bufs, n
        capacity = int(n * 1.1) + 1024


class _CudaArray:
    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}


def _device_view(ptr: int, n_int32: int, dev):
    """torch view (no copy) of device memory owned by libldb_gpu.so, so NCCL can work on it in place."""
    import torch
    return torch.as_tensor(_CudaArray(ptr, n_int32), device=dev)
