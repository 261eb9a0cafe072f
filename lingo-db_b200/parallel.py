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


def allgather_merge_state(ctx, state, world: int, rank: int, bufs: dict):
    """NCCL path used by bench.py: export → all_gather_into_tensor → merge kernel, all on the device."""
    import torch
    import torch.distributed as dist

    from . import capi
    L = ctx.L
    nbytes = int(L.ldb_gpu_groupby_export_bytes(state))
    if "send" not in bufs or bufs["send"].numel() != nbytes:
        dev = torch.device("cuda", ctx.device)
        bufs["send"] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        bufs["recv"] = torch.empty(nbytes * world, dtype=torch.uint8, device=dev)
    e = capi.Error()
    capi.check(L.ldb_gpu_groupby_export(state, C.c_void_p(bufs["send"].data_ptr()), C.byref(e)), e)
    ctx.synchronize()  # the export ran on the context's compute stream, NCCL runs on torch's
    dist.all_gather_into_tensor(bufs["recv"], bufs["send"])
    torch.cuda.current_stream().synchronize()
    capi.check(L.ldb_gpu_groupby_merge_exported(state, C.c_void_p(bufs["recv"].data_ptr()), world, rank, C.byref(e)), e)
