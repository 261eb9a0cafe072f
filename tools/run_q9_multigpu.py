#!/usr/bin/env python
"""Q9 across GPUs (torchrun --nproc-per-node N tools/run_q9_multigpu.py --sf 300) — BASELINE.json config 4.
Two plans over the same shards (lineitem and orders split by order range, part / partsupp / supplier replicated):
  copartitioned   the lineitem ⋈ orders join is co-partitioned by construction: no exchange, peer all-merge of the 175-group tables
  repartitioned   orders HASH-partitioned across the ranks (K10 peer stores), lineitem contributions shipped to the owner of their
                  order (K11 peer stores), probes against the partition, all-merge — the plan for inputs that are not co-partitioned
Both must give the same rows (checked on every rank); --check additionally compares with the single-GPU plan on rank 0 (needs the whole
table on one GPU).  Prints one JSON line: time per plan (wall clock, max over ranks, best of reps-1), shuffle volume and the NVLink
rate while the send kernels run."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from lingodb_b200 import datagen, devgen, parallel, runtime  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sf", type=float, default=10)
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--check", action="store_true")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
ctx = runtime.Context(local)
s = datagen.scale(a.sf, 42)
comm = parallel.Comm(ctx, rank, world, user_bytes=parallel.q9_heap_bytes(ctx, s.n_orders, s.n_lineitem, world))
o_lo, o_hi, r_lo, r_hi = parallel.order_range(s, rank, world)
cols = ["l_orderkey", "l_partkey", "l_suppkey", "l_quantity", "l_extendedprice", "l_discount"]
tabs = {"lineitem": devgen.lineitem(ctx, s, cols, row_begin=r_lo, n_rows=r_hi - r_lo, batch_rows=1 << 29), "orders": devgen.orders(ctx, s, row_begin=o_lo, n_rows=o_hi - o_lo),
        "supplier": devgen.supplier(ctx, s), "part": devgen.part(ctx, s), "partsupp": devgen.partsupp(ctx, s), **devgen.small_tables(ctx)}
tp = runtime.Tpch(ctx, tabs)


def barrier():
    if world > 1:
        dist.barrier()


def timed(fn):
    out, times = None, []
    for _ in range(a.reps):
        barrier()
        ctx.synchronize()
        t0 = time.perf_counter()
        out = fn()
        ctx.synchronize()
        times.append(time.perf_counter() - t0)
    tm = torch.tensor([min(times[1:] or times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    return out, float(tm.item())


rows_c, sec_c = timed(lambda: parallel.q9_sharded(ctx, tp, world, rank, {}, comm=comm))
ctx.kernel_time_reset(True)
(rows_r, stats), sec_r = timed(lambda: parallel.q9_repartitioned_peer(ctx, tp, comm, s.n_orders, s.n_lineitem))
send_ms = sum(ctx.kernel_time(f)[0] for f in ("partition_send", "star_probe_send")) / a.reps
ctx.kernel_time_reset(False)
assert rows_r == rows_c, f"rank {rank}: the repartitioned plan differs from the co-partitioned plan"
tot = torch.tensor([stats["orders_tuples_sent"], stats["lineitem_tuples_sent"], stats["shuffle_bytes_out"]], dtype=torch.int64, device=dev)
sm = torch.tensor([send_ms], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(tot)
    dist.all_reduce(sm, op=dist.ReduceOp.MAX)
ok = None
if a.check and rank == 0:
    full = dict(tabs)
    full["lineitem"] = devgen.lineitem(ctx, s, cols, batch_rows=1 << 29)
    full["orders"] = devgen.orders(ctx, s)
    ok = runtime.Tpch(ctx, full).q9() == rows_c
comm.check()
if rank == 0:
    scanned = s.n_lineitem + s.n_orders + 4 * s.n_part + s.n_part + s.n_supplier + 25  # logical rows, replicated sides counted once
    remote = (world - 1) / world
    print(json.dumps({"query": "q9", "sf": a.sf, "gpus": world, "rows_scanned": scanned, "groups": len(rows_c), "plans_agree": True, "matches_single_gpu_plan": ok,
                      "copartitioned": {"seconds": sec_c, "rows_per_s": scanned / sec_c},
                      "repartitioned": {"seconds": sec_r, "rows_per_s": scanned / sec_r, "orders_tuples": int(tot[0].item()), "lineitem_tuples": int(tot[1].item()),
                                        "shuffle_bytes": int(tot[2].item()), "bytes_over_nvlink": int(tot[2].item() * remote), "send_kernels_ms": float(sm.item()),
                                        "nvlink_gbs_during_send_kernels": int(tot[2].item() * remote) / 1e9 / (float(sm.item()) / 1000) if float(sm.item()) > 0 else None},
                      "checksum_sum_profit": sum(r["sum_profit"] for r in rows_c),
                      "timing": "wall clock around the whole plan (builds of the replicated sides included), max over ranks, best of reps-1"}), flush=True)
barrier()
comm.close()
ctx.close()
if world > 1:
    dist.destroy_process_group()
