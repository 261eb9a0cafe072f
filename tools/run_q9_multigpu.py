#!/usr/bin/env python
"""Q9 across GPUs (torchrun --nproc-per-node N tools/run_q9_multigpu.py --sf 300): lineitem and orders sharded by the same
order range (their join is co-partitioned, so no exchange), part/partsupp/supplier replicated, the 175-group partial tables
all-gathered over NCCL and merged on the device (K7).  --check compares with the single-GPU plan on rank 0 (needs the whole
table on one GPU: SF <= 100 with other ranks' shards also resident)."""
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
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = runtime.Context(local)
s = datagen.scale(a.sf, 42)
o_lo, o_hi, r_lo, r_hi = parallel.order_range(s, rank, world)
cols = ["l_orderkey", "l_partkey", "l_suppkey", "l_quantity", "l_extendedprice", "l_discount"]
small = devgen.small_tables(ctx)
tabs = {"lineitem": devgen.lineitem(ctx, s, cols, row_begin=r_lo, n_rows=r_hi - r_lo, batch_rows=1 << 29), "orders": devgen.orders(ctx, s, row_begin=o_lo, n_rows=o_hi - o_lo),
        "supplier": devgen.supplier(ctx, s), "part": devgen.part(ctx, s), "partsupp": devgen.partsupp(ctx, s), **small}
tp = runtime.Tpch(ctx, tabs)
bufs, times = {}, []
for i in range(a.reps):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rows = parallel.q9_sharded(ctx, tp, world, rank, bufs)
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t0)
tm = torch.tensor([min(times[1:] or times)], dtype=torch.float64, device=f"cuda:{local}")
if world > 1:
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
ok = None
if a.check and rank == 0:
    full = dict(tabs)
    full["lineitem"] = devgen.lineitem(ctx, s, cols, batch_rows=1 << 29)
    full["orders"] = devgen.orders(ctx, s)
    ok = runtime.Tpch(ctx, full).q9() == rows
if rank == 0:
    scanned = s.n_lineitem + s.n_orders + 4 * s.n_part + s.n_part + s.n_supplier + 25  # logical rows, replicated sides counted once
    print(json.dumps({"query": "q9_sharded", "sf": a.sf, "gpus": world, "seconds": float(tm.item()), "rows_per_s": scanned / float(tm.item()), "rows_scanned": scanned,
                      "matches_single_gpu_plan": ok, "timing": "wall clock around the whole plan incl. the NCCL all-gather + merge, max over ranks, best of reps-1",
                      "groups": len(rows), "rows_head": rows[:3]}))
if world > 1:
    dist.destroy_process_group()
