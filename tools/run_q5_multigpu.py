#!/usr/bin/env python
"""Q5 with the orders⋈lineitem join radix-partitioned across GPUs (torchrun --nproc-per-node N tools/run_q5_multigpu.py --sf 100).
Checks the result against the single-GPU plan on rank 0 when --check is given."""
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
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--check", action="store_true")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = runtime.Context(local)
s = datagen.scale(a.sf, 42)
o_lo, o_hi, r_lo, r_hi = parallel.order_range(s, rank, world)
cols = ["l_orderkey", "l_suppkey", "l_extendedprice", "l_discount"]
tabs = {"lineitem": devgen.lineitem(ctx, s, cols, row_begin=r_lo, n_rows=r_hi - r_lo), "orders": devgen.orders(ctx, s, row_begin=o_lo, n_rows=o_hi - o_lo),
        "customer": devgen.customer(ctx, s), "supplier": devgen.supplier(ctx, s), **devgen.small_tables(ctx)}
times = []
for i in range(a.reps):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rows, stats = parallel.q5_repartitioned(ctx, tabs, world, rank, s.n_orders)
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t0)
tm = torch.tensor([min(times[1:] or times)], dtype=torch.float64, device=f"cuda:{local}")
if world > 1:
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
ok = None
if a.check and rank == 0:
    full = {"lineitem": devgen.lineitem(ctx, s, cols), "orders": devgen.orders(ctx, s), "customer": tabs["customer"], "supplier": tabs["supplier"],
            "nation": tabs["nation"], "region": tabs["region"]}
    ok = runtime.Tpch(ctx, full).q5() == rows
if rank == 0:
    scanned = s.n_lineitem + s.n_orders + world * (s.n_customer + s.n_supplier)
    print(json.dumps({"query": "q5_repartitioned", "sf": a.sf, "gpus": world, "seconds": float(tm.item()), "rows_per_s": (s.n_lineitem + s.n_orders + s.n_customer + s.n_supplier) / float(tm.item()),
                      "matches_single_gpu_plan": ok, "rank0_stats": stats, "rows": rows}))
if world > 1:
    dist.destroy_process_group()
