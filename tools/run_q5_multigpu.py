#!/usr/bin/env python
"""Q5 with the orders ⋈ lineitem join repartitioned across GPUs (torchrun --nproc-per-node N tools/run_q5_multigpu.py --sf 100):
the C++ driver ldb_tpch_q5_repartitioned (fused partition → NVLink peer stores, device barriers, Bloom OR by peer loads, peer
all-merge).  --check compares with the single-GPU broadcast plan on rank 0."""
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
comm = parallel.Comm(ctx, rank, world, user_bytes=parallel.q5_heap_bytes(ctx, s.n_orders, s.n_lineitem, world))
o_lo, o_hi, r_lo, r_hi = parallel.order_range(s, rank, world)
cols = ["l_orderkey", "l_suppkey", "l_extendedprice", "l_discount"]
tabs = {"lineitem": devgen.lineitem(ctx, s, cols, row_begin=r_lo, n_rows=r_hi - r_lo), "orders": devgen.orders(ctx, s, row_begin=o_lo, n_rows=o_hi - o_lo),
        "customer": devgen.customer(ctx, s), "supplier": devgen.supplier(ctx, s), **devgen.small_tables(ctx)}
tp = runtime.Tpch(ctx, tabs)
times = []
for _ in range(a.reps):
    if world > 1:
        dist.barrier()
    ctx.synchronize()
    t0 = time.perf_counter()
    rows, stats = parallel.q5_repartitioned_peer(ctx, tp, comm, s.n_orders, s.n_lineitem)
    times.append(time.perf_counter() - t0)
tm = torch.tensor([min(times[1:] or times)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
ok = None
if a.check and rank == 0:
    full = dict(tabs)
    full["lineitem"] = devgen.lineitem(ctx, s, cols)
    full["orders"] = devgen.orders(ctx, s)
    ok = runtime.Tpch(ctx, full).q5() == rows
comm.check()
if rank == 0:
    print(json.dumps({"query": "q5_repartitioned", "sf": a.sf, "gpus": world, "seconds": float(tm.item()), "matches_single_gpu_plan": ok, "rank0_stats": stats, "result": rows}), flush=True)
if world > 1:
    dist.barrier()
comm.close()
ctx.close()
if world > 1:
    dist.destroy_process_group()
