#!/usr/bin/env python
"""Run one TPC-H query a few times on device-generated tables (used under ncu; see profiles/)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lingodb_b200 import datagen, devgen, runtime  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--query", default="q1")
ap.add_argument("--sf", type=float, default=10)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
ctx = runtime.Context(0)
s = datagen.scale(a.sf, 42)
cols = ["l_orderkey", "l_suppkey", "l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_returnflag", "l_linestatus", "l_shipdate"]
if a.query == "q9":  # only the referenced columns: SF300 lineitem is 1.8 G rows x 60 B = 108 GB of the 180 GB
    cols = ["l_orderkey", "l_partkey", "l_suppkey", "l_quantity", "l_extendedprice", "l_discount"]
tabs = {"lineitem": devgen.lineitem(ctx, s, cols, batch_rows=1 << 29)}
if a.query == "q9":
    tabs.update({"orders": devgen.orders(ctx, s), "supplier": devgen.supplier(ctx, s), "part": devgen.part(ctx, s), "partsupp": devgen.partsupp(ctx, s),
                 **devgen.small_tables(ctx)})
    print("rows scanned per run:", sum(tabs[k].num_rows for k in ("lineitem", "orders", "supplier", "part", "partsupp")) + 25, flush=True)
if a.query in ("q3", "q5"):
    tabs.update({"orders": devgen.orders(ctx, s), "customer": devgen.customer(ctx, s), "supplier": devgen.supplier(ctx, s), **devgen.small_tables(ctx)})
tp = runtime.Tpch(ctx, tabs)
fn = getattr(tp, a.query)
ctx.kernel_time_reset(True)
for i in range(a.reps):
    ctx.synchronize()
    t0 = time.perf_counter()
    res = fn()
    ctx.synchronize()
    print(f"{a.query} sf={a.sf:g} rep {i}: {1000 * (time.perf_counter() - t0):.3f} ms wall", flush=True)
for fam in ("scan_reduce", "scan_groupby", "join_build", "join_probe_agg", "join_probe2_groupby", "join_star_probe_groupby", "join_topk", "table_init"):
    ms, n = ctx.kernel_time(fam)
    if n:
        print(f"  kernels {fam}: {n} launches, {ms / n:.3f} ms avg")
print(res[:3] if isinstance(res, list) else res)
ctx.close()
