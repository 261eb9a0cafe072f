#!/usr/bin/env python
"""Phase timing of the repartitioned Q5 plan at world = 1 (profiles/r1_q5_repartitioned_n2_sf100_v2.json quotes it): wraps
parallel._materialize / _partition / _all_to_all and runtime.run_pipeline with host-synchronised wall-clock timers."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lingodb_b200 import datagen, devgen, parallel, runtime
ctx = runtime.Context(0)
s = datagen.scale(100, 42)
cols = ["l_orderkey", "l_suppkey", "l_extendedprice", "l_discount"]
tabs = {"lineitem": devgen.lineitem(ctx, s, cols), "orders": devgen.orders(ctx, s), "customer": devgen.customer(ctx, s), "supplier": devgen.supplier(ctx, s), **devgen.small_tables(ctx)}
# monkeypatch phase timers
marks = []
orig_mat, orig_part, orig_a2a = parallel._materialize, parallel._partition, parallel._all_to_all
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(*a, **k); torch.cuda.synchronize(); marks.append((name, 1000 * (time.perf_counter() - t0))); return r
    return w
parallel._materialize = timed("materialize", orig_mat); parallel._partition = timed("partition", orig_part); parallel._all_to_all = timed("a2a", orig_a2a)
orig_rp = runtime.run_pipeline
def rp(ctx_, kind, *a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = orig_rp(ctx_, kind, *a, **k); ctx_.synchronize(); marks.append((kind, 1000 * (time.perf_counter() - t0))); return r
runtime.run_pipeline = rp
for i in range(3):
    marks.clear(); torch.cuda.synchronize(); t0 = time.perf_counter()
    rows, st = parallel.q5_repartitioned(ctx, tabs, 1, 0, s.n_orders)
    torch.cuda.synchronize(); tot = 1000 * (time.perf_counter() - t0)
print("total ms", tot); print([(n, round(t, 3)) for n, t in marks]); print(st)
