"""Sweep of the tile-pipeline depth / build tile size of the join kernels at SF<sf> on one GPU (ldb_gpu_set_tuning).
Prints per configuration the Q3 / Q5 / Q9 times and their kernel-family breakdown; every result is checked against the first."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from lingodb_b200 import datagen, devgen, runtime
    sf = float(sys.argv[1]) if len(sys.argv) > 1 else 100.0
    s = datagen.scale(sf, 42)
    ctx = runtime.Context(0)
    cols = ["l_orderkey", "l_partkey", "l_suppkey", "l_quantity", "l_extendedprice", "l_discount", "l_shipdate"]
    tabs = {"lineitem": devgen.lineitem(ctx, s, cols), "orders": devgen.orders(ctx, s), "customer": devgen.customer(ctx, s), "supplier": devgen.supplier(ctx, s),
            "part": devgen.part(ctx, s), "partsupp": devgen.partsupp(ctx, s), **devgen.small_tables(ctx)}
    tp = runtime.Tpch(ctx, tabs)
    fams = ["join_build", "join_probe_agg", "join_probe2_groupby", "join_star_probe_groupby", "join_topk", "table_init", "column_range"]
    base = {}
    configs = [(3, 3, 3, 2, 2), (2, 2, 2, 2, 1), (3, 2, 2, 2, 2), (3, 4, 4, 3, 2)]
    if len(sys.argv) > 2:
        configs = [tuple(int(x) for x in c.split(":")) for c in sys.argv[2].split(",")]
    for cfg in configs:
        ctx.L.ldb_gpu_set_tuning(*cfg)
        out = {"config": dict(zip(("stages_build", "stages_probe_agg", "stages_probe2", "stages_star", "rpt_build"), cfg))}
        for name, fn in (("q3", tp.q3), ("q5", tp.q5), ("q9", tp.q9)):
            res = fn()
            if name in base:
                assert res == base[name], (name, cfg)
            base.setdefault(name, res)
            fn()
            ctx.synchronize()
            ctx.kernel_time_reset(True)
            ctx.timer_start()
            for _ in range(5):
                fn()
            ms = ctx.timer_stop() / 5
            k = {}
            for f in fams:
                kms, kn = ctx.kernel_time(f)
                if kn:
                    k[f] = round(kms / 5, 3)
            ctx.kernel_time_reset(False)
            out[name] = {"ms": round(ms, 3), "kernels": k}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
