"""A/B of the filter-shape instantiations of the join kernels (ldb_gpu_set_filter_specialisation) and of the pause between tile-barrier
polls (ldb_gpu_set_poll_pause) at SF<sf> on one GPU: Q3 / Q5 / Q9 times and their kernel-family breakdown per variant
(specialise : producer pause ns : consumer pause ns), the list repeated <rounds> times; every result is checked against the first."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from lingodb_b200 import datagen, devgen, runtime
    sf = float(sys.argv[1]) if len(sys.argv) > 1 else 100.0
    rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    s = datagen.scale(sf, 42)
    ctx = runtime.Context(0)
    cols = ["l_orderkey", "l_partkey", "l_suppkey", "l_quantity", "l_extendedprice", "l_discount", "l_shipdate"]
    tabs = {"lineitem": devgen.lineitem(ctx, s, cols), "orders": devgen.orders(ctx, s), "customer": devgen.customer(ctx, s), "supplier": devgen.supplier(ctx, s),
            "part": devgen.part(ctx, s), "partsupp": devgen.partsupp(ctx, s), **devgen.small_tables(ctx)}
    tp = runtime.Tpch(ctx, tabs)
    fams = ["join_build", "join_probe_agg", "join_probe2_groupby", "join_star_probe_groupby", "join_topk", "table_init", "column_range"]
    base = {}
    variants = [(0, 0, 0), (1, 0, 0), (1, 100, 0), (1, 300, 0), (1, 100, 40), (0, 100, 0)]
    if len(sys.argv) > 3:
        variants = [tuple(int(x) for x in v.split(":")) for v in sys.argv[3].split(",")]
    for r in range(rounds):
        for on, prod_ns, cons_ns in variants:
            ctx.L.ldb_gpu_set_filter_specialisation(on)
            ctx.L.ldb_gpu_set_poll_pause(prod_ns, cons_ns)
            out = {"specialise": on, "producer_ns": prod_ns, "consumer_ns": cons_ns, "round": r}
            for name, fn in (("q3", tp.q3), ("q5", tp.q5), ("q9", tp.q9)):
                res = fn()
                if name in base:
                    assert res == base[name], (name, on)
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
    ctx.L.ldb_gpu_set_filter_specialisation(1)
    ctx.L.ldb_gpu_set_poll_pause(0, 0)


if __name__ == "__main__":
    main()
