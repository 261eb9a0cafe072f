"""torchrun --nproc-per-node N tools/peer_selftest.py — the peer-mapped collectives across processes (CUDA IPC + NVLink):
every rank runs Q1 and Q9 on its order-range shard, all-merges the group tables over NVLink and compares the full result with
the CPU oracle over the whole table.  Used by tests/test_gpu_round2.py when >= 2 GPUs are visible and by hand on 8."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from lingodb_b200 import datagen, devgen, parallel, runtime
    from oracle import oracle as O
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sf = float(os.environ.get("SELFTEST_SF", "0.2"))
    s = datagen.scale(sf, seed=17)
    cols = ["l_orderkey", "l_partkey", "l_suppkey", "l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_returnflag", "l_linestatus", "l_shipdate"]
    ctx = runtime.Context(local)
    comm = parallel.Comm(ctx, rank, world, user_bytes=max(parallel.q5_heap_bytes(ctx, s.n_orders, s.n_lineitem, world), parallel.q9_heap_bytes(ctx, s.n_orders, s.n_lineitem, world)))
    o_lo, o_hi, r_lo, r_hi = parallel.order_range(s, rank, world)
    tabs = {"lineitem": devgen.lineitem(ctx, s, cols, row_begin=r_lo, n_rows=r_hi - r_lo), "orders": devgen.orders(ctx, s, row_begin=o_lo, n_rows=o_hi - o_lo),
            "supplier": devgen.supplier(ctx, s), "part": devgen.part(ctx, s), "partsupp": devgen.partsupp(ctx, s), "customer": devgen.customer(ctx, s), **devgen.small_tables(ctx)}
    tp = runtime.Tpch(ctx, tabs)
    host = datagen.tpch(sf, seed=17, lineitem_columns=cols, with_parts=True)
    o = O.Oracle("auto", workers=4)
    oh = {k: o.table(v) for k, v in host.items()}
    want_q1 = o.q1(oh["lineitem"])[0]
    want_q9 = o.q9(oh["part"], oh["supplier"], oh["lineitem"], oh["partsupp"], oh["orders"], oh["nation"])[0]
    want_q5 = o.q5(oh["customer"], oh["orders"], oh["lineitem"], oh["supplier"], oh["nation"], oh["region"])[0]
    for it in range(3):
        st = tp.q1_partial()
        comm.allmerge(st)
        got = tp.q1_finish(st)
        runtime.state_destroy(ctx, st)
        assert got == want_q1, f"rank {rank} iteration {it}: Q1 differs"
        got9 = parallel.q9_sharded(ctx, tp, world, rank, {}, comm=comm)
        assert got9 == want_q9, f"rank {rank} iteration {it}: Q9 differs"
        got5, st5 = parallel.q5_repartitioned_peer(ctx, tp, comm, s.n_orders, s.n_lineitem)
        assert got5 == want_q5, f"rank {rank} iteration {it}: repartitioned Q5 differs"
        got9r, st9 = parallel.q9_repartitioned_peer(ctx, tp, comm, s.n_orders, s.n_lineitem)
        assert got9r == want_q9, f"rank {rank} iteration {it}: repartitioned Q9 differs"
    comm.barrier()
    comm.check()
    dist.barrier()
    if rank == 0:
        print(f"peer selftest ok: world {world}, SF{sf:g}, Q1 {len(want_q1)} groups, Q9 {len(want_q9)} groups", flush=True)
    comm.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
