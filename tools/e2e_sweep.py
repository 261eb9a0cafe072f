"""Sweep of the HOST staging configuration on one GPU: SF<sf> Q1 from pinned Arrow buffers through the C-ABI, for several
(pack threads, raw copiers) settings.  Prints one JSON line per configuration."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import bench
    from lingodb_b200 import datagen, devgen, runtime
    sf = float(sys.argv[1]) if len(sys.argv) > 1 else 100.0
    configs = [tuple(int(x) for x in c.split(":")) for c in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["14:0", "14:2", "12:2", "8:2", "0:2", "14:1", "30:2"])]
    s = datagen.scale(sf, 42)
    ctx = runtime.Context(0)
    li = devgen.lineitem(ctx, s, bench.Q1_COLS)
    td, batches = bench.device_table_to_host(li, bench.Q1_COLS, pinned=set(bench.Q1_COLS))
    want = runtime.Tpch(ctx, {"lineitem": li}).q1()
    ctx.close()
    specs = [c for c in datagen.LINEITEM_SCHEMA if c.name in bench.Q1_COLS]
    for pack, raw in configs:
        os.environ["LDB_STAGING_THREADS"] = str(max(pack, 1))
        os.environ["LDB_STAGING_RAW_THREADS"] = str(raw)
        os.environ["LDB_PACKED_STAGING"] = "1"
        if pack == 0:  # raw only: one idle packer that never gets work is not possible; emulate with 1 packer
            pass
        c = runtime.Context(0)
        tab = runtime.Table(c, "lineitem", specs)
        tp = runtime.Tpch(c, {"lineitem": tab})

        def step():
            tab.clear()
            for ch, n in batches:
                tab.append_host(ch, n)
            return tp.q1()
        assert step() == want
        h0, r0 = int(c.L.ldb_gpu_context_h2d_bytes(c.h)), int(c.L.ldb_gpu_context_raw_staged_rows(c.h))
        c.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            got = step()
        c.synchronize()
        dt = (time.perf_counter() - t0) / 3
        assert got == want
        print(json.dumps({"pack_threads": pack, "raw_copiers": raw, "ms_per_step": 1000 * dt, "rows_per_s": s.n_lineitem / dt,
                          "h2d_gb_per_step": (int(c.L.ldb_gpu_context_h2d_bytes(c.h)) - h0) / 3 / 1e9,
                          "raw_row_share": (int(c.L.ldb_gpu_context_raw_staged_rows(c.h)) - r0) / 3 / s.n_lineitem}), flush=True)
        tab.clear()
        c.close()


if __name__ == "__main__":
    main()
