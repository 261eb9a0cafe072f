#!/usr/bin/env python
"""bench.py — TPC-H SF100 Q1 (lineitem scan + 2-key hash aggregation) rows/s on N B200s, plus Q3/Q5 (BASELINE.json's metric
names all three), Q6 and Q9 as side entries.

Contract: `python bench.py --gpus N --steps K --warmup W [--impl reference]`, one JSON line on rank 0.
  step    = one pass of the Q1 pipeline (scan → filter → expressions → group-by) over the lineitem partition resident in HBM,
            plus (N > 1) the peer-mapped all-merge of the 4-group partials over NVLink (one kernel, csrc/peer.cu).
  value   = lineitem rows of ALL ranks / max-over-ranks device time, inputs resident in HBM ("strong": SF100 is split across
            the ranks by order range).
  e2e     = the same step through the C-ABI with HOST (pinned) Arrow buffers: H2D staging of every column batch (compressed
            staging: host threads pack, the GPU unpacks — csrc/staging.cu) and the D2H result read are inside the timed region.
  roofline= scan_groupby kernel: algorithmic bytes (76 B/row) / CUDA-event kernel time.
  parity  = every timed leg is gated: the CUDA result over the FULL table must equal the CPU oracle's on the same bytes.
  cpu_baseline = the CPU oracle (reference runtime objects + restated pipelines, NOT the LLVM JIT) on the box's host cores
            over a bounded sample of the same table.
`--impl reference` times that CPU implementation alone on the same config (rank 0 only).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

Q1_COLS = ["l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_returnflag", "l_linestatus", "l_shipdate"]
Q1_BYTES_PER_ROW = 76  # 4 decimal128 + 2 fixed_size_binary(4) + date32
ALL_COLS = ["l_orderkey", "l_partkey", "l_suppkey"] + Q1_COLS
METRIC = "TPC-H SF100 Q1 rows/sec"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sf", type=float, default=100.0)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-batch-rows", type=int, default=1 << 24)
    ap.add_argument("--cpu-sample-sf", type=float, default=20.0)
    ap.add_argument("--merge", default="peer", choices=["peer", "nccl"], help="N>1: peer-mapped all-merge kernel (default) or NCCL all-gather + K7")
    ap.add_argument("--no-graph", action="store_true", help="issue every step eagerly instead of replaying the captured query (CUDA graph)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the full-size oracle gates (profiling runs only)")
    ap.add_argument("--no-extra", action="store_true", help="skip the Q6/Q3/Q5/Q9 entries")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def effective_cpus():
    """CPUs this container may burn (cgroup quota aware) — the 128-thread box gives a container 16."""
    try:
        from lingodb_b200 import capi
        return int(capi.lib().ldb_gpu_effective_cpus())
    except Exception:
        return os.cpu_count() or 1


def host_available_bytes():
    try:
        import psutil
        return int(psutil.virtual_memory().available)
    except Exception:
        return 64 << 30


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- reference arm
def host_lineitem(sf, seed, cols, n_rows, gen_rows=1 << 23):
    """lineitem prefix in host RAM as <= 2^20-row record batches.  Generated 8 Mi rows at a time by the host generator's
    thread team (one contiguous slice per thread), so the pages of every column are first-touched by threads spread over all
    cores of both sockets — the table ends up interleaved across the NUMA nodes at 64 Ki-row granularity."""
    from lingodb_b200 import datagen
    s = datagen.scale(sf, seed)
    n_rows = min(n_rows, s.n_lineitem)
    specs = [c for c in datagen.LINEITEM_SCHEMA if c.name in cols]
    td = datagen.TableData("lineitem", specs)
    b = 0
    while b < n_rows:
        n = min(gen_rows, n_rows - b)
        big = datagen.lineitem(s, cols, chunk_rows=n, row_begin=b, n_rows=n)
        arrs = big.chunks[0]
        for o in range(0, n, 1 << 20):
            m = min(1 << 20, n - o)
            td.chunks.append({k: v[o:o + m] for k, v in arrs.items()})
            td.chunk_rows.append(m)
        b += n
    return td, n_rows


def sweep_oracle_workers(o, run, cands=None):
    """Give the CPU side the worker count it is fastest with: a sweep around the CPUs the container may use.  Returns {workers: seconds}."""
    ncpu = effective_cpus()
    env = os.environ.get("ORACLE_PARALLELISM")
    if cands is None:  # around the CPU quota: more threads than the quota only get throttled (measured: 128 workers 9x slower than 16)
        cands = [int(env)] if env else sorted({max(1, int(ncpu * f)) for f in (2, 1.5, 1, 0.75, 0.5)}, reverse=True)
    seen = {}
    for w in cands:
        o.set_workers(w)
        run()
        seen[w] = min(run(), run())
    best = min(seen, key=seen.get)
    o.set_workers(best)
    return seen


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from lingodb_b200 import datagen
    from oracle import oracle as O
    s = datagen.scale(args.sf, args.seed)
    need = Q1_BYTES_PER_ROW * s.n_lineitem
    avail = host_available_bytes()
    # the full table when host RAM allows (it does on the 8xB200 box), else the largest prefix that leaves half of RAM free
    n_rows = s.n_lineitem if need < 0.6 * avail else int(0.5 * avail / Q1_BYTES_PER_ROW)
    t0 = time.perf_counter()
    t, n = host_lineitem(args.sf, args.seed, Q1_COLS, n_rows)
    gen_s = time.perf_counter() - t0
    o = O.Oracle("auto", workers=0)
    h = o.table(t)
    sweep = sweep_oracle_workers(o, lambda: o.q1(h)[1])
    for _ in range(args.warmup):
        o.q1(h)
    t0 = time.perf_counter()
    secs = []
    for _ in range(args.steps):
        _, sec = o.q1(h)
        secs.append(sec)
    wall = time.perf_counter() - t0
    total = sum(secs)
    value = n * args.steps / total
    sample = (f"the full SF{args.sf:g} lineitem table ({n} rows) per step" if n == s.n_lineitem else f"first {n} lineitem rows of the SF{args.sf:g} table per step (host RAM bound)") + \
        ", resident in host RAM, pages first-touched by generator threads on all cores (NUMA-interleaved)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000 * total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64/int128",
        "data": "synthetic", "config": {"workload": f"TPC-H SF{args.sf:g} Q1 (lineitem scan + 2-key hash aggregation)", "sf": args.sf, "lineitem_rows": n, "sample": sample,
                                        "timed_region": "pipelines only (reference executionTime)", "wall_s": wall, "generate_s": gen_s},
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": o.workers, "host_threads": os.cpu_count(), "kind": o.kind, "sample": sample,
                         "worker_sweep_s": {str(k): v for k, v in sweep.items()},
                         "note": "worker count = fastest of the sweep; reference runtime objects + restated pipelines; the MLIR/LLVM JIT cannot be built here"},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- our arm
def device_table_to_host(tab, want=None, pinned=(), batch_rows=1 << 24):
    """Copy a table's batches (DEVICE tensors from devgen, or host numpy chunks) to host RAM.  Returns (TableData with
    <= 2^20-row chunks for the oracle, [(chunk dict, rows)] of the fixed-width columns in <= `batch_rows` batches for the
    C-ABI's HOST path).  Columns in `pinned` land in pinned memory."""
    import numpy as np
    import torch

    from lingodb_b200 import datagen
    specs = [c for c in tab.columns if want is None or c.name in want]
    td = datagen.TableData(tab.name, specs)
    batches = []

    def to_host(v, pin):
        if isinstance(v, np.ndarray):
            return v
        if not pin:
            return v.cpu().numpy()
        h = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
        for b in range(0, v.shape[0], batch_rows):  # bounded transfers: one slice at a time
            h[b:b + batch_rows].copy_(v[b:b + batch_rows])
        return h.numpy()

    for item in tab._keep:
        if not isinstance(item, dict):
            continue
        host, n = {}, None
        for c in specs:
            v = item[c.name]
            if c.phys == "utf8":
                host[c.name] = (to_host(v[0], False), to_host(v[1], False))
                n = host[c.name][0].shape[0] - 1
            else:
                host[c.name] = to_host(v, c.name in pinned)
                n = host[c.name].shape[0]
        torch.cuda.synchronize()
        for b in range(0, n, 1 << 20):
            m = min(1 << 20, n - b)
            ch = {}
            for c in specs:
                a = host[c.name]
                ch[c.name] = (a[0][b:b + m + 1], a[1]) if c.phys == "utf8" else a[b:b + m]
            td.chunks.append(ch)
            td.chunk_rows.append(m)
        for b in range(0, n, batch_rows):
            m = min(batch_rows, n - b)
            batches.append(({c.name: host[c.name][b:b + m] for c in specs if c.phys != "utf8"}, m))
    return td, batches


def q1_sums(rows):
    """Q1 rows → {(flag, status): (sum_qty, sum_base, sum_disc_price, sum_charge, count)} — the exactly mergeable part."""
    return {(r["l_returnflag"], r["l_linestatus"]): (r["sum_qty"], r["sum_base_price"], r["sum_disc_price"], r["sum_charge"], r["count_order"]) for r in rows}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    ncpu = effective_cpus()
    # one staging engine per rank: the ranks of a box share the container's CPU quota
    os.environ.setdefault("LDB_STAGING_THREADS", str(max(2, ncpu - 2) if world == 1 else max(1, (ncpu - 2) // world)))

    from lingodb_b200 import datagen, devgen, parallel, runtime

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # NCCL's "NCCL version …" banner goes to stdout; the contract is ONE JSON line
        dist.init_process_group("nccl", device_id=dev)

    ctx = runtime.Context(local)
    L = ctx.L
    s = datagen.scale(args.sf, args.seed)
    comm = parallel.Comm(ctx, rank, world, user_bytes=max(parallel.q5_heap_bytes(ctx, s.n_orders, s.n_lineitem, world), parallel.q9_heap_bytes(ctx, s.n_orders, s.n_lineitem, world)) if not args.no_extra else 0) if world > 1 else None
    # strong scaling: SF-sized lineitem split by order range
    o_lo, o_hi, r_lo, r_hi = parallel.order_range(s, rank, world)
    my_rows = r_hi - r_lo
    extra = not args.no_extra
    cols = ALL_COLS if extra else Q1_COLS
    lineitem = devgen.lineitem(ctx, s, cols, row_begin=r_lo, n_rows=my_rows)
    tabs = {"lineitem": lineitem}
    tp = runtime.Tpch(ctx, tabs)
    total_rows = s.n_lineitem
    notes = []

    gather_bufs = {}

    def barrier():
        if world > 1:
            dist.barrier()

    def gather_objects(obj):
        if world == 1:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    use_graph = not args.no_graph and (world == 1 or args.merge == "peer")
    prepared = {}

    def step_resident():
        if use_graph:
            # the query is captured once (state init + scan/group-by kernel + peer all-merge) and replayed with one driver
            # call per step; the result read stays outside the graph
            if not prepared:
                ctx.graph_begin()
                prepared["state"] = tp.q1_partial()
                if world > 1:
                    comm.allmerge(prepared["state"])
                prepared["graph"] = ctx.graph_end()
            prepared["graph"].launch()
            return tp.q1_finish(prepared["state"], lazy=True)  # C rows in host memory; python dicts are built when they are compared
        st = tp.q1_partial()
        if world > 1:
            if args.merge == "peer":
                comm.allmerge(st)
            else:
                parallel.allgather_merge_state(ctx, st, world, rank, gather_bufs)
        rows = tp.q1_finish(st)
        runtime.state_destroy(ctx, st)
        return rows

    # ---- host copy of this rank's shard (pinned): input of the e2e leg AND of the full-size oracle gate
    avail = host_available_bytes()
    host_cols = Q1_COLS
    host_bytes = Q1_BYTES_PER_ROW * my_rows
    have_host = (not args.no_e2e or not args.no_parity) and host_bytes * world < 0.6 * avail
    if not have_host and not (args.no_e2e and args.no_parity):
        notes.append(f"host copy skipped: {host_bytes * world >> 30} GiB needed, {avail >> 30} GiB available")
    td_li = host_batches = None
    if have_host:
        td_li, host_batches = device_table_to_host(lineitem, host_cols, pinned=set(host_cols), batch_rows=args.e2e_batch_rows)

    # ---- parity gate 1 (full size): resident CUDA result == oracle on the same bytes, merged across ranks
    parity = {}
    oracle = None
    if have_host and not args.no_parity:
        from oracle import oracle as O
        oracle = O.Oracle("auto", workers=max(2, ncpu // world))
        oh_li = oracle.table(td_li)
        t0 = time.perf_counter()
        mine = q1_sums(oracle.q1(oh_li)[0])
        merged = {}
        for part in gather_objects(mine):
            for k, v in part.items():
                cur = merged.get(k, (0, 0, 0, 0, 0))
                merged[k] = tuple(a + b for a, b in zip(cur, v))
        got_rows = list(step_resident())
        if q1_sums(got_rows) != merged:
            raise SystemExit(f"PARITY FAILURE (Q1 resident, rank {rank}): CUDA {q1_sums(got_rows)} != oracle {merged}")
        for r in got_rows:  # avg = (sum * 10^19) sdiv count, recomputed from the merged exact sums
            for name, ssum in (("avg_qty", r["sum_qty"]), ("avg_price", r["sum_base_price"])):
                exp = abs(ssum * 10**19) // r["count_order"] * (1 if ssum >= 0 else -1)
                if r[name] != exp:
                    raise SystemExit(f"PARITY FAILURE (Q1 {name})")
        parity["q1_resident"] = {"ok": True, "rows": total_rows, "oracle": oracle.kind, "seconds": time.perf_counter() - t0,
                                 "what": "sums/counts of every group == oracle over the same bytes (all ranks' shards merged); averages recomputed"}

    # ---- warm-up + timed region (device-resident inputs, 45.6 GB at SF100 >> 126 MB L2)
    for _ in range(args.warmup):
        rows = step_resident()
    ctx.kernel_time_reset(True)
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local)
    barrier()
    ctx.synchronize()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    ctx.timer_start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rows = step_resident()
    ms_dev = ctx.timer_stop()
    ctx.synchronize()
    torch.cuda.synchronize()
    wall_ms = 1000 * (time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None
    barrier()
    launches = ctx.launch_count() - launches0
    k_ms, k_n = ctx.kernel_time("scan_groupby")
    m_ms, m_n = ctx.kernel_time("peer_group_allmerge" if args.merge == "peer" else "group_merge")
    ctx.kernel_time_reset(False)
    if comm:
        comm.check()
    # device time of the timed region, max over ranks
    tmax = torch.tensor([ms_dev, wall_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total, wall_total = float(tmax[0].item()), float(tmax[1].item())
    ms_per_step = ms_total / args.steps
    value = total_rows * args.steps / (ms_total / 1000)

    peak, peak_src = peaks()
    k_avg_ms = k_ms / max(k_n, 1)
    achieved = (Q1_BYTES_PER_ROW * my_rows) / (k_avg_ms / 1000) / 1e9 if k_avg_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "scanGroupByKernel<2 keys, 4 decimal cols, Q1 aggregates>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "frac_of": peak_src, "traffic": None, "algorithmic_bytes_per_launch": Q1_BYTES_PER_ROW * my_rows,
                "kernel_ms_avg": k_avg_ms, "kernel_launches_timed": k_n, "kernel_share_of_step": (k_ms / ms_dev) if ms_dev else None}
    if world > 1:
        roofline["merge_kernel_ms_avg"] = m_ms / max(m_n, 1)
        roofline["merge"] = "peer-mapped all-merge kernel over NVLink (csrc/peer.cu); its time includes waiting for the slowest rank" if args.merge == "peer" else "NCCL all_gather_into_tensor + K7"
    ncu = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(ncu):
        try:
            tr = json.load(open(ncu)).get("scan_groupby_q1")
            if tr:
                roofline["traffic"] = tr["dram_bytes_per_row"] * my_rows
                roofline["traffic_source"] = tr["source"]
        except (ValueError, KeyError):
            pass

    # ---- e2e: HOST (pinned) Arrow buffers → C-ABI staging → pipeline → result read
    e2e = None
    if not args.no_e2e and have_host:
        specs = [c for c in datagen.LINEITEM_SCHEMA if c.name in Q1_COLS]
        host_tab = runtime.Table(ctx, "lineitem", specs)
        tp_h = runtime.Tpch(ctx, {"lineitem": host_tab})
        h2d = Q1_BYTES_PER_ROW * my_rows

        def step_e2e():
            host_tab.clear()
            for chunk, n in host_batches:
                host_tab.append_host(chunk, n)
            st = tp_h.q1_partial()
            if world > 1 and args.merge == "peer":
                comm.allmerge(st)
            r = tp_h.q1_finish(st)
            runtime.state_destroy(ctx, st)
            return r

        rows_h = step_e2e()  # warm-up (starts the staging engine, allocates its slots)
        if rows_h != rows:  # parity gate 2: the staged (packed → unpacked) path gives the resident path's rows, bit for bit
            raise SystemExit(f"PARITY FAILURE (Q1 e2e, rank {rank}): host-staged result differs from the resident result")
        parity["q1_e2e"] = {"ok": True, "what": "rows through HOST staging == rows of the resident path (== oracle)"}
        barrier()
        ctx.synchronize()
        h2d0 = int(L.ldb_gpu_context_h2d_bytes(ctx.h))
        raw0 = int(L.ldb_gpu_context_raw_staged_rows(ctx.h))
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            rows_h = step_e2e()
        ctx.synchronize()
        e2e_s = time.perf_counter() - t0
        h2d_step = (int(L.ldb_gpu_context_h2d_bytes(ctx.h)) - h2d0) // args.e2e_steps
        raw_rows = (int(L.ldb_gpu_context_raw_staged_rows(ctx.h)) - raw0) // args.e2e_steps
        # column-cache hit: the staged table stays in HBM (LingoDBTable.cpp:294-305 ownership), the same call reads it again
        t0 = time.perf_counter()
        for _ in range(3):
            st = tp_h.q1_partial()
            if world > 1 and args.merge == "peer":
                comm.allmerge(st)
            rc = tp_h.q1_finish(st)
            runtime.state_destroy(ctx, st)
        cached_s = (time.perf_counter() - t0) / 3
        assert rc == rows
        te = torch.tensor([e2e_s, cached_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s, cached_s = float(te[0].item()), float(te[1].item())
        hsum = torch.tensor([h2d_step], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(hsum)
        e2e = {"value": total_rows * args.e2e_steps / e2e_s, "unit": "rows/s", "h2d_bytes_per_step": int(hsum.item()),
               "host_arrow_bytes_per_step": h2d * world,
               "staging": f"compressed staging: {os.environ['LDB_STAGING_THREADS']} host threads/rank (container CPU quota {ncpu}) pack 64Ki-value frame-of-reference blocks (1/2/4/8 B per value; "
                          "decimal128(12,2) → its low 8 bytes first) into pinned slots, one CUDA stream per thread copies them and a kernel unpacks into the staged layout; "
                          f"{os.environ.get('LDB_STAGING_RAW_THREADS', '2')} raw copiers fill idle link time with uncompressed Arrow cells (csrc/staging.cu)",
               "rows_shipped_raw_per_step_rank0": raw_rows, "host_cpu_quota": ncpu,
               "d2h_bytes_per_step": (64 * 140 + 16) * world, "steps": args.e2e_steps, "ms_per_step": 1000 * e2e_s / args.e2e_steps,
               "batch_rows": args.e2e_batch_rows, "host_memory": "pinned", "cold": True,
               "cached_value": total_rows / cached_s, "cached_note": "same C-ABI call with the table already staged (column-cache hit): no H2D, result D2H only; wall clock"}
        host_tab.clear()

    # ---- CPU baseline on a bounded sample of the same table (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and have_host:
        from oracle import oracle as O
        sample_sf = min(args.sf, args.cpu_sample_sf)
        s_s = datagen.scale(sample_sf, args.seed)
        n = min(my_rows, s_s.n_lineitem)
        td = datagen.TableData("lineitem", td_li.columns)
        acc = 0
        for ch, m in zip(td_li.chunks, td_li.chunk_rows):
            if acc >= n:
                break
            m2 = min(m, n - acc)
            td.chunks.append({k: v[:m2] for k, v in ch.items()})
            td.chunk_rows.append(m2)
            acc += m2
        o = O.Oracle("auto", workers=0)
        oh = o.table(td)
        sweep = sweep_oracle_workers(o, lambda: o.q1(oh)[1])
        sec = min(o.q1(oh)[1] for _ in range(3))
        cpu = {"value": n / sec, "unit": "rows/s", "cores": o.workers, "host_threads": os.cpu_count(), "kind": o.kind,
               "sample": f"first {n} lineitem rows of the SF{args.sf:g} table (the SF{sample_sf:g} prefix), best of 3 after the worker sweep, pipelines only",
               "seconds": sec, "worker_sweep_s": {str(k): v for k, v in sweep.items()},
               "note": "reference runtime objects + restated pipelines (oracle/), not the MLIR/LLVM JIT; pinned host copy of the device table"}

    # ---- the other §8 configs on the same resident tables (N = 1): Q3 and Q5 are BASELINE.json metric entries, Q6/Q9 ride along
    queries = None
    metrics = None
    if extra and world == 1:
        queries, metrics = side_queries(args, ctx, s, tabs, lineitem, td_li, oracle, peak, parity, notes)
    elif extra and world > 1:
        # The Q9 / Q5 multi-GPU plans ride along: a failure there (their own parity gates raise) must not take the gated Q1 line with it —
        # the plan is then reported as failed, without a number.  (Every rank sees the same merged rows, so a parity failure is rank-consistent.)
        try:
            queries = side_queries_multi(args, ctx, comm, s, tabs, rank, world, o_lo, o_hi, dev, notes)
        except (Exception, SystemExit) as ex:  # noqa: BLE001
            queries = {"error": f"{type(ex).__name__}: {ex}"[:600]}
            notes.append("side queries failed and are not reported: " + queries["error"])

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64/int128", "data": "synthetic",
            "config": {"workload": f"TPC-H SF{args.sf:g} Q1 (lineitem scan + 2-key hash aggregation) on {world}xB200", "sf": args.sf, "lineitem_rows": total_rows,
                       "rows_per_gpu": my_rows, "partitioning": ("order-range split; partial group tables merged by a peer-mapped NVLink kernel" if args.merge == "peer" else "order-range split; NCCL all-gather of the partials") if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (45.6 GB of columns per pass at SF100 vs 126 MB L2)", "arrow_layout": "decimal128 16 B/value, date32, fixed_size_binary(4)",
                       "generator": "deterministic TPC-H-shaped generator on device (csrc/datagen.cu), seed %d" % args.seed, "wall_ms_per_step": wall_total / args.steps,
                       "result_rows": len(rows), "launch": "captured query replayed as one CUDA graph per step (csrc/runtime.cpp ldb_gpu_graph_*)" if use_graph else "eager"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "parity": parity,
        }
        if metrics:
            line["metrics"] = metrics
        if queries:
            line["other_queries"] = queries
        if notes:
            line["notes"] = notes
        print(json.dumps(line), flush=True)
    if prepared:
        ctx.synchronize()
        prepared["graph"].destroy()
        runtime.state_destroy(ctx, prepared["state"])
    if comm:
        ctx.synchronize()
        barrier()
        comm.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


KERNEL_FAMILIES = ["scan_reduce", "scan_groupby", "join_build", "join_probe_agg", "join_probe2_groupby", "join_star_probe_groupby", "join_topk", "table_init",
                   "column_range", "materialize", "partition", "group_merge"]


def side_queries(args, ctx, s, tabs, lineitem, td_li, oracle, peak, parity, notes):
    """Q6/Q3/Q5/Q9 on one GPU: parity gate (full-size oracle where host RAM and time allow), then 5 timed repetitions with the
    per-kernel-family CUDA-event times of the timed region (roofline of the whole query = algorithmic bytes / query time)."""
    from lingodb_b200 import devgen, runtime
    tabs.update({"orders": devgen.orders(ctx, s), "customer": devgen.customer(ctx, s), "supplier": devgen.supplier(ctx, s), **devgen.small_tables(ctx)})
    tabs.update({"part": devgen.part(ctx, s), "partsupp": devgen.partsupp(ctx, s)})
    tpx = runtime.Tpch(ctx, tabs)
    n_ps = 4 * s.n_part
    scanned = {"q6": s.n_lineitem, "q3": s.n_lineitem + s.n_orders + s.n_customer, "q5": s.n_lineitem + s.n_orders + s.n_customer + s.n_supplier + 30,
               "q9": s.n_lineitem + s.n_orders + n_ps + s.n_part + s.n_supplier + 25}
    # Arrow physical widths of the referenced columns, each read once (p_name: 4 B offset + ~33 B text)
    algo = {"q6": 52 * s.n_lineitem, "q3": 40 * s.n_lineitem + 16 * s.n_orders + 21 * s.n_customer,
            "q5": 40 * s.n_lineitem + 12 * s.n_orders + 8 * s.n_customer + 8 * s.n_supplier,
            "q9": 60 * s.n_lineitem + 24 * n_ps + 8 * s.n_orders + 41 * s.n_part + 8 * s.n_supplier}
    # ---- parity gates at full size: host copies of the build sides + the lineitem key columns (pageable)
    gate = {}
    if oracle is not None and td_li is not None and not args.no_parity:
        try:
            t0 = time.perf_counter()
            td_keys, _ = device_table_to_host(lineitem, ["l_orderkey", "l_partkey", "l_suppkey"])
            for ch, kc in zip(td_li.chunks, td_keys.chunks):
                ch.update(kc)
            td_li.columns = list(td_li.columns) + [c for c in td_keys.columns]
            oh = {"lineitem": oracle.table(td_li)}
            for name in ("orders", "customer", "supplier", "nation", "region", "part", "partsupp"):
                oh[name] = oracle.table(device_table_to_host(tabs[name])[0])
            oracle.set_workers(max(2, effective_cpus()))
            gate["q6"] = lambda: oracle.q6(oh["lineitem"])[0]
            gate["q3"] = lambda: oracle.q3(oh["customer"], oh["orders"], oh["lineitem"])[0]
            gate["q5"] = lambda: oracle.q5(oh["customer"], oh["orders"], oh["lineitem"], oh["supplier"], oh["nation"], oh["region"])[0]
            gate["q9"] = lambda: oracle.q9(oh["part"], oh["supplier"], oh["lineitem"], oh["partsupp"], oh["orders"], oh["nation"])[0]
            notes.append(f"host copies for the Q3/Q5/Q9 gates: {time.perf_counter() - t0:.1f} s")
        except MemoryError:
            notes.append("Q3/Q5/Q9 gates skipped: host RAM")
    queries, metrics = {}, []
    for name, fn in (("q6", tpx.q6), ("q3", tpx.q3), ("q5", tpx.q5), ("q9", tpx.q9)):
        res = fn()
        if name in gate:  # parity gate in front of the timed leg
            t0 = time.perf_counter()
            exp = gate[name]()
            if res != exp:
                raise SystemExit(f"PARITY FAILURE ({name} at SF{args.sf:g}): CUDA {str(res)[:300]} != oracle {str(exp)[:300]}")
            parity[name] = {"ok": True, "oracle_seconds": time.perf_counter() - t0, "what": f"full result rows == oracle at SF{args.sf:g}"}
        fn()
        ctx.synchronize()
        reps = 5
        ctx.kernel_time_reset(True)
        ctx.timer_start()
        for _ in range(reps):
            res = fn()
        ms = ctx.timer_stop() / reps
        kern = {}
        for fam in KERNEL_FAMILIES:
            kms, kn = ctx.kernel_time(fam)
            if kn:
                kern[fam] = {"ms_per_query": kms / reps, "launches_per_query": kn / reps}
        ctx.kernel_time_reset(False)
        gbs = algo[name] / (ms / 1000) / 1e9
        queries[name] = {"ms": ms, "rows_per_s": scanned[name] / (ms / 1000), "rows_scanned": scanned[name], "kernels": kern,
                         "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "algorithmic_bytes": algo[name],
                                      "kernel_ms_sum": sum(k["ms_per_query"] for k in kern.values())},
                         "parity": "oracle ok" if name in parity else "not gated"}
        if name == "q5":
            queries[name]["result"] = [[r["n_name"], r["revenue"]] for r in res]
        if name in ("q3", "q5"):
            metrics.append({"metric": f"TPC-H SF{args.sf:g} {name.upper()} rows/sec", "value": scanned[name] / (ms / 1000), "unit": "rows/s", "ms_per_query": ms,
                            "roofline_frac": gbs / peak, "parity": queries[name]["parity"]})
    return queries, metrics


def side_queries_multi(args, ctx, comm, s, tabs, rank, world, o_lo, o_hi, dev, notes):
    """N > 1: Q9 with lineitem ⋈ orders co-partitioned by order range (no exchange), group tables merged over NVLink."""
    import torch
    import torch.distributed as dist

    from lingodb_b200 import devgen, parallel, runtime
    tabs.update({"orders": devgen.orders(ctx, s, row_begin=o_lo, n_rows=o_hi - o_lo), "customer": devgen.customer(ctx, s),
                 "supplier": devgen.supplier(ctx, s), **devgen.small_tables(ctx)})
    tabs.update({"part": devgen.part(ctx, s), "partsupp": devgen.partsupp(ctx, s)})
    tpx = runtime.Tpch(ctx, tabs)
    queries = {}
    res = parallel.q9_sharded(ctx, tpx, world, rank, {}, comm=comm)
    parallel.q9_sharded(ctx, tpx, world, rank, {}, comm=comm)
    ctx.synchronize()
    dist.barrier()
    reps = 3
    ctx.timer_start()
    for _ in range(reps):
        res = parallel.q9_sharded(ctx, tpx, world, rank, {}, comm=comm)
    ms = ctx.timer_stop() / reps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n_ps = 4 * s.n_part
    scanned = s.n_lineitem + s.n_orders + world * (n_ps + s.n_part + s.n_supplier + 25)
    queries["q9_sharded"] = {"ms": float(t.item()), "rows_per_s": scanned / (float(t.item()) / 1000), "rows_scanned": scanned, "groups": len(res),
                             "checksum_sum_profit": sum(r["sum_profit"] for r in res),
                             "plan": "lineitem and orders sharded by the same order range (co-partitioned join), part/partsupp/supplier replicated, peer-mapped all-merge of the 175-group tables"}
    comm.check()
    # ---- Q9 with orders HASH-partitioned (config 4): K10 + K11 peer stores; must give the co-partitioned plan's rows
    rows9, st9 = parallel.q9_repartitioned_peer(ctx, tpx, comm, s.n_orders, s.n_lineitem)
    if rows9 != res:
        raise SystemExit(f"PARITY FAILURE (repartitioned Q9, rank {rank}): differs from the co-partitioned plan")
    secs = []
    for _ in range(3):
        dist.barrier()
        ctx.synchronize()
        t0 = time.perf_counter()
        rows9, st9 = parallel.q9_repartitioned_peer(ctx, tpx, comm, s.n_orders, s.n_lineitem)
        secs.append(time.perf_counter() - t0)
    t9 = torch.tensor([min(secs[1:])], dtype=torch.float64, device=dev)
    dist.all_reduce(t9, op=dist.ReduceOp.MAX)
    st = torch.tensor([st9["orders_tuples_sent"], st9["lineitem_tuples_sent"], st9["shuffle_bytes_out"]], dtype=torch.int64, device=dev)
    dist.all_reduce(st)
    queries["q9_repartitioned"] = {"ms": 1000 * float(t9.item()), "rows_per_s": scanned / float(t9.item()), "rows_scanned": scanned, "orders_tuples_shuffled": int(st[0].item()),
                                   "lineitem_tuples_shuffled": int(st[1].item()), "shuffle_bytes_all_ranks": int(st[2].item()),
                                   "parity": "rows == the co-partitioned plan's (whose single-GPU twin is oracle-gated at SF100)",
                                   "plan": "orders hash-partitioned across the ranks (K10), lineitem contributions shipped to the owner of their order (K11), peer all-merge"}
    comm.check()
    # ---- Q5 with the orders ⋈ lineitem join REPARTITIONED across the ranks (BASELINE.json config 3): fused partition + NVLink peer
    # stores, device-side barriers, Bloom OR by peer loads, peer all-merge — C++ driver ldb_tpch_q5_repartitioned, no NCCL
    golden = os.path.join(ROOT, "tests", "golden", "bench_answers_sf100_seed42.json")
    want5 = json.load(open(golden)).get("q5") if (os.path.exists(golden) and args.sf == 100.0 and args.seed == 42) else None
    rows5, st5 = parallel.q5_repartitioned_peer(ctx, tpx, comm, s.n_orders, s.n_lineitem)
    if want5 is not None and [[r["n_name"], r["revenue"]] for r in rows5] != want5:
        raise SystemExit(f"PARITY FAILURE (repartitioned Q5, rank {rank}): {rows5} != golden {want5}")
    secs = []
    ctx.kernel_time_reset(True)
    for _ in range(4):
        dist.barrier()
        ctx.synchronize()
        t0 = time.perf_counter()
        rows5, st5 = parallel.q5_repartitioned_peer(ctx, tpx, comm, s.n_orders, s.n_lineitem)
        secs.append(time.perf_counter() - t0)
    send_ms, send_n = ctx.kernel_time("partition_send")
    ctx.kernel_time_reset(False)
    t5 = torch.tensor([min(secs[1:]), send_ms / max(send_n, 1) * 2], dtype=torch.float64, device=dev)
    dist.all_reduce(t5, op=dist.ReduceOp.MAX)
    stats = torch.tensor([st5["orders_tuples_sent"], st5["lineitem_tuples_sent"], st5["shuffle_bytes_out"]], dtype=torch.int64, device=dev)
    dist.all_reduce(stats)
    scanned5 = s.n_lineitem + s.n_orders + world * (s.n_customer + s.n_supplier + 30)
    q5_s = float(t5[0].item())
    queries["q5_repartitioned"] = {"ms": 1000 * q5_s, "rows_per_s": scanned5 / q5_s, "rows_scanned": scanned5,
                                   "timing": "wall clock around the C++ driver (one host synchronisation at the result read), max over ranks, best of 3",
                                   "orders_tuples_shuffled": int(stats[0].item()), "lineitem_tuples_shuffled": int(stats[1].item()), "shuffle_bytes_all_ranks": int(stats[2].item()),
                                   "partition_send_kernels_ms_per_query": float(t5[1].item()),
                                   "shuffle_gbs_during_send_kernels": int(stats[2].item()) / 1e9 / (float(t5[1].item()) / 1000) if float(t5[1].item()) > 0 else None,
                                   "note": "the send kernels are lineitem/orders SCANS (40 / 12 B per row read from HBM) that store ~4 % of the rows to peers; the NVLink volume is small by design (Bloom semi-join before the shuffle)",
                                   "parity": "golden answer of the oracle-gated single-GPU run" if want5 is not None else "not gated (no golden for this sf/seed)",
                                   "result": [[r["n_name"], r["revenue"]] for r in rows5]}
    comm.check()
    return queries


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
