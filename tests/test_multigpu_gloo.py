"""N>1 host logic on CPU: world_size-2 gloo run of the Q1 sharding protocol (order-range split, table-image
all-gather, K7 merge semantics, host finish).  The per-rank partial is computed by the CPU oracle here —
the GPU kernels are covered by the -m gpu tests; this covers partitioning and the exchange format."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CAP = 64


def _worker(rank, world, port, sf, seed, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lingodb_b200 import datagen, parallel
    from oracle import oracle as O
    s = datagen.scale(sf, seed)
    o_lo, o_hi, r_lo, r_hi = parallel.order_range(s, rank, world)
    mine = datagen.lineitem(s, chunk_rows=4096, row_begin=r_lo, n_rows=r_hi - r_lo)
    o = O.Oracle("port", workers=1)
    part, _ = o.q1(o.table(mine))
    groups = []
    for r in part:
        # invert the avg: the exchanged partial carries sums and counts (SimplifyAggregations: avg = sum / count)
        sum_disc = -(-(r["avg_disc"] * r["count_order"]) // 10**19)
        groups.append(((r["l_returnflag"], r["l_linestatus"]),
                       [r["sum_qty"], r["sum_base_price"], r["sum_disc_price"], r["sum_charge"], sum_disc, r["count_order"], 0, 0]))
    img = torch.from_numpy(parallel.pack_image(CAP, groups))
    gathered = [torch.empty_like(img) for _ in range(world)]
    dist.all_gather(gathered, img)
    rows = parallel.q1_rows_from_groups(parallel.merge_images_host([g.numpy() for g in gathered], CAP))
    cover = torch.tensor([r_hi - r_lo], dtype=torch.int64)
    dist.all_reduce(cover)
    if rank == 0:
        whole = datagen.lineitem(s, chunk_rows=1 << 20)
        want, _ = o.q1(o.table(whole))
        out.put((rows == want, int(cover.item()) == s.n_lineitem, len(rows)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_q1_sharding_protocol_gloo(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 1000 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, 0.02, 13, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    same, covered, n = out.get(timeout=10)
    assert covered, "order-range split must cover every lineitem row exactly once"
    assert same and n == 4


def test_image_roundtrip_and_wraparound():
    sys.path.insert(0, ROOT)
    from lingodb_b200 import parallel
    big = (1 << 127) + 12345
    img = parallel.pack_image(CAP, [((65, 70), [1, -2, big, -big, 5, 6, 0, 0])])
    assert len(img) == parallel.image_bytes(CAP)
    (key, aggs), = parallel.unpack_image(img, CAP)
    assert key == (65, 70) and aggs[0] == 1 and aggs[1] == (1 << 128) - 2
    merged = parallel.merge_images_host([img, img], CAP)
    assert merged[(65, 70)][2] == (2 * big) % (1 << 128)  # sums wrap mod 2^128 like the device atomics


def test_order_range_partition_is_exact():
    sys.path.insert(0, ROOT)
    from lingodb_b200 import datagen, parallel
    s = datagen.scale(0.013, 3)
    for world in (1, 2, 4, 8):
        prev = 0
        for r in range(world):
            o_lo, o_hi, r_lo, r_hi = parallel.order_range(s, r, world)
            assert r_lo == prev and r_hi >= r_lo
            prev = r_hi
        assert prev == s.n_lineitem


def test_q9_shaped_group_images_merge_on_the_host():
    """Two-key (nation, year) partial group tables of two ranks fold into the totals: the host twin of what
    parallel.q9_sharded does with all_gather + the K7 merge kernel."""
    from lingodb_b200 import parallel
    cap = 1024
    r0 = [((n, 1992 + y), [1000 * n + y]) for n in range(25) for y in range(7) if (n + y) % 3 != 0]
    r1 = [((n, 1992 + y), [-(7 * n) + (1 << 70) * y]) for n in range(25) for y in range(7) if (n + y) % 2 == 0]
    merged = dict(parallel.merge_images_host([parallel.pack_image(cap, r0), parallel.pack_image(cap, r1)], cap))
    want = {}
    for k, v in r0 + r1:
        want[k] = (want.get(k, 0) + v[0]) % (1 << 128)
    assert {k: v[0] % (1 << 128) for k, v in merged.items()} == want


def test_repartitioned_q5_heap_layout_is_a_pure_function_of_the_cardinalities():
    """ldb_tpch_q5_repartitioned_heap_bytes (csrc/tpch_plans.cpp q5Layout): every rank derives the same symmetric-heap layout from
    (orders, lineitem, world) alone — no device needed; the receive sub-regions shrink with world^2, the Bloom filter does not."""
    from lingodb_b200 import capi
    L = capi.lib()
    n_o, n_l = 150_000_000, 600_000_004
    sizes = {w: int(L.ldb_tpch_q5_repartitioned_heap_bytes(n_o, n_l, w)) for w in (1, 2, 4, 8)}
    assert all(v % 256 == 0 and v > 0 for v in sizes.values())
    assert sizes[1] > sizes[2] > sizes[4] > sizes[8] > 16 << 20  # the 16 MiB Bloom filter of the global orders key set stays
    per_rank_lineitem_region = lambda w: w * (n_l // 10 // (w * w) + 8192) * 24
    assert abs((sizes[1] - sizes[8]) - (per_rank_lineitem_region(1) - per_rank_lineitem_region(8))) < 64 << 20
    assert int(L.ldb_tpch_q5_repartitioned_heap_bytes(1500, 6000, 2)) < 4 << 20
