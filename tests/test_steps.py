"""Serialised execution steps (csrc/step_json.cpp): structure validation without a device; the five TPC-H plans of
tests/golden/plans/ run through ldb_gpu_run_step[_hex] on the GPU and compared with the oracle."""
import ctypes as C
import glob
import json
import os
import re

import pytest

from lingodb_b200 import capi, datagen

PLANS = os.path.join(os.path.dirname(__file__), "golden", "plans")


def _steps(name, sizes):
    plan = json.load(open(os.path.join(PLANS, name + ".json")))

    def resolve(v):
        if isinstance(v, str) and v.startswith("$n_"):
            m = re.fullmatch(r"\$n_(\w+?)(?:/(\d+))?", v)
            return sizes[m.group(1)] // int(m.group(2) or 1) + 1024
        if isinstance(v, dict):
            return {k: resolve(x) for k, x in v.items()}
        if isinstance(v, list):
            return [resolve(x) for x in v]
        return v
    return [resolve(s) for s in plan["steps"]]


def test_every_golden_plan_validates_and_bad_documents_are_refused():
    L = capi.lib()
    sizes = {"customer": 1500, "orders": 15000, "supplier": 100, "part": 2000, "partsupp": 8000}
    for path in sorted(glob.glob(os.path.join(PLANS, "*.json"))):
        for step in _steps(os.path.basename(path)[:-5], sizes):
            e = capi.Error()
            assert L.ldb_gpu_step_validate(json.dumps(step).encode(), C.byref(e)) == capi.LDB_OK, (path, e.message)
    for bad, code in (('{"kind": "scan_groupby"}', capi.LDB_ERR_INVALID), ('{"kind": "warp_drive", "source": "t", "sink": {"name": "s"}}', capi.LDB_ERR_UNSUPPORTED),
                      ('{"kind": "scan_reduce", "source": "t", "filters": [{"column": "a", "op": "~", "value": 1}], "sink": {"name": "s"}}', capi.LDB_ERR_UNSUPPORTED),
                      ('{"kind": "scan_reduce", "source": "t", "filters": [{"column": "a", "op": "<", "value": 1.5}], "sink": {"name": "s"}}', capi.LDB_ERR_INVALID),
                      ('{"kind": "scan_reduce", "source": "t"', capi.LDB_ERR_INVALID), ("[1, 2]", capi.LDB_ERR_INVALID)):
        e = capi.Error()
        assert L.ldb_gpu_step_validate(bad.encode(), C.byref(e)) == code, (bad, e.message)


@pytest.mark.gpu
def test_golden_plans_run_through_the_step_interface(oracle):
    """Every step of the five plans goes through ldb_gpu_run_step (Q3's also hex-encoded); the states the steps created are read
    with the ordinary C-ABI calls and must equal the oracle's rows."""
    from lingodb_b200 import runtime
    from lingodb_b200.runtime import Tpch
    gpu_ctx = runtime.Context(0)  # its own context: steps resolve tables and states by NAME
    t = datagen.tpch(0.05, seed=19, chunk_rows=20_000, with_parts=True)
    oh = {k: oracle.table(v) for k, v in t.items()}
    tabs = {k: gpu_ctx.table_from_host(v) for k, v in t.items()}
    sizes = {k: v.num_rows for k, v in t.items()}
    L = gpu_ctx.L

    def run(name, hex_=False):
        for step in _steps(name, sizes):
            e = capi.Error()
            doc = json.dumps(step).encode()
            rc = L.ldb_gpu_run_step_hex(gpu_ctx.h, doc.hex().encode(), C.byref(e)) if hex_ else L.ldb_gpu_run_step(gpu_ctx.h, doc, C.byref(e))
            capi.check(rc, e)

    def state(name):
        s = L.ldb_gpu_find_state(gpu_ctx.h, name.encode())
        assert s, name
        return C.c_void_p(s)

    tp = Tpch(gpu_ctx, tabs)
    run("q6")
    out, e = (capi.I128 * 8)(), capi.Error()
    capi.check(L.ldb_gpu_simple_state_read(state("q6_revenue"), out, C.byref(e)), e)
    assert {"revenue": out[0].value()} == oracle.q6(oh["lineitem"])[0]
    run("q1")
    assert tp.q1_finish(state("q1_groups")) == oracle.q1(oh["lineitem"])[0]
    run("q3", hex_=True)
    rows, n = (capi.TopKRow * 10)(), C.c_int32()
    capi.check(L.ldb_gpu_join_table_topk(state("q3_map"), 10, rows, C.byref(n), C.byref(e)), e)
    got3 = [{"l_orderkey": r.key, "revenue": r.agg.value(), "o_orderdate": r.side[0], "o_shippriority": r.side[1]} for r in rows[: n.value]]
    assert got3 == oracle.q3(oh["customer"], oh["orders"], oh["lineitem"])[0]
    run("q5")
    grows, gn = runtime.groupby_read(gpu_ctx, state("q5_groups"))
    names = [nm for nm, _ in datagen.NATIONS]
    got5 = sorted(({"n_name": names[grows[i].keys[0]], "revenue": grows[i].aggs[0].value()} for i in range(gn)), key=lambda r: (-r["revenue"], r["n_name"]))
    assert got5 == oracle.q5(oh["customer"], oh["orders"], oh["lineitem"], oh["supplier"], oh["nation"], oh["region"])[0]
    run("q9")
    assert tp.q9_finish(state("q9_groups")) == oracle.q9(oh["part"], oh["supplier"], oh["lineitem"], oh["partsupp"], oh["orders"], oh["nation"])[0]
    for nm in ("q6_revenue", "q1_groups", "q3_cust", "q3_map", "q5_region", "q5_nation", "q5_cust", "q5_ord", "q5_supp", "q5_groups", "q9_part", "q9_ps", "q9_supp", "q9_ord", "q9_groups"):
        L.ldb_gpu_state_destroy(state(nm))
        assert not L.ldb_gpu_find_state(gpu_ctx.h, nm.encode())
    gpu_ctx.close()
