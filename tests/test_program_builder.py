"""Host logic of the program pipeline (lingodb_b200/program.py): expression trees → LdbInstr lists, without a GPU."""
import ctypes as C

import pytest

from lingodb_b200 import program as P

col, const = (lambda n: ("col", n)), (lambda v: ("const", v))


def test_common_subexpressions_columns_and_constants_are_emitted_once():
    b = P.Builder()
    disc_price = ("mul", col("l_extendedprice"), ("sub", const(100), col("l_discount")))
    r1 = b.expr(disc_price)
    r2 = b.expr(("mul", disc_price, ("add", const(100), col("l_tax"))))
    assert b.expr(disc_price) == r1 and r2 != r1
    ops = [i[0] for i in b.instr]
    assert ops.count(P.OPS["load"]) == 3 and ops.count(P.OPS["const"]) == 1 and ops.count(P.OPS["mul"]) == 2
    assert b.columns == ["l_extendedprice", "l_discount", "l_tax"] and b.consts == [100]
    # every operand register is written by an earlier instruction (what ldb_gpu_run_program validates again on the C++ side)
    written = set()
    for op, dst, a, bb, arg in b.instr:
        if op in (P.OPS["add"], P.OPS["sub"], P.OPS["mul"]):
            assert a in written and bb in written
        written.add(dst)


def test_between_case_strings_and_negative_constants():
    b = P.Builder()
    b.expr(("between", col("d"), const(5), const(7)))
    ops = [i[0] for i in b.instr]
    assert ops.count(P.OPS["cmp"]) == 2 and ops.count(P.OPS["and"]) == 1
    cmps = [i for i in b.instr if i[0] == P.OPS["cmp"]]
    assert [c[4] for c in cmps] == [P.CMP[">="], P.CMP["<="]]
    r = b.expr(("case", ("strcmp", "=", "s", "MAIL"), const(-1), const(0)))
    sel = b.instr[-1]
    assert sel[0] == P.OPS["select"] and sel[1] == r
    assert b.strings == ["MAIL"] and (1 << 128) - 1 in b.consts  # -1 as a 128-bit two's complement constant
    like = b.expr(("like", "prefix", "s", "PROMO"))
    assert b.instr[-1][:1] == (P.OPS["strlike"],) and b.instr[-1][3] == P.LIKE["prefix"] and like == b.instr[-1][1]


def test_probe_tables_are_listed_once_and_probes_nest():
    b = P.Builder()
    orders, cust = C.c_void_p(1), C.c_void_p(2)
    cn = ("probe", cust, ("probe", orders, col("l_orderkey")))
    b.expr(("cmp", "=", cn, const(7)))
    b.expr(cn)
    assert len(b.tables) == 2 and b.tables[0] is cust and b.tables[1] is orders  # listed once each, outer probe first
    probes = [i for i in b.instr if i[0] == P.OPS["probe"]]
    assert [p[4] for p in probes] == [1, 0, 1, 0]  # inner (orders) before outer (customer); probes are re-evaluated, the column load is not
    assert [i[0] for i in b.instr].count(P.OPS["load"]) == 1


def test_register_budget_is_enforced():
    b = P.Builder()
    e = col("a")
    for k in range(60):
        e = ("add", e, const(k))
    with pytest.raises(ValueError, match="48 registers"):
        b.expr(e)
    with pytest.raises(ValueError, match="unknown expression"):
        P.Builder().expr(("frobnicate", col("a")))
