"""Shared by the nullable-scan parity tests: a generator lineitem with validity bitmaps injected, and the filter sets the
reference's pushdown produces for nullable columns (Pushdown.cpp:346-352: [NOTNULL, cmp]; :365-372: [NOTNULL, lower, upper])."""
import numpy as np

from lingodb_b200 import datagen

NULL_FRACTION = {"l_quantity": 0.07, "l_extendedprice": 0.11, "l_discount": 0.03, "l_shipdate": 0.05, "l_suppkey": 0.02}


def nullable_lineitem(sf=0.02, seed=91, chunk_rows=30_011):
    """lineitem batches (ragged against the 20 000-row morsels) whose listed columns carry Arrow validity bitmaps; returns the table
    and, per column, the concatenated boolean validity for the numpy cross-check."""
    s = datagen.scale(sf, seed=seed)
    li = datagen.lineitem(s, ["l_suppkey", "l_quantity", "l_extendedprice", "l_discount", "l_shipdate"], chunk_rows=chunk_rows)
    rng = np.random.default_rng(seed)
    valid = {k: [] for k in NULL_FRACTION}
    for chunk, n in zip(li.chunks, li.chunk_rows):
        for k, p in NULL_FRACTION.items():
            v = rng.random(n) > p
            valid[k].append(v)
            chunk[k + "$valid"] = np.packbits(v, bitorder="little")
    return li, {k: np.concatenate(v) for k, v in valid.items()}


# (pushed-down filters as the reference writes them, SQL predicate as a program expression, summed column)
def cases(date):
    col = lambda n: ("col", n)
    const = lambda v: ("const", v)
    return [
        ([("l_quantity", "notnull", None), ("l_quantity", "<", 24)], ("cmp", "<", col("l_quantity"), const(2400)), "l_extendedprice"),
        ([("l_shipdate", "notnull", None), ("l_shipdate", ">=", "1994-01-01"), ("l_discount", "notnull", None), ("l_discount", ">=", "0.05"), ("l_discount", "<=", "0.07")],
         ("and", ("cmp", ">=", col("l_shipdate"), const(date("1994-01-01"))), ("between", col("l_discount"), const(5), const(7))), "l_quantity"),
        ([("l_suppkey", "notnull", None), ("l_suppkey", "!=", 7), ("l_extendedprice", "notnull", None)],
         ("and", ("cmp", "!=", col("l_suppkey"), const(7)), ("not", ("isnull", col("l_extendedprice")))), "l_suppkey"),
        ([("l_extendedprice", "notnull", None)], ("not", ("isnull", col("l_extendedprice"))), None),
        ([], None, "l_discount"),
    ]
