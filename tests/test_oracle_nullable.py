"""Nullable scans on the CPU side: the oracle's pushed-down filters over columns with validity bitmaps (the reference's own
Restrictions when oracle/_ref is built, Restrictions.cpp:67-162,392-405) against a numpy evaluation, and the port against the
reference-compiled build — this is what pins the NULL semantics the GPU program pipeline is then compared with."""
import numpy as np
import pytest

from oracle import oracle as O

from _nullable import cases, nullable_lineitem


def _np_eval(li, valid, filters, sum_column, date):
    cat = lambda k: np.concatenate([c[k] for c in li.chunks])
    lo64 = lambda a: a[:, :8].copy().view(np.int64).reshape(-1)
    vals = {}
    for k in valid:
        a = cat(k)
        vals[k] = lo64(a) if a.ndim == 2 else a.astype(np.int64)
    keep = np.ones(li.num_rows, bool)
    for c, op, v in filters:
        if op == "notnull":
            keep &= valid[c]
            continue
        if isinstance(v, str):
            v = date(v) if "-" in v else int(round(float(v) * 100))
        elif li.spec(c).phys == "decimal128":
            v *= 100
        keep &= {"<": vals[c] < v, "<=": vals[c] <= v, ">": vals[c] > v, ">=": vals[c] >= v, "=": vals[c] == v, "!=": vals[c] != v}[op]
    total = int(vals[sum_column][keep & valid[sum_column]].sum()) if sum_column else 0
    return int(keep.sum()), total


@pytest.mark.parametrize("kind", ["port", "reference"])
def test_pushed_down_filters_over_nullable_columns(kind):
    try:
        o = O.Oracle(kind, workers=3)
    except OSError:
        pytest.skip(f"oracle build '{kind}' not present")
    li, valid = nullable_lineitem()
    h = o.table(li)
    date = lambda s: o.lib.oracle_parse_date(s.encode())
    for filters, _, sum_column in cases(date):
        assert o.scan_count_sum(h, filters, sum_column) == _np_eval(li, valid, filters, sum_column, date), filters
    # a comparison WITHOUT its NOTNULL partner reads the cell under a NULL like any other (the reference relies on the optimizer to add it)
    got = o.scan_count_sum(h, [("l_suppkey", ">=", 0)], None)
    assert got[0] == li.num_rows
    o.free(h)
