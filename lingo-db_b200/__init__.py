"""lingo-db_b200 — Blackwell-native backend for LingoDB's three data-parallel hot paths
(Arrow scan + predicates, hash-join build/probe, hash group-by), behind a C-ABI (include/ldb_gpu.h).

The directory name carries a hyphen (fixed by the task layout); import it as `lingodb_b200`
(the sibling alias package) — both names resolve to this directory.
"""
from . import build  # noqa: F401

__all__ = ["build", "datagen", "capi", "runtime", "parallel", "devgen", "dbgen", "arrow_io"]
