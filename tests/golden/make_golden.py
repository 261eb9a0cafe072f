#!/usr/bin/env python
"""Extract the reference's own known-answer vectors for the hot path into reference_kats.json.

Run in the build container (needs /root/reference); the JSON is committed so the tests run anywhere.
Sources:
  test/lit/DB/hash.mlir                      CHECK lines 27-34 (db.hash of constants, executed through the JIT)
  test/unittests/storage/TestStorage.cpp     dbHash64(int8 1) literal (:289, :411)
  test/lit/DB/dates.mlir                     ExtractFromDate(year, 2020-06-13) = 2020 (:23-26)
  test/sqlite-datasets/tpchSf1.test          Q9 answer keys (nation, o_year): 25 nations x 1992..1998 (:20632-20807)
  test/sqlite-datasets/tpchSf1.test          Q1 answer rows (:25-28): avg = (sum * 10^19) / count, Q6 (:20506)
"""
import json
import os
import re

REF = "/root/reference"
out = {"source_commit": "fab813e8", "hash": {}, "tpch_sf1": {}}

mlir = open(os.path.join(REF, "test/lit/DB/hash.mlir")).read()
checks = [int(x) for x in re.findall(r"//CHECK: index\((\d+)\)", mlir)]
names = ["i32_10", "i64_10", "bool_true", "decimal15_2_100.01", "date_2020-06-11", "timestamp_s_2020-06-11_12:30:00", "string_hello_world!", "tuple7"]
assert len(checks) == len(names)
out["hash"] = dict(zip(names, checks))
out["hash"]["file"] = "test/lit/DB/hash.mlir:27-34"

ts = open(os.path.join(REF, "test/unittests/storage/TestStorage.cpp")).read()
m = re.findall(r"(-3\d{18})", ts)  # the literal next to dbHash64 / hash(1) checks
out["hash"]["int8_1"] = int(m[0])
out["hash"]["int8_1_file"] = "test/unittests/storage/TestStorage.cpp:289,411"

lines = open(os.path.join(REF, "test/sqlite-datasets/tpchSf1.test")).read().split("\n")
q1 = []
for ln in lines[24:28]:
    f = ln.split("\t")
    q1.append({"l_returnflag": f[0], "l_linestatus": f[1], "sum_qty": f[2], "sum_base_price": f[3], "sum_disc_price": f[4], "sum_charge": f[5],
               "avg_qty": f[6], "avg_price": f[7], "avg_disc": f[8], "count_order": f[9]})
out["tpch_sf1"]["q1"] = q1
out["tpch_sf1"]["q1_file"] = "test/sqlite-datasets/tpchSf1.test:25-28"
out["tpch_sf1"]["q6"] = lines[20505].strip()
out["tpch_sf1"]["q6_file"] = "test/sqlite-datasets/tpchSf1.test:20506"
dates = open(os.path.join(REF, "test/lit/DB/dates.mlir")).read()
m = re.search(r'db.constant \( "(\d{4}-\d{2}-\d{2})"\) : !db.date<day>', dates)
years = re.findall(r"//CHECK: int\((\d{4})\)", dates)
out["dates"] = {"date": m.group(1), "extract_year": int(years[0]), "file": "test/lit/DB/dates.mlir:6,23-26"}
q9 = [ln.split("\t")[:2] for ln in lines[20631:20807] if ln.count("\t") == 2]
out["tpch_sf1"]["q9_keys"] = [[n.strip(), int(y)] for n, y in q9]
out["tpch_sf1"]["q9_file"] = "test/sqlite-datasets/tpchSf1.test:20632-20807"
def answer_rows(tag):
    """rows between the '----' after `query … tag` and the next blank line"""
    i = next(k for k, ln in enumerate(lines) if ln.startswith("query") and ln.rstrip().endswith(tag))
    j = next(k for k in range(i, len(lines)) if lines[k].strip() == "----")
    rows = []
    for ln in lines[j + 1:]:
        if not ln.strip():
            break
        rows.append([f.strip() for f in ln.split("\t")])
    return rows, f"test/sqlite-datasets/tpchSf1.test:{j + 2}-{j + 1 + len(rows)}"


for q in ("q3", "q5", "q9", "q4", "q12", "q18", "q7", "q21", "q10", "q15", "q14", "q17", "q19", "q2", "q8", "q11", "q20", "q22"):
    out["tpch_sf1"][q + "_rows"], out["tpch_sf1"][q + "_rows_file"] = answer_rows("tpch" + q)
# Q16 has 18 314 answer rows: keep their count, a digest of the tab-joined rows and the two ends instead of 900 KB of text
import hashlib
q16, q16_file = answer_rows("tpchq16")
out["tpch_sf1"]["q16"] = {"rows": len(q16), "sha256": hashlib.sha256("\n".join("\t".join(r) for r in q16).encode()).hexdigest(), "first": q16[:3], "last": q16[-3:], "file": q16_file}
out["tpch_sf1"]["q2_rows"] = [r[:5] for r in out["tpch_sf1"]["q2_rows"]]  # s_acctbal, s_name, n_name, p_partkey, p_mfgr
out["tpch_sf1"]["q20_rows"] = [r[:1] for r in out["tpch_sf1"]["q20_rows"]]  # s_name
out["tpch_sf1"]["q10_rows"] = [r[:5] for r in out["tpch_sf1"]["q10_rows"]]  # c_custkey, c_name, revenue, c_acctbal, n_name (address / phone / comment are not generated here)
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json"), "w"), indent=1)
print(json.dumps(out, indent=1)[:1500])
