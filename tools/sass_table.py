#!/usr/bin/env python
"""Instruction-mix table of the kernels in libldb_gpu.so (profiles/*_sass_excerpts.md): `cuobjdump -sass` per function, counted by mnemonic."""
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "lingo-db_b200/libldb_gpu.so"
want = sys.argv[2].split(";") if len(sys.argv) > 2 else None
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs, cur = {}, None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append(m.group(1).strip())
names = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
print("| kernel | SASS instr | UBLKCP (TMA bulk) | SYNCS (mbarrier) | ATOMS (smem atomics) | ATOMG/RED (global atomics) | MATCH/VOTE/SHFL | LDG | STG/ST | BAR.SYNC | LDCU+BRA.U (uniform control) | tensor (HMMA/UTC*) |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
tot = {"n": 0, "instr": 0}
for (mangled, ins), name in zip(funcs.items(), names):
    tot["n"] += 1
    tot["instr"] += len(ins)
    short = re.sub(r"\(.*$", "", name).replace("(int)", "")
    if want and not any(w in short for w in want):
        continue
    def c(*pre):
        return sum(1 for i in ins if any(re.sub(r"^@!?U?P\d\s+", "", i).startswith(p) for p in pre))
    print(f"| `{short[:90]}` | {len(ins)} | {c('UBLKCP')} | {c('SYNCS')} | {c('ATOMS')} | {c('ATOMG', 'RED')} | {c('MATCH')}/{c('VOTE')}/{c('SHFL')} | {c('LDG')} | {c('STG', 'ST.')} | {c('BAR.SYNC')} | {c('LDCU', 'BRA.U')} | {c('HMMA', 'UTC', 'IMMA', 'QMMA')} |")
print(f"\nWhole library: {tot['n']} kernels, {tot['instr']} SASS instructions.")
