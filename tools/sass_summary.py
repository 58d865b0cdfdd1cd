#!/usr/bin/env python3
"""Static SASS summary of the shipped library (cuobjdump -sass): per kernel the instruction count and the mnemonics that
show how data moves and where the arithmetic runs -- bulk async copies (UBLKCP = cp.async.bulk, SYNCS = mbarrier), LDGSTS
(cp.async), 128-bit global accesses, FP64 (DFMA/DADD/DMUL/DSETP), MUFU, barriers, atomics, reductions.
usage: sass_summary.py [libfcz_engine.so]"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "foldcomp_b200", "csrc", "libfcz_engine.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
print(f"# {os.path.basename(lib)}: cubins for {', '.join(arch)}")
cols = ["total", "UBLKCP", "SYNCS", "LDGSTS", "LDG", "LDG.128", "STG", "STG.128", "LDS", "STS", "DFMA", "DADD", "DMUL", "DSETP", "MUFU", "FFMA", "FMUL", "FADD",
        "BAR", "ATOM", "RED", "REDUX", "SHFL", "CCTL", "BRA"]
print("kernel".ljust(34) + " ".join(c.rjust(8) for c in cols))
cur, cnt = None, None
def flush():
    if cur:
        print(cur[:33].ljust(34) + " ".join(str(cnt[c]).rjust(8) for c in cols))
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        flush()
        name = m.group(1)
        d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        cur, cnt = d.split("(")[0], collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        op = m.group(1)
        base = op.split(".")[0]
        cnt["total"] += 1
        cnt[base] += 1
        if base in ("LDG", "STG") and ".128" in op:
            cnt[base + ".128"] += 1
        if base in ("ATOMS", "ATOMG"):
            cnt["ATOM"] += 1
flush()
