#!/usr/bin/env python3
"""Attribute an ncu SASS source page (per-instruction executed counts / stall samples) to the OUTERMOST source line of a
given file and line range (e.g. the body of encode_chain in fcz_codec.h), following nvdisasm -gi's inline chains, so that
shared helpers (cross3, ld3, acos ...) are charged to the phase that called them.
usage: ncu_phases.py <report.ncu-rep> <mangled kernel> <cubin> <file> <first line> <last line> [bucket,bucket,...]
  buckets: "name:lo-hi" ranges over the lines of <file>; without them every line is its own bucket."""
import collections, csv, os, re, subprocess, sys

rep, kern, cubin, fname, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
buckets = []
if len(sys.argv) > 7:
    for b in sys.argv[7].split(","):
        name, rng = b.split(":")
        a, z = rng.split("-")
        buckets.append((name, int(a), int(z)))
dis = subprocess.run(["nvdisasm", "-g", "-gi", cubin], capture_output=True, text=True).stdout
m0 = re.search(r"^\s*\.section\s+\.text\." + re.escape(kern) + r"\b.*$", dis, re.M)
sec = dis[m0.end():]
nxt = re.search(r"^\s*\.section\s", sec, re.M)
sec = sec[: nxt.start() if nxt else None]
key_of, leaf_of = {}, {}
chain, in_block = [], False  # an annotation block (consecutive //## lines, innermost frame first) holds until the next block
for ln in sec.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if not in_block:
            chain, in_block = [], True
        chain.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m:
        in_block = False
        if not chain:
            continue
        pick = None
        for f, l in chain:  # innermost -> outermost; keep the outermost frame inside the range
            if f == fname and lo <= l <= hi:
                pick = (f, l)
        addr = int(m.group(1), 16)
        key_of[addr] = pick if pick else ("other:" + chain[-1][0], chain[-1][1])
        leaf_of[addr] = chain[0]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + os.environ.get("KNAME", ".*")], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi_ = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi_]
si, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
base = None
smp, ins = collections.Counter(), collections.Counter()
for r in rows[hi_ + 1:]:
    try:
        addr = int(r[0], 16)
    except ValueError:
        continue
    if base is None:
        base = addr
    f, l = key_of.get(addr - base, ("?", 0))
    name = None
    if f == fname:
        for bn, a, z in buckets:
            if a <= l <= z:
                name = bn
        if name is None:
            name = f"{f}:{l}"
    else:
        name = f if buckets else f"{f}:{l}"
    if os.environ.get("LEAF"):  # split every bucket by the innermost (leaf) source line
        lf = leaf_of.get(addr - base, ("?", 0))
        name = f"{name:14s} {lf[0]}:{lf[1]}"
    smp[name] += int(r[si]); ins[name] += int(r[ie])
ts, ti = sum(smp.values()), sum(ins.values())
print(f"total samples {ts}, warp instructions {ti}")
for name, n in sorted(ins.items(), key=lambda kv: -kv[1])[: int(os.environ.get("TOP", "40"))]:
    print(f"{100*n/ti:5.1f}% inst {100*smp[name]/ts:5.1f}% samples  {name}")
