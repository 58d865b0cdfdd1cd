#!/usr/bin/env python3
"""Join an ncu SASS source page (per-instruction samples) with nvdisasm line info and print the
hottest source lines.  usage: ncu_lines.py <report.ncu-rep> <mangled kernel> <cubin> [topN]"""
import csv, re, subprocess, sys, collections
rep, kern, cubin = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
m0 = re.search(r"^\s*\.section\s+\.text\." + re.escape(kern) + r"\b.*$", dis, re.M)
sec = dis[m0.end():]
nxt = re.search(r"^\s*\.section\s", sec, re.M)
sec = sec[: nxt.start() if nxt else None]
line_of = {}
cur = None
inl = None
for ln in sec.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + __import__("os").environ.get("KNAME", ".*")], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
si, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
base = None
agg = collections.Counter(); inst = collections.Counter()
for r in rows[hi + 1:]:
    try:
        addr = int(r[0], 16)
    except ValueError:
        continue
    if base is None:
        base = addr
    key = line_of.get(addr - base, ("?", 0))
    agg[key] += int(r[si]); inst[key] += int(r[ie])
tot = sum(agg.values()); ti = sum(inst.values())
print(f"total samples {tot}, warp instructions {ti}")
srcs = {}
for (f, l), n in agg.most_common(top):
    if f not in srcs:
        try:
            srcs[f] = open(__import__("os").environ.get("SRCDIR", "/root/repo/foldcomp_b200/csrc/") + f).read().splitlines()
        except OSError:
            srcs[f] = []
    text = srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
    print(f"{100*n/tot:5.1f}% samples {100*inst[(f,l)]/ti:5.1f}% inst  {f}:{l}  {text}")
