#!/usr/bin/env python3
"""Static SASS instruction counts of one kernel, attributed to the OUTERMOST source line inside a file/line range
(following nvdisasm -gi inline chains), optionally grouped into named buckets.  No GPU needed: a per-iteration size
check of the hot loops before spending GPU time.
usage: sass_static.py <cubin|.so> <mangled kernel> <file> <first> <last> [name:lo-hi,...]"""
import collections, re, subprocess, sys, tempfile, os

obj, kern, fname, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
buckets = []
if len(sys.argv) > 6:
    for b in sys.argv[6].split(","):
        name, rng = b.split(":"); a, z = rng.split("-"); buckets.append((name, int(a), int(z)))
cubin = obj
if obj.endswith(".so"):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
    cubin = os.path.join(d, sorted(os.listdir(d))[0])
dis = subprocess.run(["nvdisasm", "-g", "-gi", cubin], capture_output=True, text=True).stdout
m0 = re.search(r"^\s*\.section\s+\.text\." + re.escape(kern) + r"\b.*$", dis, re.M)
sec = dis[m0.end():]
nxt = re.search(r"^\s*\.section\s", sec, re.M)
sec = sec[: nxt.start() if nxt else None]
cnt = collections.Counter(); ops = collections.defaultdict(collections.Counter)
chain, in_block = [], False
for ln in sec.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if not in_block:
            chain, in_block = [], True
        chain.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?(\S+)", ln)
    if m:
        in_block = False
        pick = None
        for f, l in chain:
            if f == fname and lo <= l <= hi:
                pick = l
        name = "other"
        if pick is not None:
            name = f"{fname}:{pick}"
            for bn, a, z in buckets:
                if a <= pick <= z:
                    name = bn
        cnt[name] += 1
        ops[name][m.group(2).split(".")[0]] += 1
tot = sum(cnt.values())
print(f"{kern}: {tot} SASS instructions")
for name, n in sorted(cnt.items(), key=lambda kv: -kv[1])[: int(os.environ.get("TOP", "30"))]:
    top = ", ".join(f"{o} {c}" for o, c in ops[name].most_common(8))
    print(f"{n:6d}  {name:24s} {top}")
