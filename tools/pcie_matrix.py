#!/usr/bin/env python3
"""Pinned-copy throughput with both directions busy, by copy size: 336 MB each way issued from two host threads on two
streams as copies of 1 / 4 / 12 / 48 / 336 MB.  Prints ms for each direction alone and together."""
import json, threading, time
import torch
dev = torch.device("cuda:0")
TOTAL = 336 << 20
h_a = torch.empty(TOTAL, dtype=torch.uint8).pin_memory(); h_b = torch.empty(TOTAL, dtype=torch.uint8).pin_memory()
d_a = torch.empty(TOTAL, dtype=torch.uint8, device=dev); d_b = torch.empty(TOTAL, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def h2d(chunk):
    with torch.cuda.stream(s1):
        for o in range(0, TOTAL, chunk): d_a[o:o + chunk].copy_(h_a[o:o + chunk], non_blocking=True)
    s1.synchronize()
def d2h(chunk):
    with torch.cuda.stream(s2):
        for o in range(0, TOTAL, chunk): h_b[o:o + chunk].copy_(d_b[o:o + chunk], non_blocking=True)
    s2.synchronize()
def timed(fn, *a):
    t0 = time.perf_counter(); fn(*a); return 1e3 * (time.perf_counter() - t0)
def pair(ca, cb):
    r = {}
    ta = threading.Thread(target=lambda: r.__setitem__("h2d", timed(h2d, ca))); tb = threading.Thread(target=lambda: r.__setitem__("d2h", timed(d2h, cb)))
    t0 = time.perf_counter(); ta.start(); tb.start(); ta.join(); tb.join()
    return round(r["h2d"], 2), round(r["d2h"], 2), round(1e3 * (time.perf_counter() - t0), 2)
out = {}
for mb in (1, 4, 12, 48, 336):
    c = mb << 20
    h2d(c); d2h(c)
    out[f"{mb}MB"] = {"h2d_alone": round(timed(h2d, c), 2), "d2h_alone": round(timed(d2h, c), 2), "both": pair(c, c), "both_again": pair(c, c)}
out["h2d 12MB | d2h 336MB"] = pair(12 << 20, TOTAL)
out["h2d 336MB | d2h 12MB"] = pair(TOTAL, 12 << 20)
print(json.dumps(out, indent=1))
