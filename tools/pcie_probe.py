#!/usr/bin/env python3
"""Pinned-memory PCIe bandwidth of the box: H2D alone, D2H alone, both directions at once (two streams).
The end-to-end figure of bench.py is bounded by these (a round trip moves ~115 B/residue each way)."""
import json
import torch

n = 512 << 20
h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_a, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_b.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


out = {"bytes": n, "h2d_gbs": n / timed(h2d) / 1e6, "d2h_gbs": n / timed(d2h) / 1e6}
t = timed(both)
out["bidir_each_way_gbs"] = n / t / 1e6
print(json.dumps(out))
