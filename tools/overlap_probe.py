#!/usr/bin/env python3
"""Which resource makes a host-memory decode crawl next to a host-memory encode?  Times encode_host / decode_host alone,
next to each other, and next to plain pinned copies in the opposite direction."""
import json, os, sys, threading, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from foldcomp_b200 import abi, synth
from foldcomp_b200.abi import HostBlobBatch, HostChainBatch
from foldcomp_b200.engine import Engine

dev = torch.device("cuda:0")
batch = synth.generate(10000, 350, seed=synth.SEED)
keep = []
def pin(a):
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).pin_memory(); keep.append(t)
    return t.numpy().view(a.dtype).reshape(a.shape)
hb = HostChainBatch(pin(batch.res_off), pin(batch.atom_off), pin(batch.title_off), pin(batch.res_type), pin(batch.bfactor), pin(batch.xyz), pin(batch.titles), pin(batch.meta), pin(np.zeros(batch.n_chains, np.int32)))
cap = abi.encode_bound(batch.n_chains, batch.n_res, batch.n_atoms, len(batch.titles), 25)
hblob = HostBlobBatch(pin(np.zeros(batch.n_chains + 1, np.uint64)), pin(np.zeros(cap, np.uint8)), pin(np.zeros(batch.n_chains, np.int32)))
hblob2 = HostBlobBatch(pin(np.zeros(batch.n_chains + 1, np.uint64)), pin(np.zeros(cap, np.uint8)), pin(np.zeros(batch.n_chains, np.int32)))
nt = len(batch.titles)
hout = HostChainBatch(pin(np.zeros(batch.n_chains + 1, np.uint32)), pin(np.zeros(batch.n_chains + 1, np.uint64)), pin(np.zeros(batch.n_chains + 1, np.uint32)), pin(np.zeros(batch.n_res, np.uint8)), pin(np.zeros(batch.n_res, np.float32)), pin(np.zeros((batch.n_atoms, 3), np.float32)), pin(np.zeros(nt, np.uint8)), pin(np.zeros(batch.n_chains, abi.META_DTYPE)), pin(np.zeros(batch.n_chains, np.int32)))
e1, e2 = Engine(0), Engine(0)
e1.encode_host(hb, hblob); e1.encode_host(hb, hblob2); e2.decode_host(hblob, out=hout)
chunk = 12 << 20
hbuf = torch.empty(28 * chunk, dtype=torch.uint8).pin_memory()
dbuf = torch.empty(28 * chunk, dtype=torch.uint8, device=dev)
st = torch.cuda.Stream()
def plain(direction):
    with torch.cuda.stream(st):
        for k in range(28):
            sl = slice(k * chunk, (k + 1) * chunk)
            if direction == "h2d": dbuf[sl].copy_(hbuf[sl], non_blocking=True)
            else: hbuf[sl].copy_(dbuf[sl], non_blocking=True)
    st.synchronize()
def timed(fn):
    t0 = time.perf_counter(); fn(); return 1e3 * (time.perf_counter() - t0)
def pair(fa, fb):
    res = {}
    def run(name, fn): res[name] = timed(fn)
    ta, tb = threading.Thread(target=run, args=("a", fa)), threading.Thread(target=run, args=("b", fb))
    t0 = time.perf_counter(); ta.start(); tb.start(); ta.join(); tb.join()
    return round(res["a"], 2), round(res["b"], 2), round(1e3 * (time.perf_counter() - t0), 2)
enc = lambda: e1.encode_host(hb, hblob2)
dec = lambda: e2.decode_host(hblob, out=hout)
out = {}
if os.environ.get("FCZ_TRACE_HOST"):
    print("---- dec alone", file=sys.stderr); dec()
    print("---- dec next to enc", file=sys.stderr); print(pair(enc, dec), file=sys.stderr)
    print("---- dec next to plain h2d", file=sys.stderr); print(pair(dec, lambda: plain("h2d")), file=sys.stderr)
    sys.exit(0)
for _ in range(2):
    out["enc_alone"] = round(timed(enc), 2); out["dec_alone"] = round(timed(dec), 2)
    out["h2d_alone"] = round(timed(lambda: plain("h2d")), 2); out["d2h_alone"] = round(timed(lambda: plain("d2h")), 2)
    out["enc|d2h"] = pair(enc, lambda: plain("d2h")); out["dec|h2d"] = pair(dec, lambda: plain("h2d"))
    out["h2d|d2h"] = None
    out["enc|dec"] = pair(enc, dec)
print(json.dumps(out))
