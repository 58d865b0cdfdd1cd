#!/usr/bin/env python3
"""Sweep of the host-memory path's transfer parameters (FCZ_CHUNK_MB, FCZ_H2D_QUEUE, FCZ_D2H_QUEUE) and of the number of
sub-batches per step: end-to-end round-trip residues/s with encode and decode on two engines from two host threads.
One process per configuration would regenerate the batch each time, so the engines are re-created inside one process
(they read the variables at creation)."""
import itertools, json, os, sys, threading, time
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
def pin_in(b): return HostChainBatch(pin(b.res_off), pin(b.atom_off), pin(b.title_off), pin(b.res_type), pin(b.bfactor), pin(b.xyz), pin(b.titles), pin(b.meta), pin(np.zeros(b.n_chains, np.int32)))
def pin_out(b): return HostChainBatch(pin(np.zeros(b.n_chains + 1, np.uint32)), pin(np.zeros(b.n_chains + 1, np.uint64)), pin(np.zeros(b.n_chains + 1, np.uint32)), pin(np.zeros(b.n_res, np.uint8)), pin(np.zeros(b.n_res, np.float32)), pin(np.zeros((b.n_atoms, 3), np.float32)), pin(np.zeros(max(len(b.titles), 1), np.uint8)), pin(np.zeros(b.n_chains, abi.META_DTYPE)), pin(np.zeros(b.n_chains, np.int32)))
def pin_blob(b):
    c = abi.encode_bound(b.n_chains, b.n_res, b.n_atoms, len(b.titles), 25)
    return HostBlobBatch(pin(np.zeros(b.n_chains + 1, np.uint64)), pin(np.zeros(c, np.uint8)), pin(np.zeros(b.n_chains, np.int32)))
sets = {}
for P in (1, 2, 4):
    bounds = [round(i * batch.n_chains / P) for i in range(P + 1)]
    parts = [batch.select(range(bounds[i], bounds[i + 1])) for i in range(P)]
    NB = max(2, P)
    sets[P] = (parts, [pin_in(p) for p in parts], [pin_out(p) for p in parts], [pin_blob(max(parts, key=lambda q: q.n_atoms)) for _ in range(NB)])

def run(P, steps, e_enc, e_dec):
    parts, h_in, h_out, h_blob = sets[P]
    NB = len(h_blob)
    ready = [threading.Semaphore(0) for _ in range(NB)]; free = [threading.Semaphore(1) for _ in range(NB)]
    def view(slot, j):
        hb, nj = h_blob[slot], parts[j].n_chains
        return HostBlobBatch(hb.blob_off[: nj + 1], hb.bytes, hb.status[:nj])
    def enc():
        for i in range(steps * P):
            free[i % NB].acquire(); e_enc.encode_host(h_in[i % P], view(i % NB, i % P)); ready[i % NB].release()
    def dec():
        for i in range(steps * P):
            ready[i % NB].acquire(); e_dec.decode_host(view(i % NB, i % P), out=h_out[i % P]); free[i % NB].release()
    th = [threading.Thread(target=enc), threading.Thread(target=dec)]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    torch.cuda.synchronize()
    return time.perf_counter() - t0

res = []
for chunk, hq, dq in itertools.product((12, 32, 96, 400), (0, 2), (0, 2)):
    if hq != dq: continue
    os.environ["FCZ_CHUNK_MB"], os.environ["FCZ_H2D_QUEUE"], os.environ["FCZ_D2H_QUEUE"] = str(chunk), str(hq), str(dq)
    e1, e2 = Engine(0), Engine(0)
    for P in (1, 2, 4):
        run(P, 2, e1, e2)
        dt = run(P, 8, e1, e2)
        res.append({"chunk_mb": chunk, "queue": hq, "parts": P, "ms_per_step": round(1e3 * dt / 8, 3), "M_res_s": round(batch.n_res * 8 / dt / 1e6, 1)})
        print(res[-1], flush=True)
    e1.close(); e2.close()
best = max(res, key=lambda r: r["M_res_s"])
print(json.dumps({"best": best}))
