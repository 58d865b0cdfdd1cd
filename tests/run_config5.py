#!/usr/bin/env python3
"""BASELINE.json configs[4]: mixed-length batch (50-2000 residues, clipped log-normal AFDB proxy), -b anchor sweep
10 / 25 / 50 / 200.  Per threshold: device-resident round-trip residues/s (CUDA events), FCZ bytes per residue, backbone
and all-atom RMSD of the round trip against the ORIGINAL coordinates for the engine and -- on a sample -- for the oracle
(the reference's own loss at that threshold), and the engine-vs-oracle decode deviation on that sample."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402  (the checker; this script is a measurement tool, not product code)
from foldcomp_b200 import abi, synth  # noqa: E402
from foldcomp_b200.abi import HostChainBatch  # noqa: E402
from foldcomp_b200.engine import DeviceBlobBatch, DeviceChainBatch, Engine  # noqa: E402

n = int(os.environ.get("N_CHAINS", "10000"))
steps = int(os.environ.get("STEPS", "10"))
dev = torch.device("cuda:0")
rng = np.random.default_rng(synth.SEED)
lens = synth.mixed_lengths(rng, n)
batch = synth.generate(n, lens, seed=synth.SEED)
sample = list(range(0, n, max(n // 100, 1)))
sub = batch.select(sample)
eng = Engine(0)
dbatch = DeviceChainBatch.from_host(batch, dev)


def pooled(a, b):
    """(backbone RMSD, all-atom RMSD) pooled over every atom of two batches with identical layout."""
    d2 = ((a.xyz.astype(np.float64) - b.xyz.astype(np.float64)) ** 2).sum(axis=1)
    m = H.backbone_mask_fast(a.res_type)
    return float(np.sqrt(d2[m].mean())), float(np.sqrt(d2.mean()))


rows = []
for b in (10, 25, 50, 200):
    eng.set_opts(anchor_threshold=b)
    cap = abi.encode_bound(n, batch.n_res, batch.n_atoms, len(batch.titles), b)
    dblob = DeviceBlobBatch(n, cap, dev)
    dout = DeviceChainBatch(n, batch.n_res, batch.n_atoms, len(batch.titles), dev)
    torch.cuda.synchronize()

    def step():
        eng.encode_device(dbatch, dblob)
        eng.decode_plan_device(dblob, dout)
        eng.decode_device(dblob, dout)

    for _ in range(3):
        step()
    eng.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    torch.cuda.synchronize()
    for _ in range(steps):
        step()
    eng.sync()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    got = dout.to_host()
    assert not got.status.any()
    orig = HostChainBatch(batch.res_off, batch.atom_off, batch.title_off, batch.res_type, got.bfactor, batch.xyz, batch.titles, batch.meta)
    bb, allr = pooled(got, orig)
    wbb, _, _ = H.per_chain_deviation(got, orig)
    # the oracle on the sample: its own round-trip loss, and how far the engine's decode is from it
    ob = H.oracle_encode_batch(sub, b)
    od = H.oracle_decode_batch(ob)
    so = HostChainBatch(sub.res_off, sub.atom_off, sub.title_off, sub.res_type, od.bfactor, sub.xyz, sub.titles, sub.meta)
    obb, oall = pooled(od, so)
    gs = got.select(sample)
    dbb, dall, dmax = H.per_chain_deviation(gs, od)
    rows.append({"b": b, "ms_per_round_trip": ms, "residues_per_s": batch.n_res / (ms * 1e-3), "fcz_bytes_per_residue": int(dblob.blob_off[-1].item()) / batch.n_res,
                 "engine_roundtrip_rmsd_bb": bb, "engine_roundtrip_rmsd_all": allr, "engine_roundtrip_rmsd_bb_worst_chain": wbb, "oracle_roundtrip_rmsd_bb_sample": obb, "oracle_roundtrip_rmsd_all_sample": oall,
                 "engine_vs_oracle_decode_bb": dbb, "engine_vs_oracle_decode_max": dmax})
    print(rows[-1], file=sys.stderr, flush=True)
print(json.dumps({"config": f"BASELINE.json configs[4]: {n} chains, lengths 50..2000 (median 280), {batch.n_res} residues, -b sweep", "rows": rows}))
