#!/usr/bin/env python3
"""Device-resident throughput of the text kernels on bench.py's workload: 10 000 decoded chains of
350 residues -> PDB text (k_pdb_plan + k_pdb_emit), and pLDDT / sequence extraction from the FCZ blobs.  Prints one
JSON line; kernel times from the engine's per-launch CUDA events, bytes = text written + coordinates read."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from foldcomp_b200 import abi, synth  # noqa: E402
from foldcomp_b200.engine import DeviceBlobBatch, DeviceChainBatch, DeviceTextBatch, Engine  # noqa: E402

n_chains = int(os.environ.get("N_CHAINS", "10000"))
steps = int(os.environ.get("STEPS", "10"))
dev = torch.device("cuda:0")
batch = synth.generate(n_chains, 350, seed=synth.SEED)
eng = Engine(0)
cap = abi.encode_bound(batch.n_chains, batch.n_res, batch.n_atoms, len(batch.titles), 25)
dbatch = DeviceChainBatch.from_host(batch, dev)
dblob = DeviceBlobBatch(batch.n_chains, cap, dev)
dout = DeviceChainBatch(batch.n_chains, batch.n_res, batch.n_atoms, len(batch.titles), dev)
torch.cuda.synchronize()
eng.encode_device(dbatch, dblob)
eng.decode_plan_device(dblob, dout)
eng.decode_device(dblob, dout)
dtext = DeviceTextBatch(batch.n_chains, 16, dev)
total = eng.pdb_text_plan_device(dout, dtext)
dtext.bytes = torch.zeros(total, dtype=torch.uint8, device=dev)
eng.pdb_text_device(dout, dtext)
eng.sync()
for _ in range(2):
    eng.pdb_text_plan_device(dout, dtext)
    eng.pdb_text_device(dout, dtext)
eng.sync()
eng.set_profiling(True)
eng.get_profile()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
ev0.record()
torch.cuda.synchronize()
for _ in range(steps):
    eng.pdb_text_plan_device(dout, dtext)
    eng.pdb_text_device(dout, dtext)
eng.sync()
ev1.record()
torch.cuda.synchronize()
wall_ms = ev0.elapsed_time(ev1) / steps
p = eng.get_profile()
eng.set_profiling(False)
plan_ms = p.kernel_ms[6] / max(p.kernel_launches[6], 1)
emit_ms = p.kernel_ms[7] / max(p.kernel_launches[7], 1)
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
peak = float(peaks.get("hbm_gbs", 6650.0))
emit_bytes = total + 12 * batch.n_atoms + 5 * batch.n_res
# spot check against the host formatter on one chain
txt = dtext.to_host()
from foldcomp_b200 import pdbio  # noqa: E402

dec = dout.to_host()
assert txt.text(17).decode() == pdbio.format_pdb(dec, 17), "GPU text differs from the host formatter"
print(json.dumps({
    "workload": f"{n_chains} decoded chains x 350 residues -> PDB text", "text_bytes": int(total), "atoms": int(batch.n_atoms),
    "k_pdb_plan_ms": plan_ms, "k_pdb_emit_ms": emit_ms, "plan_plus_emit_wall_ms": wall_ms,
    "emit_algorithmic_bytes": int(emit_bytes), "emit_gbs": emit_bytes / emit_ms / 1e6, "hbm_peak_gbs": peak,
    "emit_frac_of_hbm": emit_bytes / emit_ms / 1e6 / peak, "residues_per_s": batch.n_res / (wall_ms * 1e-3),
}))
