#!/usr/bin/env python3
"""A/B of engine build/run-time variants on bench.py's device-resident round trip: for each environment setting in
VARIANTS (name=ENV1:val,ENV2:val;...) creates an engine, checks encode bytes against the default engine's, and times the
per-kernel events over STEPS round trips."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from foldcomp_b200 import abi, synth
from foldcomp_b200.engine import DeviceBlobBatch, DeviceChainBatch, Engine

steps = int(os.environ.get("STEPS", "20"))
variants = [("default", {})]
for spec in os.environ.get("VARIANTS", "").split(";"):
    if spec.strip():
        name, _, kv = spec.partition("=")
        variants.append((name, dict(p.split(":") for p in kv.split(",") if p)))
dev = torch.device("cuda:0")
lengths = os.environ.get("LENGTHS", "fixed")
if lengths == "mixed":
    L = synth.mixed_lengths(np.random.default_rng(synth.SEED), 10000)
else:
    L = 350
batch = synth.generate(10000, L, seed=synth.SEED)
cap = abi.encode_bound(batch.n_chains, batch.n_res, batch.n_atoms, len(batch.titles), 25)
dbatch = DeviceChainBatch.from_host(batch, dev)
ref_bytes = None
for name, env in variants:
    for k, v in env.items():
        os.environ[k] = v
    eng = Engine(0)
    dblob = DeviceBlobBatch(batch.n_chains, cap, dev)
    dout = DeviceChainBatch(batch.n_chains, batch.n_res, batch.n_atoms, len(batch.titles), dev)
    torch.cuda.synchronize()
    def step():
        eng.encode_device(dbatch, dblob); eng.decode_plan_device(dblob, dout); eng.decode_device(dblob, dout)
    for _ in range(3): step()
    eng.sync()
    nb = int(dblob.blob_off[-1].item())
    got = dblob.bytes[:nb].clone()
    xyz = dout.xyz.clone()
    if ref_bytes is None: ref_bytes, ref_xyz = got, xyz
    same = bool(torch.equal(got, ref_bytes)); dmax = float((xyz - ref_xyz).abs().max().item())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); torch.cuda.synchronize()
    for _ in range(steps): step()
    eng.sync(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    eng.set_profiling(True); eng.get_profile()
    for _ in range(steps): step()
    p = eng.get_profile(); eng.set_profiling(False)
    ker = {k: round(p.kernel_ms[i] / max(p.kernel_launches[i], 1) * (p.kernel_launches[i] / steps), 4) for k, i in abi.PROF_KINDS.items()}
    print(json.dumps({"variant": name, "env": env, "ms_per_step": round(ms, 4), "G_res_s": round(batch.n_res / ms / 1e6, 3), "blobs_identical": same, "decode_max_diff": dmax, "kernel_ms_per_step": ker}), flush=True)
    eng.close()
    for k in env: os.environ.pop(k, None)
