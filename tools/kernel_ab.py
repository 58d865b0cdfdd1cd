#!/usr/bin/env python3
"""A/B of engine build / run-time variants on bench.py's device-resident round trip.  Every variant runs in its own process
(so that FCZ_ENGINE_LIB can point at another build of the library): creates an engine under the variant's environment, checks
the encode bytes (sha256) against the first variant's, and times the per-kernel events over STEPS round trips.
  VARIANTS="name=ENV1:val,ENV2:val;name2=..."   (the default configuration always runs first)
  LENGTHS=fixed|mixed   STEPS=20"""
import hashlib, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def one():
    import numpy as np, torch
    sys.path.insert(0, ROOT)
    from foldcomp_b200 import abi, synth
    from foldcomp_b200.engine import DeviceBlobBatch, DeviceChainBatch, Engine

    steps = int(os.environ.get("STEPS", "20"))
    dev = torch.device("cuda:0")
    if os.environ.get("LENGTHS", "fixed") == "mixed":
        L = synth.mixed_lengths(np.random.default_rng(synth.SEED), 10000)
    else:
        L = 350
    batch = synth.generate(10000, L, seed=synth.SEED)
    cap = abi.encode_bound(batch.n_chains, batch.n_res, batch.n_atoms, len(batch.titles), 25)
    dbatch = DeviceChainBatch.from_host(batch, dev)
    eng = Engine(0)
    dblob = DeviceBlobBatch(batch.n_chains, cap, dev)
    dout = DeviceChainBatch(batch.n_chains, batch.n_res, batch.n_atoms, len(batch.titles), dev)
    torch.cuda.synchronize()

    def step():
        eng.encode_device(dbatch, dblob); eng.decode_plan_device(dblob, dout); eng.decode_device(dblob, dout)

    for _ in range(3):
        step()
    eng.sync()
    nb = int(dblob.blob_off[-1].item())
    sha = hashlib.sha256(dblob.bytes[:nb].cpu().numpy().tobytes()).hexdigest()[:16]
    xsum = float(dout.xyz.double().abs().sum().item())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); torch.cuda.synchronize()
    for _ in range(steps):
        step()
    eng.sync(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    eng.set_profiling(True); eng.get_profile()
    for _ in range(steps):
        step()
    p = eng.get_profile(); eng.set_profiling(False)
    ker = {k: round(p.kernel_ms[i] / steps, 4) for k, i in abi.PROF_KINDS.items()}
    print(json.dumps({"variant": os.environ.get("AB_NAME", "?"), "ms_per_step": round(ms, 4), "G_res_s": round(batch.n_res / ms / 1e6, 3),
                      "blob_sha": sha, "decode_abs_sum": xsum, "kernel_ms_per_step": ker}), flush=True)
    eng.close()


if __name__ == "__main__":
    if os.environ.get("AB_NAME"):
        one()
        sys.exit(0)
    variants = [("default", {})]
    for spec in os.environ.get("VARIANTS", "").split(";"):
        if spec.strip():
            name, _, kv = spec.partition("=")
            variants.append((name, dict(p.split(":", 1) for p in kv.split(",") if p)))
    first = None
    for name, env in variants:
        e = dict(os.environ, AB_NAME=name, **env)
        r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=e, capture_output=True, text=True, timeout=600)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
        try:
            d = json.loads(line)
        except ValueError:
            print(json.dumps({"variant": name, "env": env, "error": (r.stderr or r.stdout)[-400:]}), flush=True)
            continue
        if first is None:
            first = d
        d["env"] = env
        d["blobs_identical_to_default"] = d["blob_sha"] == first["blob_sha"]
        print(json.dumps(d), flush=True)
