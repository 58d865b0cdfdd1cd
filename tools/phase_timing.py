#!/usr/bin/env python3
"""Per-phase cycle shares of the encode / decode kernels from the debug build libfcz_engine_timing.so (clock64 marks of
thread 0 of every CTA, summed over chains; make -C foldcomp_b200/csrc libfcz_engine_timing.so).  Never loaded by the package."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["FCZ_ENGINE_LIB"] = os.path.join(ROOT, "foldcomp_b200", "csrc", "libfcz_engine_timing.so")
sys.path.insert(0, ROOT)
import numpy as np, torch
from foldcomp_b200 import abi, synth
from foldcomp_b200.engine import DeviceBlobBatch, DeviceChainBatch, Engine

NAMES = {0: "enc scan", 1: "enc side chains", 2: "enc backbone (float first)", 5: "enc decide", 6: "enc exact list", 7: "enc params + list q",
         3: "enc pack", 4: "enc copy-out", 8: "dec unpack", 9: "dec passes", 10: "dec stitch", 11: "dec blend", 12: "dec side", 13: "dec stage-in", 14: "dec copy-out"}
dev = torch.device("cuda:0")
batch = synth.generate(10000, 350, seed=synth.SEED)
cap = abi.encode_bound(batch.n_chains, batch.n_res, batch.n_atoms, len(batch.titles), 25)
dbatch = DeviceChainBatch.from_host(batch, dev)
eng = Engine(0)
dblob = DeviceBlobBatch(batch.n_chains, cap, dev)
dout = DeviceChainBatch(batch.n_chains, batch.n_res, batch.n_atoms, len(batch.titles), dev)
buf = (C.c_ulonglong * 32)()
eng.lib.fcz_debug_phase_cycles.argtypes = [C.c_void_p, C.c_void_p]
for it in range(3):
    eng.encode_device(dbatch, dblob); eng.decode_plan_device(dblob, dout); eng.decode_device(dblob, dout)
    eng.lib.fcz_debug_phase_cycles(eng.h, buf)
cyc, cnt = np.array(buf[:16], float), np.array(buf[16:], float)
enc = sum(cyc[i] for i in (0, 1, 2, 3, 4, 5, 6, 7)); dec = sum(cyc[i] for i in range(8, 15))
out = {}
for i, n in NAMES.items():
    if cnt[i]:
        out[n] = {"cycles_per_chain": round(cyc[i] / cnt[i], 1), "share": round(cyc[i] / (enc if i < 8 else dec), 4)}
print(json.dumps(out, indent=1))
