"""Per-phase cycle breakdown of k_encode / k_decode (debug build libfcz_engine_timing.so)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import foldcomp_b200._lib as L
L.LIB_PATH = os.path.join(ROOT, "foldcomp_b200", "csrc", "libfcz_engine_timing.so")
import numpy as np, torch
from foldcomp_b200 import abi, synth
from foldcomp_b200.engine import Engine, DeviceChainBatch, DeviceBlobBatch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
batch = synth.generate(n, 350, seed=1)
dev = torch.device("cuda:0")
eng = Engine(0)
db = DeviceChainBatch.from_host(batch, dev)
bl = DeviceBlobBatch(n, abi.encode_bound(n, batch.n_res, batch.n_atoms, len(batch.titles), 25), dev)
do = DeviceChainBatch(n, batch.n_res, batch.n_atoms, len(batch.titles), dev)
torch.cuda.synchronize()
out = (C.c_ulonglong * 32)()
f = eng.lib.fcz_debug_phase_cycles
f.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
for it in range(3):
    eng.encode_device(db, bl); eng.decode_plan_device(bl, do); eng.decode_device(bl, do)
    f(eng.h, out)
names = {0: "enc scan+stage wait", 1: "enc side-chain bytes", 2: "enc backbone angles+min/max", 3: "enc pack", 4: "enc copy-out", 8: "dec unpack", 9: "dec fwd/rev passes",
         10: "dec stitch", 11: "dec blend", 12: "dec side chains", 13: "dec stage-in wait", 14: "dec copy-out"}
for k in sorted(names):
    if out[16 + k]:
        print(f"{names[k]:24s} {out[k] / out[16 + k]:10.0f} cycles/chain  (n={out[16 + k]})")
