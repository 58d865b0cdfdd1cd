import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import helpers as H
from foldcomp_b200 import synth
from foldcomp_b200.abi import HostBlobBatch
from foldcomp_b200.engine import Engine
from foldcomp_b200.tables import tables

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
L = int(sys.argv[2]) if len(sys.argv) > 2 else 350
batch = synth.generate(n, L, seed=5)
want = H.oracle_encode_batch(batch, 25)
ref = H.oracle_decode_batch(want)
with Engine(0) as eng:
    dec = eng.decode_host(HostBlobBatch(want.blob_off, want.bytes))
d = np.sqrt(((dec.xyz.astype(np.float64) - ref.xyz) ** 2).sum(1))
bad = np.nonzero(~(d < 0.05))[0]
print("atoms", len(d), "bad", len(bad), "nan", int(np.isnan(dec.xyz).any(axis=1).sum()))
tb = tables()
nat = tb.natoms[ref.res_type]
off = np.concatenate([[0], np.cumsum(nat)])
res_of = np.searchsorted(off, bad, side="right") - 1
slots = bad - off[res_of]
for a, r, s in list(zip(bad, res_of, slots))[:40]:
    print("atom", a, "res", r, "(in chain:", r % L, ") slot", s, "dev", d[a], dec.xyz[a], ref.xyz[a])
print("bad slots histogram", np.bincount(slots, minlength=14) if len(bad) else None)
print("bad residues-in-chain", sorted(set((res_of % L).tolist()))[:60])
