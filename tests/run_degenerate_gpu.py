#!/usr/bin/env python3
"""GPU engine vs oracle on degenerate inputs (the generator of test_codec_model.py's degenerate test), one batch."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers as H
from foldcomp_b200 import abi, synth
from foldcomp_b200.engine import Engine

rng = np.random.default_rng(7)
parts, kinds = [], []
for trial in range(120):
    L = int(rng.integers(2, 60))
    batch = synth.generate(1, L, seed=1000 + trial)
    x, bf = batch.xyz.copy(), batch.bfactor.copy()
    A, kind = len(x), trial % 10
    if kind == 0: bf[:] = 50.0
    elif kind == 1: x[rng.integers(0, A)] = x[rng.integers(0, A)]
    elif kind == 2: x[rng.integers(0, A, 3)] = 0.0
    elif kind == 3: x *= np.float32(100.0)
    elif kind == 4: x[2] = x[1] + (x[1] - x[0])
    elif kind == 5: x[rng.integers(0, A)] = np.nan
    elif kind == 6: bf[rng.integers(0, L)] = np.nan
    elif kind == 7: x[:] = np.round(x)
    elif kind == 8: x[rng.integers(0, A)] = np.inf
    else: bf[:] = rng.choice([0.0, 100.0], L)
    batch.xyz, batch.bfactor = x, bf
    parts.append(batch); kinds.append(kind)
big = abi.concat_batches(parts)
with Engine(0) as eng:
    for b in (25, 10):
        eng.set_opts(anchor_threshold=b)
        got = eng.encode_host(big)
        want = H.oracle_encode_batch(big, b)
        bad = [(c, kinds[c]) for c in range(big.n_chains) if got.blob(c) != want.blob(c)]
        print("b", b, "status nonzero", int(np.count_nonzero(got.status)), "mismatching chains", bad[:12], "of", big.n_chains)
