#!/usr/bin/env python3
"""Writes tests/golden/text_golden.npz from the UNMODIFIED reference (oracle/_ref, built from /root/reference by
oracle/Makefile): for every chain of golden.npz, the PDB text `foldcomp decompress` produces for its committed -b 25
blob (Foldcomp::read + decompress + writeAtomCoordinatesToPDB) as sha256 + length (the smallest one verbatim), and
Foldcomp::extract's output for pLDDT digits 1..4 and the sequence.  Run in the build container only."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import helpers as H  # noqa: E402

z = np.load(os.path.join(HERE, "golden.npz"))
names = [str(x) for x in z["names"]]
off, data = z["fcz_off_25"], z["fcz_25"]
out = {}
for c, name in enumerate(names):
    blob = bytes(data[int(off[c]) : int(off[c + 1])])
    txt = H.ref_decompress_to_pdb(blob)
    out[f"pdb_sha256_{name}"] = np.array(hashlib.sha256(txt).hexdigest())
    out[f"pdb_len_{name}"] = np.array(len(txt))
    if name == "test_af.pdb":
        out[f"pdb_text_{name}"] = np.frombuffer(txt, np.uint8)
    for t, d in ((0, 1), (0, 2), (0, 3), (0, 4), (1, 0)):
        out[f"extract_{name}_{t}_{d}"] = np.frombuffer(H.ref_extract(blob, t, d), np.uint8)
# the reference's own committed extract fixtures (test/test_af.fcz, test/test_af.plddt, test/test_af.plddt.tsv), verbatim
REF_TEST = "/root/reference/test"
out["upstream_test_af_fcz"] = np.frombuffer(open(os.path.join(REF_TEST, "test_af.fcz"), "rb").read(), np.uint8)
out["upstream_test_af_plddt"] = np.frombuffer(open(os.path.join(REF_TEST, "test_af.plddt"), "rb").read(), np.uint8)
out["upstream_test_af_plddt_tsv"] = np.frombuffer(open(os.path.join(REF_TEST, "test_af.plddt.tsv"), "rb").read(), np.uint8)
np.savez_compressed(os.path.join(HERE, "text_golden.npz"), **out)
print("wrote", len(out), "entries")
