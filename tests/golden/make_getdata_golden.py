#!/usr/bin/env python3
"""Writes tests/golden/getdata_golden.npz from the UNMODIFIED reference's CPython module (oracle/_ref/pyref, built from
/root/reference/foldcomp/foldcomp.cxx by oracle/Makefile): get_data(pdb_text) -- getDataFromPDB, foldcomp.cxx:633-671 --
for the PDB text of every chain of golden.npz (the text is foldcomp_b200.pdbio.format_pdb of the committed chain, so the
tests rebuild it without the reference's files): torsion_angles, bond_angles, phi, psi, omega, b_factors, residues and the
number of coordinates.  Run in the build container only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import dbutil  # noqa: E402
from foldcomp_b200 import abi, pdbio  # noqa: E402

ref = dbutil.reference_module()
assert ref is not None, "oracle/_ref/pyref is not built (make -C oracle pyref)"
z = np.load(os.path.join(HERE, "golden.npz"))
batch = abi.HostChainBatch(res_off=z["res_off"], atom_off=z["atom_off"], title_off=z["title_off"], res_type=z["res_type"],
                           bfactor=z["bfactor"], xyz=z["xyz"], titles=z["titles"],
                           meta=np.ascontiguousarray(z["meta"]).view(abi.META_DTYPE).reshape(-1))
names = [str(x) for x in z["names"]]
out = {}
for c, name in enumerate(names):
    if batch.res_off[c + 1] - batch.res_off[c] > 700:
        continue  # the two small fixtures and the synthetic chains are enough
    text = pdbio.format_pdb(batch.chain(c), 0)
    d = ref.get_data(text)
    for k in ("torsion_angles", "bond_angles", "phi", "psi", "omega", "b_factors"):
        out[f"{name}|{k}"] = np.array(d[k], np.float32)
    out[f"{name}|residues"] = np.array(d["residues"])
    out[f"{name}|n_coordinates"] = np.array(len(d["coordinates"]))
    out[f"{name}|coordinates_head"] = np.array(d["coordinates"][:8], np.float32)
np.savez_compressed(os.path.join(HERE, "getdata_golden.npz"), **out)
print("wrote", len(out), "entries for", len(set(k.split("|")[0] for k in out)), "chains")
