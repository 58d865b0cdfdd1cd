#!/usr/bin/env python3
"""Generate tests/golden/golden.npz from the UNMODIFIED reference (oracle/_ref) in this container.

Run here (needs /root/reference and oracle/_ref/libfoldcomp_ref.so); the output is committed so
that the GPU box, which has no /root/reference, can still check against reference outputs.

Contents (all produced by the reference's own Foldcomp::compress/writeStream/read/decompress
through oracle/ref_shim.cpp):
  * inputs: canonical-slot chains parsed from the reference fixtures test/test.pdb,
    test/test_af.pdb, test/multichain.pdb (chain A, and the two fragments of chain B) and 16
    synthetic chains of mixed length (foldcomp_b200.synth, seed 424242);
  * for every input and every anchor threshold in ANCHORS: the reference FCZ bytes and the
    reference's decode of them (coordinates, B-factors);
  * the 24 upstream-encoded blobs of test/example_db with the reference's decode (decode-only).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers as H  # noqa: E402
from foldcomp_b200 import synth  # noqa: E402
from foldcomp_b200.abi import concat_chains  # noqa: E402
from foldcomp_b200.pdbio import _atom_records, canonicalize  # noqa: E402

REF = "/root/reference/test"
ANCHORS = (25, 10, 50, 200)


def fixture_chains():
    parts, names = [], []
    for fn, title in (("test.pdb", "test"), ("test_af.pdb", "test_af")):
        recs = _atom_records(open(os.path.join(REF, fn)).read())
        parts.append(canonicalize(recs, title))
        names.append(fn)
    # multichain.pdb: chain A, and chain B split at its residue-number gap (main.cpp:469-484)
    txt = open(os.path.join(REF, "multichain.pdb")).read().splitlines()
    for ch in "AB":
        lines = [l for l in txt if l.startswith("ATOM") and l[21] == ch]
        recs = _atom_records("\n".join(lines))
        frags, cur = [], [recs[0]]
        for a, b in zip(recs, recs[1:]):
            if b[4] - a[4] > 1:  # identifyDiscontinousResInd, src/atom_coordinate.cpp:506-530
                frags.append(cur)
                cur = []
            cur.append(b)
        frags.append(cur)
        for j, fr in enumerate(frags):
            parts.append(canonicalize(fr, "multichain" + ch + (f"_{j}" if len(frags) > 1 else "")))
            names.append(f"multichain.pdb:{ch}:{j}")
    lens = [2, 3, 7, 24, 25, 26, 49, 50, 51, 64, 100, 129, 257, 350, 400, 700]
    syn = synth.generate(len(lens), np.array(lens), seed=424242)
    parts.append(syn)
    names += [f"synthetic:{l}" for l in lens]
    allp = []
    for p in parts:
        for c in range(p.n_chains):
            allp.append(p._slice(c))
    batch = concat_chains(allp)
    # The shim rebuilds atoms from the canonical slots, so the reference sees table atoms (+OXT) only;
    # keep header nAtom consistent with that (inputs with missing atoms differ only in this field).
    batch.meta["n_atom"] = (np.diff(batch.atom_off.astype(np.int64)) + batch.meta["has_oxt"]).astype(np.uint16)
    return batch, names


def main():
    assert H.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    batch, names = fixture_chains()
    out = {
        "names": np.array(names),
        "res_off": batch.res_off, "atom_off": batch.atom_off, "title_off": batch.title_off,
        "res_type": batch.res_type, "bfactor": batch.bfactor, "xyz": batch.xyz, "titles": batch.titles,
        "meta": batch.meta.view(np.uint8).reshape(batch.n_chains, -1),
        "anchors": np.array(ANCHORS),
    }
    for b in ANCHORS:
        blobs, xyzs, bfs = [], [], []
        for c in range(batch.n_chains):
            blob = H.ref_encode(batch, c, b)
            assert isinstance(blob, bytes), (names[c], b, blob)
            blob = H.masked(blob)  # the four uninitialised padding bytes are not reproducible
            dec = H.ref_decode(blob)
            blobs.append(np.frombuffer(blob, np.uint8))
            xyzs.append(dec.xyz)
            bfs.append(dec.bfactor)
        out[f"fcz_{b}"] = np.concatenate(blobs)
        out[f"fcz_off_{b}"] = np.cumsum([0] + [len(x) for x in blobs]).astype(np.uint64)
        out[f"dec_xyz_{b}"] = np.concatenate(xyzs)
        out[f"dec_bfac_{b}"] = np.concatenate(bfs)
    # upstream-encoded example_db (NUL-terminated entries)
    data = open(os.path.join(REF, "example_db"), "rb").read()
    blobs, xyzs, offs = [], [], [0]
    for line in open(os.path.join(REF, "example_db.index")):
        key, off, ln = (int(x) for x in line.split())
        blob = data[off : off + ln - 1]
        dec = H.ref_decode(blob)
        assert not isinstance(dec, int), key
        blobs.append(np.frombuffer(blob, np.uint8))
        xyzs.append(dec.xyz)
        offs.append(offs[-1] + len(blob))
    out["db_fcz"] = np.concatenate(blobs)
    out["db_fcz_off"] = np.array(offs, np.uint64)
    out["db_dec_xyz"] = np.concatenate(xyzs)
    path = os.path.join(HERE, "golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", batch.n_chains, "chains,", len(blobs), "db blobs")


if __name__ == "__main__":
    main()
