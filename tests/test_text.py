"""PDB text emitter and extract scans -- no GPU.

Three implementations are compared byte for byte:
  * the unmodified reference (oracle/_ref: writeAtomCoordinatesToPDB src/atom_coordinate.cpp:220-291,
    Foldcomp::extract src/foldcomp.cpp:1260-1336) when it is present,
  * the oracle's plain-C restatement (oracle/fcz_oracle.c),
  * the product's formatting code (foldcomp_b200/csrc/fcz_text.h) run by the one-thread model in tests/emu/ --
    the same plan + emit-unit functions the CUDA kernels instantiate,
and all three against committed reference outputs (tests/golden/text_golden.npz, made by make_text_golden.py).
"""
import hashlib
import os

import numpy as np
import pytest

import helpers as H
from foldcomp_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def text_golden():
    return np.load(os.path.join(HERE, "golden", "text_golden.npz"))


def _golden_chains(golden):
    """The decoded golden chains (reference decode of the committed blobs at -b 25) as 1-chain batches."""
    out = []
    for c, name in enumerate(golden.names):
        blob = golden.blobs(25)[c]
        out.append((name, blob, H.decoded_as_batch(H.oracle_decode(blob))))
    return out


def test_oracle_text_matches_committed_reference_output(golden, text_golden):
    for name, blob, b in _golden_chains(golden):
        txt = H.oracle_format_pdb(b)
        assert hashlib.sha256(txt).hexdigest() == str(text_golden[f"pdb_sha256_{name}"]), name
        assert len(txt) == int(text_golden[f"pdb_len_{name}"])
        for t, d in ((0, 1), (0, 2), (0, 3), (0, 4), (1, 0)):
            assert H.oracle_extract(blob, t, d) == bytes(text_golden[f"extract_{name}_{t}_{d}"]), (name, t, d)
    # the full text of the smallest fixture is committed verbatim
    i = golden.names.index("test_af.pdb")
    assert H.oracle_format_pdb(_golden_chains(golden)[i][2]) == bytes(text_golden["pdb_text_test_af.pdb"])


def test_model_text_matches_oracle_on_goldens(golden):
    for name, blob, b in _golden_chains(golden):
        for alt in (False, True):
            assert H.emu_format_pdb(b, 0, alt) == H.oracle_format_pdb(b, 0, alt), (name, alt)
        for t, d in ((0, 1), (0, 2), (0, 3), (0, 4), (1, 0)):
            assert H.emu_extract(blob, t, d) == H.oracle_extract(blob, t, d), (name, t, d)


def test_overflowing_columns_and_long_titles():
    """std::setw is a minimum width: every over-long field must shift the rest of the line exactly as in the reference."""
    for seed in range(3):
        b = H.extreme_text_chain(seed)
        want = H.oracle_format_pdb(b)
        assert H.emu_format_pdb(b) == want
        if H.have_ref():
            assert H.ref_format_pdb(b) == want
        lines = want.split(b"\n")
        assert lines[0].startswith(b"TITLE     a very long") and lines[1].startswith(b"TITLE    2") and lines[2].startswith(b"TITLE    3")
        assert max(len(l) for l in lines) > 81 and lines[-2].startswith(b"TER")


def test_zero_to_one_plddt_scale():
    """extract switches to a 0..1 scale when the largest representable B-factor is <= 1 (src/foldcomp.cpp:1290-1296)."""
    b = synth.generate(1, 40, seed=5)
    b.bfactor = (b.bfactor / np.float32(100.0)).astype(np.float32)
    blob = H.oracle_encode(b, 0, 25)
    for d in (1, 2, 3, 4):
        want = H.oracle_extract(blob, 0, d)
        assert H.emu_extract(blob, 0, d) == want
        if H.have_ref():
            assert H.ref_extract(blob, 0, d) == want


@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built (no /root/reference here)")
def test_oracle_text_matches_reference_live(golden):
    rng = np.random.default_rng(2)
    lens = synth.mixed_lengths(rng, 12, lo=2, hi=600)
    batch = synth.generate(len(lens), lens, seed=77)
    for c in range(batch.n_chains):
        for alt in (False, True):
            assert H.oracle_format_pdb(batch, c, alt) == H.ref_format_pdb(batch, c, alt), (c, alt)
    for name, blob, b in _golden_chains(golden):
        # decode + format exactly as `foldcomp decompress` does (src/main.cpp:612-689)
        assert H.ref_decompress_to_pdb(blob) == H.oracle_format_pdb(b), name
        alt = H.decoded_as_batch(H.oracle_decode(blob, use_alt=True))
        assert H.ref_decompress_to_pdb(blob, True) == H.oracle_format_pdb(alt, 0, True), name
        for t, d in ((0, 1), (0, 2), (0, 3), (0, 4), (1, 0)):
            assert H.ref_extract(blob, t, d) == H.oracle_extract(blob, t, d), (name, t, d)


def test_extract_matches_the_references_committed_fixtures(text_golden):
    """test/test_af.plddt (one digit per residue) and test/test_af.plddt.tsv (four digits) of the reference repository,
    produced upstream from the upstream-encoded test/test_af.fcz."""
    blob = bytes(text_golden["upstream_test_af_fcz"])
    header, digits1 = bytes(text_golden["upstream_test_af_plddt"]).split(b"\n")[:2]
    assert header.startswith(b">")
    name, n_res, digits4 = bytes(text_golden["upstream_test_af_plddt_tsv"]).rstrip(b"\n").split(b"\t")
    for fn in (H.oracle_extract, H.emu_extract):
        assert fn(blob, 0, 1) == digits1
        assert fn(blob, 0, 4) == digits4
        assert len(fn(blob, 1, 0)) == int(n_res)
