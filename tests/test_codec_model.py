"""The product's per-chain codec (foldcomp_b200/csrc/fcz_codec.h) instantiated with a one-thread
host context (tests/emu/) against the oracle: this checks, without a GPU, the algorithm the CUDA
kernels run -- phase structure, indexing, bit packing and the anchor-segment stitch decomposition.
Encode must be byte-identical; decode within the BASELINE.md tolerances (backbone RMSD <= 0.01 A,
max deviation <= 0.05 A, all-atom RMSD <= 0.02 A -- measured values are ~1e-4 A)."""
import numpy as np

import helpers as H
from foldcomp_b200 import synth

TOL_BB_RMSD, TOL_MAX, TOL_ALL_RMSD = 0.01, 0.05, 0.02


def test_model_encode_matches_golden(golden):
    for b in golden.anchors:
        blobs = golden.blobs(b)
        for c, name in enumerate(golden.names):
            assert H.emu_encode(golden.batch, c, b) == blobs[c], (name, b)


def test_model_decode_within_tolerance_of_reference(golden):
    worst = 0.0
    for b in golden.anchors:
        blobs = golden.blobs(b)
        for c, name in enumerate(golden.names):
            dec = H.emu_decode(blobs[c])
            xyz, bf = golden.decoded(b, c)
            bb = H.backbone_mask(dec.res_type)
            assert H.rmsd(dec.xyz[bb], xyz[bb]) <= TOL_BB_RMSD, (name, b)
            assert H.max_dev(dec.xyz, xyz) <= TOL_MAX, (name, b)
            assert H.rmsd(dec.xyz, xyz) <= TOL_ALL_RMSD, (name, b)
            assert np.array_equal(dec.bfactor, bf)
            worst = max(worst, H.max_dev(dec.xyz, xyz))
    assert worst < 5e-3  # calibration: the decomposition costs ~1e-4 A, not 1e-2


def test_model_decodes_upstream_example_db(golden):
    a = 0
    for blob in golden.db_blobs:
        dec = H.emu_decode(blob)
        n = len(dec.xyz)
        assert H.max_dev(dec.xyz, golden.db_xyz[a : a + n]) <= TOL_MAX
        a += n


def test_model_vs_oracle_synthetic_mixed():
    rng = np.random.default_rng(11)
    lens = synth.mixed_lengths(rng, 60, lo=2, hi=1200)
    batch = synth.generate(len(lens), lens, seed=3)
    for c in range(batch.n_chains):
        for b in (25, 10, 50, 200):
            o = H.oracle_encode(batch, c, b)
            assert H.emu_encode(batch, c, b) == o, (c, b)
            do, de = H.oracle_decode(o), H.emu_decode(o)
            bb = H.backbone_mask(do.res_type)
            assert H.rmsd(de.xyz[bb], do.xyz[bb]) <= TOL_BB_RMSD
            assert H.max_dev(de.xyz, do.xyz) <= TOL_MAX
            assert de.title == do.title and de.meta == do.meta


def test_model_alt_order(golden):
    c = golden.names.index("test.pdb")
    blob = golden.blobs(25)[c]
    do, de = H.oracle_decode(blob, use_alt=True), H.emu_decode(blob, use_alt=True)
    assert H.max_dev(de.xyz, do.xyz) <= TOL_MAX


def test_model_long_chains_up_to_format_maximum():
    """Chains beyond the shared-memory tiers, up to the format's 16-bit residue count (src/foldcomp.h:118-131)."""
    for L, b in ((2721, 25), (6200, 25), (50000, 200), (65535, 300)):
        batch = H.long_chain(L)
        o = H.oracle_encode(batch, 0, b)
        assert H.emu_encode(batch, 0, b) == o, (L, b)
        do, de = H.oracle_decode(o), H.emu_decode(o)
        bb = H.backbone_mask(do.res_type)
        assert H.rmsd(de.xyz[bb], do.xyz[bb]) <= TOL_BB_RMSD and H.max_dev(de.xyz, do.xyz) <= TOL_MAX, (L, b)


def test_model_encode_degenerate_inputs_match_oracle_and_reference():
    """Inputs on which the reference's arithmetic leaves the beaten path -- constant B-factors (its discretiser divides
    by zero and casts NaN), coincident atoms (NaN cosines), missing atoms at (0,0,0), collinear triples, NaN / inf
    coordinates, integer-lattice coordinates (exact ties), two-valued B-factors: the product codec, the oracle and, when
    present, the unmodified reference must still agree byte for byte."""
    for trial, (kind, batch) in enumerate(H.degenerate_chains()):
        for b in (25, 10):
            o = H.oracle_encode(batch, 0, b)
            assert H.emu_encode(batch, 0, b) == o, (kind, trial, b)
            if H.have_ref():
                assert H.masked(H.ref_encode(batch, 0, b)) == H.masked(o), (kind, trial, b)


def test_model_decode_degenerate_blobs():
    """The product codec's decode on the blobs of the degenerate chains, against the oracle (which equals the unmodified
    reference there).  Same contract as tests/test_gpu_parity.py::test_degenerate_decode: exact fields always; NaN masks
    identical and finite atoms within tolerance unless the input has an exactly collinear backbone triple, where the
    reference's frame normalisation (src/nerf.cpp:52-85) is 0/0 or amplifies sincosf's last bit to Angstroms."""
    n_ill = 0
    chains = H.degenerate_chains()
    for trial, (kind, batch) in enumerate(chains):
        for b in (25, 10):
            blob = H.oracle_encode(batch, 0, b)
            do, de = H.oracle_decode(blob), H.emu_decode(blob)
            if H.have_ref():
                dr = H.ref_decode(blob)
                assert np.array_equal(dr.xyz, do.xyz, equal_nan=True), (kind, trial, b)
            assert np.array_equal(de.res_type, do.res_type) and de.title == do.title
            assert np.array_equal(de.bfactor, do.bfactor, equal_nan=True)
            if H.has_collinear_backbone(batch):
                n_ill += 1
                continue
            why = H.decode_mismatch(de.xyz, do.xyz, do.res_type, TOL_BB_RMSD, TOL_MAX)
            assert why is None, (kind, trial, b, why)
    assert n_ill < len(chains) // 2
