"""Backbone angles before quantisation -- Foldcomp::preprocess's backboneTorsionAngles / backboneBondAngles
(src/foldcomp.cpp:484-496), the lists get_data(pdb_text) returns (foldcomp/foldcomp.cxx:633-671).  CPU: the oracle's
restatement against tests/golden/getdata_golden.npz (made from the reference's own CPython module by
tests/golden/make_getdata_golden.py) and the one-thread model of the product code (enc_raw_angles) against the oracle.
GPU: k_raw_angles through the C ABI and foldcomp_b200.get_data(pdb_text) against both."""
import os

import numpy as np
import pytest

import helpers as H
from foldcomp_b200 import abi, pdbio, synth

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ("torsion_angles", "bond_angles", "phi", "psi", "omega")


def _golden_getdata():
    return np.load(os.path.join(HERE, "golden", "getdata_golden.npz"))


def _parsed(golden, c):
    """The chain as get_data reads it: through its PDB text (coordinates rounded by the %8.3f columns)."""
    return pdbio.parse_pdb_chain(pdbio.format_pdb(golden.batch.chain(c), 0), "")


def _same_bits(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_oracle_and_model_match_reference_get_data(golden):
    g = _golden_getdata()
    n = 0
    for c, name in enumerate(golden.names):
        if f"{name}|phi" not in g:
            continue
        one = _parsed(golden, c)
        for fn in (H.oracle_backbone_angles, H.emu_backbone_angles):
            lists = H.get_data_lists(fn(one, 0))
            for k, v in zip(KEYS, lists):
                assert _same_bits(v, g[f"{name}|{k}"]), (fn.__name__, name, k)
        assert _same_bits(one.bfactor, g[f"{name}|b_factors"]), name
        n += 1
    assert n >= 15


def test_model_matches_oracle_on_synthetic_and_degenerate():
    batch = synth.generate(30, synth.mixed_lengths(np.random.default_rng(3), 30, 1, 600), seed=21)
    for c in range(batch.n_chains):
        assert _same_bits(H.emu_backbone_angles(batch, c), H.oracle_backbone_angles(batch, c)), c
    for _, one in H.degenerate_chains():  # NaN / inf / coincident / collinear atoms: the same NaNs in the same places
        a, b = H.emu_backbone_angles(one, 0), H.oracle_backbone_angles(one, 0)
        nan = np.isnan(a) | np.isnan(b)
        assert np.array_equal(np.isnan(a), np.isnan(b))
        assert np.array_equal(a[~nan].view(np.uint32), b[~nan].view(np.uint32))


@pytest.mark.gpu
def test_gpu_backbone_angles_match_oracle(engine):
    batch = abi.concat_batches([synth.generate(300, synth.mixed_lengths(np.random.default_rng(5), 300, 1, 1500), seed=9), H.long_chain(9000)])
    got = engine.backbone_angles_host(batch)
    for c in range(batch.n_chains):
        r0, r1 = int(batch.res_off[c]), int(batch.res_off[c + 1])
        assert _same_bits(got[r0:r1], H.oracle_backbone_angles(batch, c)), c


@pytest.mark.gpu
def test_python_get_data_from_pdb_text_matches_reference(engine, golden):
    import foldcomp_b200

    g = _golden_getdata()
    n = 0
    for c, name in enumerate(golden.names):
        if f"{name}|phi" not in g:
            continue
        text = pdbio.format_pdb(golden.batch.chain(c), 0)
        d = foldcomp_b200.get_data(text)
        for k in KEYS + ("b_factors",):
            assert _same_bits(d[k], g[f"{name}|{k}"]), (name, k)
        assert d["residues"] == str(g[f"{name}|residues"]) and len(d["coordinates"]) == int(g[f"{name}|n_coordinates"])
        assert _same_bits(np.array(d["coordinates"][:8], np.float32), g[f"{name}|coordinates_head"])
        n += 1
    assert n >= 15
    with pytest.raises(ValueError):
        foldcomp_b200.get_data("HEADER\nEND\n")
