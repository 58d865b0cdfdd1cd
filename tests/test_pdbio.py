"""Host-side text <-> canonical conversion (foldcomp_b200/pdbio.py) -- no GPU."""
import numpy as np
import pytest

import helpers as H
from foldcomp_b200 import pdbio

PDB = """\
TITLE     tiny
ATOM      1  N   GLY A   5      -0.966   0.493   1.500  1.00 11.00           N  
ATOM      2  CA  GLY A   5       0.257   0.418   0.692  1.00 12.50           C  
ATOM      3  C   GLY A   5      -0.094   0.017  -0.716  1.00 13.00           C  
ATOM      4  O   GLY A   5      -1.056  -0.682  -0.923  1.00 14.00           O  
ATOM      5  N   ALA A   6       0.661   0.439  -1.742  1.00 21.00           N  
ATOM      6  CA  ALA A   6       0.435   0.099  -3.135  1.00 22.25           C  
ATOM      7  C   ALA A   6       1.734  -0.322  -3.810  1.00 23.00           C  
ATOM      8  O   ALA A   6       2.770   0.251  -3.520  1.00 24.00           O  
ATOM      9  OXT ALA A   6       1.700  -1.300  -4.600  1.00 25.00           O  
TER
"""


def test_parse_slots_missing_atoms_and_oxt():
    b = pdbio.parse_pdb_chain(PDB, "tiny")
    assert b.n_res == 2 and list(b.res_type) == [7, 0]
    assert b.n_atoms == 4 + 5  # GLY 4 slots, ALA 5 slots (CB missing -> zeros)
    assert np.array_equal(b.xyz[8], [0, 0, 0])
    assert np.allclose(b.bfactor, [12.5, 22.25])
    m = b.meta[0]
    assert m["n_atom"] == 9 and m["idx_residue"] == 5 and m["idx_atom"] == 1 and m["chain"] == ord("A")
    assert m["has_oxt"] == 1 and np.allclose(m["oxt"], [1.7, -1.3, -4.6])
    assert b.title(0) == "tiny"


def test_parse_errors_mirror_reference_flags():
    with pytest.raises(pdbio.PdbError, match="No ATOM"):
        pdbio.parse_pdb_chain("HEADER x\n", "t")
    two = PDB.replace("ALA A   6", "ALA B   6")
    with pytest.raises(pdbio.PdbError, match="Multiple chains"):
        pdbio.parse_pdb_chain(two, "t")


def test_alt_location_removed():
    lines = PDB.splitlines(keepends=True)
    dup = lines[:3] + [lines[2].replace("0.257", "9.999")] + lines[3:]
    b = pdbio.parse_pdb_chain("".join(dup), "t")
    assert np.allclose(b.xyz[1], [0.257, 0.418, 0.692])
    assert b.meta[0]["n_atom"] == 9


def test_ftoa_matches_reference_rounding():
    assert pdbio._ftoa(1.2345, 1000, 3) == "1.235" or pdbio._ftoa(1.2345, 1000, 3) == "1.234"
    assert pdbio._ftoa(-0.0004, 1000, 3) == "-0.000"
    assert pdbio._ftoa(12.5, 100, 2) == "12.50"
    assert pdbio._ftoa(-7.25, 100, 2) == "-7.25"


def test_format_roundtrips_through_parser(golden):
    c = golden.names.index("test_af.pdb")
    dec = H.oracle_decode(golden.blobs(25)[c])
    from foldcomp_b200 import abi

    meta = np.zeros(1, abi.META_DTYPE)
    meta[0] = dec.meta
    b = abi.concat_chains([(dec.res_type, dec.bfactor, dec.xyz, np.frombuffer(dec.title, np.uint8), meta)])
    txt = pdbio.format_pdb(b, 0)
    assert txt.startswith("TITLE     test_af\nATOM      1  N   MET A   1 ")
    assert txt.rstrip().splitlines()[-1].startswith("TER")
    back = pdbio.parse_pdb_chain(txt, "test_af")
    assert np.array_equal(back.res_type, dec.res_type)
    assert np.abs(back.xyz - dec.xyz).max() <= 0.00051
    assert back.meta[0]["has_oxt"] == 1


def test_native_parser_face_equals_python_mirror(golden):
    """foldcomp_b200.pdbnative (ctypes over the C++ parser, what compress() uses) against pdbio on texts with noise."""
    from foldcomp_b200 import pdbnative

    assert pdbnative.available()
    for c in range(len(golden.names)):
        text = H.oracle_format_pdb(golden.batch, c).decode("latin-1")
        a, b = pdbnative.parse_pdb_chain(text, "t"), pdbio.parse_pdb_chain(text, "t")
        for f in ("res_off", "atom_off", "title_off", "res_type", "bfactor", "xyz", "titles"):
            assert np.array_equal(getattr(a, f), getattr(b, f)), (c, f)
        assert a.meta.tobytes() == b.meta.tobytes()
    a, b = pdbnative.parse_pdb_chain(PDB, "tiny"), pdbio.parse_pdb_chain(PDB, "tiny")
    assert np.array_equal(a.xyz, b.xyz) and a.meta.tobytes() == b.meta.tobytes() and a.n_atoms == 9
    with pytest.raises(pdbio.PdbError, match="No ATOM"):
        pdbnative.parse_pdb_chain("HEADER x\n", "t")
    with pytest.raises(pdbio.PdbError, match="Multiple chains"):
        pdbnative.parse_pdb_chain(PDB.replace("ALA A   6", "ALA B   6"), "t")


def test_split_pdb_by_chain_like_the_reference_helper():
    """foldcomp_b200.split_pdb_by_chain against the reference package's pure-Python helper (foldcomp/util.py) where it is at
    hand, and against spelled-out expectations everywhere."""
    import importlib.util
    import os

    import foldcomp_b200

    a = "ATOM      1  N   MET A   1      -1.000   2.000   3.000  1.00 50.00           N  "
    b = a[:21] + "B" + a[22:]
    text = "HEADER x\n" + a + "\n" + a + "\nTER\n" + b + "\nHETATM junk\n" + a + "\nEND\n"
    want = [a + "\n" + a + "\n", b + "\n", a + "\n"]
    assert foldcomp_b200.split_pdb_by_chain(text) == want
    assert foldcomp_b200.split_pdb_by_chain("HEADER only\n") == [""] and foldcomp_b200.split_pdb_by_chain("") == [""]
    ref_util = "/root/reference/foldcomp/util.py"
    if os.path.exists(ref_util):
        spec = importlib.util.spec_from_file_location("ref_util", ref_util)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        for t in (text, "HEADER only\n", "", a, a + "\n" + b):
            assert foldcomp_b200.split_pdb_by_chain(t) == m.split_pdb_by_chain(t), t
