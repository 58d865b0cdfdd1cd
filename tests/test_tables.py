"""The integer-slot residue table (foldcomp_b200/csrc/fcz_tables.h) against a dump of the reference's
own AminoAcid::AminoAcids() (src/amino_acid.h:69-406) committed as tests/golden/aa_table_dump.txt."""
import os
import subprocess

import helpers as H


def test_table_matches_reference_dump(tmp_path):
    exe = tmp_path / "dump_my"
    subprocess.check_call(["gcc", "-O1", "-o", str(exe), os.path.join(H.ROOT, "tests", "dump_my_tables.c")])
    mine = subprocess.check_output([str(exe)]).decode()
    gold = open(os.path.join(H.ROOT, "tests", "golden", "aa_table_dump.txt")).read()
    assert mine == gold


def test_reference_dump_is_current():
    """When the reference is present, the committed dump must equal a fresh one."""
    exe = os.path.join(H.ROOT, "oracle", "_ref", "dump_tables")
    if not (os.path.isdir(H.REFERENCE_DIR) and os.path.exists(exe)):
        import pytest

        pytest.skip("reference not present")
    fresh = subprocess.check_output([exe]).decode()
    assert fresh == open(os.path.join(H.ROOT, "tests", "golden", "aa_table_dump.txt")).read()


def test_python_tables_agree_with_header():
    from foldcomp_b200.tables import tables

    tb = tables()
    assert list(tb.natoms[:20]) == [5, 11, 8, 8, 6, 9, 9, 4, 10, 8, 8, 9, 8, 11, 7, 6, 7, 14, 12, 7]
    assert tb.natoms[23] == 3 and tb.natoms[20] == 0
    assert float(tb.natoms[:20].mean()) > 7  # sanity
    # torsion counts per type (src/foldcomp.cpp:1761-1807) = atoms - 3
    assert [int(n) - 3 for n in tb.natoms[:20]] == [2, 8, 5, 5, 3, 6, 6, 1, 7, 5, 5, 6, 5, 8, 4, 3, 4, 11, 9, 4]
