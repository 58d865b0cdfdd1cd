"""The C-ABI library loads on a machine without a GPU and exports every symbol that
include/fcz_engine.h declares (no compute calls here)."""
import os
import re

import helpers as H


def _declared_symbols():
    text = open(os.path.join(H.ROOT, "include", "fcz_engine.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fcz_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from foldcomp_b200._lib import SYMBOLS, load

    lib = load()
    declared = _declared_symbols()
    assert declared == sorted(SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s


def test_host_only_entry_points():
    from foldcomp_b200 import abi
    from foldcomp_b200._lib import load
    from foldcomp_b200.tables import tables

    lib = load()
    tb = tables()
    for c in range(24):
        assert lib.fcz_type_natoms(c) == tb.natoms[c]
        for k in range(tb.natoms[c]):
            assert lib.fcz_type_atom_name(c, k).decode() == tb.atom_names[c][k]
            assert lib.fcz_type_alt_slot(c, k) == tb.alt[c, k]
            if k >= 3:
                assert [lib.fcz_type_pred(c, k, w) for w in range(3)] == list(tb.pred[c, k])
                assert lib.fcz_type_bond_length(c, k) == tb.blen[c, k]
                assert lib.fcz_type_bond_angle(c, k) == tb.bang[c, k]
    assert lib.fcz_strerror(0) == b"ok"
    assert lib.fcz_encode_bound(3, 1000, 7800, 33, 25) == abi.encode_bound(3, 1000, 7800, 33, 25)


def test_product_does_not_reference_oracle():
    """The product tree must not import, link or call anything under oracle/."""
    pkg = os.path.join(H.ROOT, "foldcomp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".h", ".cu", ".cpp", ".c", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "fcz_oracle" not in txt and "libfoldcomp_ref" not in txt and "oracle/" not in txt, os.path.join(dirpath, f)
