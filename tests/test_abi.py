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


def test_ctypes_mirrors_match_the_header_layout(tmp_path):
    """sizeof / offsetof of every struct of include/fcz_engine.h, as a C compiler sees the header, against the ctypes mirrors
    of foldcomp_b200/abi.py (what a binding in any other language would have to reproduce), and the error / validity codes."""
    import ctypes as C
    import subprocess

    from foldcomp_b200 import abi

    mirrors = {
        "fcz_opts": abi.FczOpts, "fcz_chain_meta": abi.FczChainMeta, "fcz_chain_batch": abi.FczChainBatch, "fcz_blob_batch": abi.FczBlobBatch,
        "fcz_text_batch": abi.FczTextBatch, "fcz_sizes": abi.FczSizes, "fcz_profile": abi.FczProfile,
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(H.ROOT, "include", "fcz_engine.h")}"', "int main(void) {"]
    for name, cls in mirrors.items():
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for field, _ in cls._fields_:
            lines.append(f'  printf("{name}.{field} %zu\\n", offsetof({name}, {field}));')
    consts = ["FCZ_OK", "FCZ_E_MAGIC", "FCZ_E_TRUNCATED", "FCZ_E_RESIDUE", "FCZ_E_LIMIT", "FCZ_E_CAPACITY", "FCZ_E_CUDA", "FCZ_E_ARG",
              "FCZ_E_PARSE_NOATOM", "FCZ_E_PARSE_CHAINS", "FCZ_E_PARSE_RECORD", "FCZ_E_PARSE_NUMBER", "FCZ_E_PARSE_GAPS", "FCZ_MEM_HOST", "FCZ_MEM_DEVICE"]
    for c in consts:
        lines.append(f'  printf("{c} %d\\n", (int){c});')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-o", str(exe), str(src)])  # the header is plain C
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for name, cls in mirrors.items():
        assert int(got[name]) == C.sizeof(cls), name
        for field, _ in cls._fields_:
            assert int(got[f"{name}.{field}"]) == getattr(cls, field).offset, (name, field)
    for c in consts:
        assert int(got[c]) == getattr(abi, c), c
