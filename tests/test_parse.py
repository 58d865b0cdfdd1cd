"""SURVEY.md section 8 f3 on the GPU: the per-entry parser of foldcomp_b200/csrc/fcz_parse.h (the algorithm the CUDA kernels
k_parse_* run), on one host thread, against the host parser parsePdbChain (itself pinned to the reference's CPython
module by tests/test_db_host.py) and against strtof for the numeric fields."""
import ctypes as C

import numpy as np
import pytest

import dbutil
import helpers as H
from foldcomp_b200 import abi, pdbio, synth


def _emu_parse(text: bytes):
    lib = H.emu()
    lib.emu_parse_pdb.restype = C.c_int
    lib.emu_parse_pdb.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32]
    cap_r, cap_a = len(text) // 60 + 8, 14 * (len(text) // 60 + 8)
    rt, bf, xyz = np.zeros(cap_r, np.uint8), np.zeros(cap_r, np.float32), np.zeros((cap_a, 3), np.float32)
    meta = np.zeros(1, abi.META_DTYPE)
    nr, na = C.c_uint32(), C.c_uint32()
    rc = lib.emu_parse_pdb(text, len(text), rt.ctypes.data, bf.ctypes.data, xyz.ctypes.data, meta.ctypes.data, C.byref(nr), C.byref(na), cap_r, cap_a)
    return rc, rt[: nr.value], bf[: nr.value], xyz[: na.value], meta[0]


def _host_parse(text: bytes):
    lib = dbutil.gpu_host_lib()
    cap_r, cap_a = len(text) // 60 + 8, 14 * (len(text) // 60 + 8)
    rt, bf, xyz = np.zeros(cap_r, np.uint8), np.zeros(cap_r, np.float32), np.zeros((cap_a, 3), np.float32)
    meta = np.zeros(1, abi.META_DTYPE)
    nr, na = C.c_uint32(), C.c_uint32()
    rc = lib.fczgpu_parse_pdb(text, len(text), rt.ctypes.data, bf.ctypes.data, xyz.ctypes.data, meta.ctypes.data, C.byref(nr), C.byref(na), cap_r, cap_a)
    return rc, rt[: nr.value], bf[: nr.value], xyz[: na.value], meta[0]


def _same(a, b):
    assert a[0] == b[0], (a[0], b[0])
    if a[0] == 0:
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))
        assert np.array_equal(a[3].view(np.uint32), b[3].view(np.uint32)) and a[4].tobytes() == b[4].tobytes()


def test_parser_model_matches_host_parser_on_written_text():
    batch = synth.generate(12, synth.mixed_lengths(np.random.default_rng(2), 12, 30, 500), seed=5)
    for c in range(batch.n_chains):
        text = pdbio.format_pdb(batch.chain(c), 0).encode()
        got = _emu_parse(text)
        _same(got, _host_parse(text))
        assert got[0] == 0 and np.array_equal(got[1], batch.chain(c).res_type) and np.array_equal(got[3], batch.chain(c).xyz)


def messy_variants(golden):
    """Other records between the atoms, CRLF, alternative positions, atoms out of table order, a missing atom, an
    unknown residue, a last atom that is not OXT, short B-factor column, no trailing newline."""
    base = pdbio.format_pdb(golden.batch.chain(golden.names.index("test_af.pdb")), 0)
    lines = base.splitlines()
    atoms = [l for l in lines if l.startswith("ATOM")]
    return {
        "plain": base,
        "crlf": base.replace("\n", "\r\n"),
        "no_final_newline": base.rstrip("\n"),
        "remarks": "HEADER    X\nREMARK 1\n" + "\n".join(lines[:40]) + "\nANISOU junk\nHETATM 9999  O   HOH A 999       0.000   0.000   0.000  1.00  0.00\n" + "\n".join(lines[40:]) + "\nEND\n",
        "altloc": "\n".join(atoms[:5] + [atoms[4]] + [atoms[4][:30] + "   9.999   9.999   9.999" + atoms[4][54:]] + atoms[5:]) + "\n",
        "shuffled": "\n".join(atoms[:3][::-1] + atoms[3:]) + "\n",
        "missing_atom": "\n".join(atoms[:1] + atoms[2:]) + "\n",
        "missing_sidechain_atom": "\n".join(atoms[:4] + atoms[5:]) + "\n",
        "unknown_residue": "\n".join(l[:17] + "XYZ" + l[20:] if 8 <= i < 16 else l for i, l in enumerate(atoms)) + "\n",
        "no_oxt": "\n".join(atoms[:-1]) + "\n",
        "short_b": "\n".join(l[:64] for l in atoms) + "\n",
        "negative_numbers": "\n".join(l[:22] + " -12" + l[26:] if l[22:26] == atoms[0][22:26] else l for l in atoms) + "\n",  # the whole first residue
        "two_chains": "\n".join(atoms[:10] + [l[:21] + "B" + l[22:] for l in atoms[10:20]]) + "\n",
        "no_atoms": "HEADER\nREMARK\nEND\n",
        "empty": "",
        "short_record": "\n".join(atoms[:4] + [atoms[4][:40]] + atoms[5:]) + "\n",
    }


def test_parser_model_on_messy_text(golden):
    variants = messy_variants(golden)
    for name, text in variants.items():
        a, b = _emu_parse(text.encode()), _host_parse(text.encode())
        _same(a, b)
    assert _emu_parse(variants["two_chains"].encode())[0] == 2 and _emu_parse(variants["no_atoms"].encode())[0] == 1
    assert _emu_parse(variants["short_record"].encode())[0] == 3 and _emu_parse(b"")[0] == 1


def test_fixed_float_fields_equal_strtof():
    """Every %8.3f field from -999.999 to 9999.999 (11 M fields) and 4 M random fields of other fixed-point shapes:
    parse_fixed_float == strtof bit for bit; shapes outside the fast grammar are rejected, never mis-parsed."""
    lib = H.emu()
    lib.emu_parse_float_check.restype = C.c_uint64
    lib.emu_parse_float_check.argtypes = [C.c_char_p, C.c_uint32, C.c_uint64, C.POINTER(C.c_uint64)]
    v = np.arange(-999999, 10000000, dtype=np.int64)
    txt = np.char.mod("%8.3f", v / 1000.0)
    buf = "".join(txt.tolist()).encode()
    rej = C.c_uint64()
    assert lib.emu_parse_float_check(buf, 8, len(v), C.byref(rej)) == 0 and rej.value == 0
    rng = np.random.default_rng(3)
    fields = []
    for _ in range(400000):
        d = int(rng.integers(1, 10))
        f = int(rng.integers(0, d + 1))
        digs = "".join(str(x) for x in rng.integers(0, 10, d))
        s = (("-" if rng.random() < 0.3 else "") + digs[: d - f] + ("." + digs[d - f:] if f or rng.random() < 0.5 else ""))
        fields.append(s.rjust(12)[:12])
    buf = "".join(fields).encode()
    assert lib.emu_parse_float_check(buf, 12, len(fields), C.byref(rej)) == 0
    weird = ["1e3".rjust(12), "nan".rjust(12), "0x1p3".rjust(12), "1234567890".rjust(12), "1.2.3".rjust(12), "abc".rjust(12)]
    assert lib.emu_parse_float_check("".join(weird).encode(), 12, len(weird), C.byref(rej)) == 0 and rej.value >= 5
