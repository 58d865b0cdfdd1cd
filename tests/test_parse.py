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


def _stoi(field: str) -> int:
    """parse_int_field of the parsers: blanks, sign, leading digits (0 without any)."""
    import re

    m = re.match(r"[ \t]*([+-]?)(\d*)", field)
    v = int(m.group(2)) if m.group(2) else 0
    return -v if m.group(1) == "-" else v


def cli_fragments(text: bytes) -> bool:
    """True when `foldcomp compress` would not treat the (single-chain) text as ONE unit: the first kept atom is not an N, or
    an N atom's residue number exceeds the previous N atom's by more than one (identifyDiscontinousResInd,
    src/atom_coordinate.cpp:506-530) -- the GPU parser's flag 5."""
    prev_name, first, last_n = None, True, None
    for line in text.decode("latin-1").split("\n"):
        if not line.startswith("ATOM") or len(line) < 61:
            continue
        name = line[12:16].strip(" \t")
        if name == prev_name:  # removeAlternativePosition
            continue
        prev_name = name
        if first and name != "N":
            return True
        first = False
        if name == "N":
            num = _stoi(line[22:26])
            if last_n is not None and num - last_n > 1:
                return True
            last_n = num
    return False


def _host_units(text: bytes):
    """parsePdbUnits: [(res_type, bfactor, xyz, meta)] per chain / fragment, or the flag."""
    lib = dbutil.gpu_host_lib()
    lib.fczgpu_parse_pdb_units.restype = C.c_int
    lib.fczgpu_parse_pdb_units.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint32), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64]
    cap_u, cap_r = 256, len(text) // 60 + 8
    cap_a = 14 * cap_r
    ro, ao = np.zeros(cap_u + 1, np.uint32), np.zeros(cap_u + 1, np.uint64)
    rt, bf, xyz = np.zeros(cap_r, np.uint8), np.zeros(cap_r, np.float32), np.zeros((cap_a, 3), np.float32)
    meta = np.zeros(cap_u, abi.META_DTYPE)
    nu = C.c_uint32()
    rc = lib.fczgpu_parse_pdb_units(text, len(text), C.byref(nu), ro.ctypes.data, ao.ctypes.data, rt.ctypes.data, bf.ctypes.data, xyz.ctypes.data,
                                    meta.ctypes.data, cap_u, cap_r, cap_a)
    if rc:
        return rc
    return [(rt[ro[u] : ro[u + 1]].copy(), bf[ro[u] : ro[u + 1]].copy(), xyz[int(ao[u]) : int(ao[u + 1])].copy(), meta[u].copy()) for u in range(nu.value)]


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
        "negative_numbers": "\n".join(l[:22] + "%4d" % (int(l[22:26]) - 20) + l[26:] for l in atoms) + "\n",  # contiguous, from -19 on
        "two_chains": "\n".join(atoms[:10] + [l[:21] + "B" + l[22:] for l in atoms[10:20]]) + "\n",
        "no_atoms": "HEADER\nREMARK\nEND\n",
        "empty": "",
        "short_record": "\n".join(atoms[:4] + [atoms[4][:40]] + atoms[5:]) + "\n",
        "numbering_gap": "\n".join(l[:22] + "%4d" % (int(l[22:26]) + (5 if int(l[22:26]) > 12 else 0)) + l[26:] for l in atoms) + "\n",
        "numbering_step_back": "\n".join(l[:22] + "%4d" % (int(l[22:26]) - (7 if int(l[22:26]) > 12 else 0)) + l[26:] for l in atoms) + "\n",
    }


def test_parser_model_on_messy_text(golden):
    variants = messy_variants(golden)
    for name, text in variants.items():
        a, b = _emu_parse(text.encode()), _host_parse(text.encode())
        if a[0] == 5:  # the CLI would cut the text into fragments: the single-chain host parser does not look for that
            assert b[0] == 0 and cli_fragments(text.encode()), name
            continue
        assert not (b[0] == 0 and cli_fragments(text.encode())), name
        _same(a, b)
    assert _emu_parse(variants["shuffled"].encode())[0] == 5 and _emu_parse(variants["negative_numbers"].encode())[0] == 0
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


def test_host_units_match_reference_cli(golden, tmp_path):
    """parsePdbUnits (what compress-db does with an entry the GPU parser flags as several chains / fragments) against the
    reference's own CLI on the same file: `foldcomp compress` writes one .fcz per chain and fragment
    (src/main.cpp:466-530); every one of them equals the oracle's encoding of the matching unit."""
    import os
    import subprocess

    cli = os.path.join(H.ROOT, "integration", "_build", "foldcomp_ref")
    if not os.path.exists(cli):
        pytest.skip("integration/_build/foldcomp_ref not built")
    v = messy_variants(golden)
    base_atoms = [l for l in v["plain"].splitlines() if l.startswith("ATOM")]
    chain_b = [l[:21] + "B" + l[22:] for l in v["numbering_gap"].splitlines() if l.startswith("ATOM")]
    # chain A whole, chain B with a numbering gap, chain C starting in the middle of a residue (its first atoms are dropped)
    chain_c = [l[:21] + "C" + l[22:] for l in base_atoms[2:]]
    text = "\n".join(base_atoms + chain_b + chain_c) + "\nEND\n"
    indir, out = tmp_path / "in", tmp_path / "out"
    indir.mkdir()
    out.mkdir()
    (indir / "multi.pdb").write_text(text)
    r = subprocess.run([cli, "compress", "-y", str(indir), str(out)], capture_output=True, text=True, timeout=120)
    files = sorted(os.listdir(out))
    assert r.returncode == 0 and files, (r.stdout[-300:], r.stderr[-300:])
    units = _host_units(text.encode())
    assert isinstance(units, list) and len(units) == 4  # A, B_0, B_1, C
    want = []
    for rt, bf, xyz, meta in units:
        one = abi.concat_chains([(rt, bf, xyz, np.frombuffer(b"multi", np.uint8), np.array([meta]))])
        want.append(H.masked(H.oracle_encode(one, 0, 25)))
    got = [H.masked(open(out / f, "rb").read()) for f in files]
    assert sorted(got) == sorted(want), files
    # and the single-chain texts keep their flags
    assert _host_units(b"HEADER\n") == 1 and len(_host_units(v["plain"].encode())) == 1 and len(_host_units(v["numbering_gap"].encode())) == 2
    assert len(_host_units(v["numbering_step_back"].encode())) == 1 and _emu_parse(v["numbering_step_back"].encode())[0] == 0
    assert _emu_parse(v["numbering_gap"].encode())[0] == 5


def test_title_rule_matches_reference_cli(golden, tmp_path):
    """pdbTitle (the title compress-db stores in every blob) against the reference CLI: HEADER idCode, else TITLE records,
    else the file name without extension (src/structure_reader.cpp:31-46, src/main.cpp:466-467)."""
    import os
    import subprocess

    cli = os.path.join(H.ROOT, "integration", "_build", "foldcomp_ref")
    if not os.path.exists(cli):
        pytest.skip("integration/_build/foldcomp_ref not built")
    lib = dbutil.gpu_host_lib()
    lib.fczgpu_pdb_title.restype = C.c_int
    lib.fczgpu_pdb_title.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_char_p, C.c_size_t]
    atoms = "\n".join(l for l in messy_variants(golden)["plain"].splitlines() if l.startswith("ATOM")) + "\nEND\n"
    header = lambda idc: ("HEADER    HYDROLASE                               01-JAN-20   " + idc).ljust(80) + "\n"
    heads = {
        "none": "",
        "remark_only": "REMARK   1 nothing\n",
        "title": "TITLE     A SHORT TITLE\n",
        "title_padded": "TITLE     PADDED TO EIGHTY COLUMNS".ljust(80) + "\n",
        "title_two_lines": "TITLE     FIRST LINE OF THE TITLE WHICH IS LONG ENOUGH TO CONTINUE ON THE NEXT".ljust(80) + "\nTITLE    2 LINE, AND ENDS HERE".ljust(80) + "\n",
        "header_id": header("1ABC") + "TITLE     IGNORED BECAUSE OF THE ID\n",
        "header_blank_id": header("    ") + "TITLE     TAKEN BECAUSE THE ID IS BLANK\n",
        "header_short": "HEADER    TOO SHORT FOR AN ID\nTITLE     FROM TITLE\n",
        "crlf": "TITLE     WITH CARRIAGE RETURN\r\n",
    }
    indir, out = tmp_path / "in", tmp_path / "out"
    indir.mkdir()
    out.mkdir()
    for k, h in heads.items():
        (indir / f"{k}.pdb").write_text(h + atoms, newline="")
    r = subprocess.run([cli, "compress", "-y", str(indir), str(out)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.stdout[-300:], r.stderr[-300:])
    for k, h in heads.items():
        blob = open(out / f"{k}.fcz", "rb").read()
        tl = int.from_bytes(blob[24:28], "little")
        t0 = 76 + 4 * blob[12]
        want = blob[t0 : t0 + tl]
        buf = C.create_string_buffer(4096)
        text = (h + atoms).encode()
        n = lib.fczgpu_pdb_title(text, len(text), f"{k}.pdb".encode(), buf, 4096)
        assert n >= 0 and buf.raw[:n] == want, (k, buf.raw[:n], want)


def test_parser_model_fuzz_against_host_parser():
    """Random damage to PDB text (bytes flipped, runs deleted / inserted / duplicated, truncation, NULs, stray ATOM tags,
    exponents): the GPU parser's per-entry algorithm (one host thread) never disagrees with the host parser except in the
    ways it is allowed to -- it rejects number shapes outside its grammar (flag 4), flags what the CLI would cut (5), and
    may name another of the fatal flags 1..3 when a text has several faults (it looks at every line, the host parser stops
    at the first).  The same loop was run under AddressSanitizer / UBSan for 18 000 mutated texts without a report."""
    import random

    rng = random.Random(20260925)
    batch = synth.generate(6, [3, 12, 40, 25, 8, 60], seed=5)
    bases = [pdbio.format_pdb(batch.chain(c), 0).encode() for c in range(6)]
    same = rejected = 0
    for it in range(1200):
        t = bytearray(rng.choice(bases))
        for _ in range(rng.randint(1, 6)):
            if not t:
                break
            op, i = rng.randint(0, 7), rng.randrange(len(t))
            if op == 0:
                t[i] = rng.randrange(256)
            elif op == 1:
                del t[i : i + rng.randint(1, 90)]
            elif op == 2:
                t[i:i] = bytes(rng.randrange(256) for _ in range(rng.randint(1, 40)))
            elif op == 3:
                t[i:i] = b"\n" * rng.randint(1, 5)
            elif op == 4:
                t = t[:i]
            elif op == 5:
                t[i:i] = t[max(0, i - 200) : i]
            elif op == 6:
                t[i] = ord(rng.choice(" \t\r\n\0-+.eE0123456789A"))
            else:
                t[i : i + 1] = rng.choice([b"ATOM", b"\nATOM", b"nan", b"1e5", b"0x1p3", b"          "])
        t = bytes(t)
        a, b = _emu_parse(t), _host_parse(t)
        if a[0] == 4:
            rejected += 1
            continue
        if a[0] == 5:
            assert b[0] == 0 and cli_fragments(t), it
            continue
        if a[0] != b[0]:
            assert {a[0], b[0]} <= {1, 2, 3}, (it, a[0], b[0])
            continue
        if a[0] == 0:
            assert not cli_fragments(t), it
            _same(a, b)
            same += 1
    assert same > 100 and rejected > 50
