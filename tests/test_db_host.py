"""Host side of the batch path (foldcomp_b200/csrc/fcz_db.{h,cpp}) -- no GPU:
the fixed-column ATOM parser against the reference's CPython module (oracle/_ref/pyref, foldcomp/foldcomp.cxx:253-293)
and the Python mirror in pdbio.py; the certified float-field fast path against strtof; the database reader/writer
against plain-Python files and the reference's own reader."""
import ctypes as C
import ctypes.util
import os
import struct

import numpy as np
import pytest

import dbutil
import helpers as H
from foldcomp_b200 import abi, pdbio, synth


@pytest.fixture(scope="module")
def lib():
    return dbutil.gpu_host_lib()


def cxx_parse(lib, text: bytes, title: str = "t"):
    cap_r, cap_a = 70000, 70000 * 14
    rt, bf, xyz = np.zeros(cap_r, np.uint8), np.zeros(cap_r, np.float32), np.zeros((cap_a, 3), np.float32)
    meta = np.zeros(1, abi.META_DTYPE)
    nr, na = C.c_uint32(), C.c_uint32()
    flag = lib.fczgpu_parse_pdb(text, len(text), rt.ctypes.data, bf.ctypes.data, xyz.ctypes.data, meta.ctypes.data, C.byref(nr), C.byref(na), cap_r, cap_a)
    if flag:
        return flag
    return abi.concat_chains([(rt[: nr.value], bf[: nr.value], xyz[: na.value], np.frombuffer(title.encode(), np.uint8), meta)])


def _texts(golden):
    out = [(n, H.oracle_format_pdb(golden.batch, c)) for c, n in enumerate(golden.names)]
    b = synth.generate(6, np.array([2, 3, 17, 90, 350, 700]), seed=12)
    out += [(f"syn{c}", H.oracle_format_pdb(b, c)) for c in range(b.n_chains)]
    return out


def test_parser_matches_python_mirror_and_reference_module(lib, golden):
    ref = dbutil.reference_module()
    for name, text in _texts(golden):
        # noise the reference's parser skips or resolves: other record types, an alternate location, CRLF-free short lines
        lines = text.split(b"\n")
        atom_i = [i for i, l in enumerate(lines) if l.startswith(b"ATOM")]
        noisy = lines[: atom_i[3]] + [lines[atom_i[2]].replace(b"  1.00", b"  0.50")] + lines[atom_i[3] :]
        noisy = [b"HEADER    something", b"REMARK 1"] + noisy + [b"HETATM 9999  O   HOH A 999       1.000   2.000   3.000  1.00 20.00           O  ", b"END"]
        for t in (text, b"\n".join(noisy)):
            got = cxx_parse(lib, t, "x")
            want = pdbio.parse_pdb_chain(t.decode("latin-1"), "x")
            assert np.array_equal(got.res_type, want.res_type) and np.array_equal(got.xyz, want.xyz)
            assert np.array_equal(got.bfactor, want.bfactor) and got.meta.tobytes() == want.meta.tobytes()
            if ref is not None:
                for b in (25, 200):
                    theirs = ref.compress("x", t.decode("latin-1"), anchor_residue_threshold=b)
                    assert H.masked(H.oracle_encode(got, 0, b)) == H.masked(theirs), (name, b)


def test_parser_flags(lib):
    assert cxx_parse(lib, b"HEADER x\nREMARK\n") == 1  # no ATOM line (foldcomp.cxx:288-290)
    b = synth.generate(1, 5, seed=1)
    text = H.oracle_format_pdb(b, 0)
    two = text.replace(b" A   4 ", b" B   4 ")
    assert cxx_parse(lib, two) == 2  # multiple chains (foldcomp.cxx:266-268)
    assert cxx_parse(lib, text[:200] + b"\nATOM      9  N   GLY A\n") == 3  # a truncated ATOM record (the reference throws)


def test_float_field_fast_path_is_strtof(lib):
    libc = C.CDLL(ctypes.util.find_library("c"))
    libc.strtof.restype = C.c_float
    libc.strtof.argtypes = [C.c_char_p, C.c_void_p]
    rng = np.random.default_rng(0)
    fields = [b"   0.000", b"  -0.000", b"9999.999", b"-999.999", b"  12.5  ", b"1.5e3   ", b" .5     ", b"     nan", b"  +3.250", b"1234567.", b"16777217", b"99999999", b"0.000001"]
    for _ in range(200000):
        w = int(rng.integers(1, 9))
        v = rng.uniform(-10 ** min(w, 4), 10 ** min(w, 5))
        fields.append((b"%8.3f" % v)[:8] if rng.random() < 0.8 else (b"%8.*f" % (int(rng.integers(0, 7)), v))[:8])
    for f in fields:
        a, b = lib.fczgpu_parse_float(f, len(f)), libc.strtof(f, None)
        assert (a == b) or (a != a and b != b), (f, a, b)


def test_db_reader_writer_roundtrip(lib, golden, tmp_path):
    entries = [(k, f"name_{k}.fcz", golden.db_blobs[i]) for i, k in enumerate([7, 3, 11, 0, 5])]
    src = str(tmp_path / "src_db")
    dbutil.write_db(src, entries, nul=False)  # `compress --db` of the reference writes no terminator
    dst = str(tmp_path / "dst_db")
    assert lib.fczgpu_db_copy(src.encode(), dst.encode()) == len(entries)
    got = dbutil.read_db(dst)
    assert got == sorted(entries)  # sorted by key on close; payloads intact; names kept
    raw = open(dst, "rb").read()
    assert sum(len(e[2]) + 1 for e in entries) == len(raw)  # every entry NUL-terminated
    ref = dbutil.reference_module()
    if ref is not None:  # the reference's own reader opens what we wrote and decodes every entry
        with ref.open(dst) as db:
            assert len(db) == len(entries)
            for i, (k, name, blob) in enumerate(sorted(entries)):
                n, pdb = db[i]
                assert pdb == ref.decompress(blob)[1]


def test_oracle_angles_match_reference_get_data(golden):
    """The oracle's restatement of decompressBackboneChain against the reference's own get_data() (foldcomp.cxx:497-640)."""
    ref = dbutil.reference_module()
    if ref is None:
        pytest.skip("oracle/_ref/pyref not built")
    for blob in golden.blobs(25) + list(golden.db_blobs[:4]):
        d = ref.get_data(blob)
        ang = H.oracle_unpack_angles(blob)
        L = len(ang)
        f32 = lambda x: np.asarray(x, np.float32)
        assert np.array_equal(f32(d["phi"]), ang[:, 0]) and np.array_equal(f32(d["psi"]), ang[:, 1]) and np.array_equal(f32(d["omega"]), ang[:, 2])
        assert np.array_equal(f32(d["torsion_angles"]), ang[: L - 1][:, [1, 2, 0]].reshape(-1))
        assert np.array_equal(f32(d["bond_angles"]), ang[:, [4, 5, 3]].reshape(-1))
        dec = H.oracle_decode(blob)
        assert np.array_equal(f32(d["b_factors"]), dec.bfactor)
        assert d["residues"] == "".join("ARNDCQEGHILKMFPSTWYVBZ*X"[c] for c in dec.res_type)


def test_python_database_raw_mode_and_ids(golden, tmp_path):
    """foldcomp_b200.open(..., decompress=False): the reference's raw mode (foldcomp.cxx:52-80) needs no GPU."""
    import foldcomp_b200

    entries = [(k, f"name_{k}", golden.db_blobs[i]) for i, k in enumerate([4, 9, 2, 7])]
    path = str(tmp_path / "db")
    dbutil.write_db(path, entries)
    with foldcomp_b200.open(path, decompress=False) as db:
        by_key = [e[2] for e in sorted(entries)]  # entries come in key order, as from the reference's reader
        assert len(db) == 4 and [db[i] for i in range(4)] == by_key and db[-1] == by_key[-1]
        with pytest.raises(IndexError):
            db[4]
    with foldcomp_b200.open(path, ids=["name_7", "nope", "name_4"], decompress=False) as db:
        assert len(db) == 2 and db[0] == entries[3][2] and db[1] == entries[0][2]
    with pytest.raises(KeyError):
        foldcomp_b200.open(path, ids=["nope"], decompress=False, err_on_missing=True)
    for bad in ({"ids": "name_4"}, {"decompress": 1}, {"err_on_missing": "yes"}):
        with pytest.raises(TypeError):
            foldcomp_b200.open(path, **bad)
    ref = dbutil.reference_module()
    if ref is not None:
        with ref.open(path, decompress=False) as rdb, foldcomp_b200.open(path, decompress=False) as db:
            assert len(rdb) == len(db) and all(rdb[i] == db[i] for i in range(len(db)))
        with ref.open(path, ids=["name_7", "name_4"], decompress=False) as rdb:
            assert [rdb[i] for i in range(len(rdb))] == [entries[3][2], entries[0][2]]


def test_db_copy_handles_unsorted_and_gapped_files(lib, tmp_path):
    """Entries out of key order in the index, gaps between entries, and no lookup file."""
    src = str(tmp_path / "src")
    blobs = {5: b"five", 1: b"one-one", 3: b""}
    with open(src, "wb") as d:
        d.write(b"XX" + blobs[5] + b"\0" + b"GAP" + blobs[1] + b"\0" + blobs[3] + b"\0")
    with open(src + ".index", "w") as ix:
        ix.write("5\t2\t5\n1\t10\t8\n3\t18\t1\n")
    dst = str(tmp_path / "dst")
    assert lib.fczgpu_db_copy(src.encode(), dst.encode()) == 3
    got = dbutil.read_db(dst)
    assert got == [(1, "1", b"one-one"), (3, "3", b""), (5, "5", b"five")]  # names default to the key
    assert lib.fczgpu_db_copy((src + "_missing").encode(), dst.encode()) < 0


def test_reference_style_db_handles(tmp_path):
    """make_reader ... writer_append (src/database_reader.h:11-27, src/database_writer.h:12-15) exported by
    libfoldcomp_gpu.so over DbReader / DbWriter: same signatures (C++ linkage, as in the reference) and conventions --
    ids are positions in the key-sorted index, lengths include the entry's NUL, lookups by name and by key, absent ->
    -1 / UINT32_MAX / "", writer_append stores the data as given and free_writer sorts by key."""
    import ctypes as C

    lib = C.CDLL(dbutil.GPU_SO)
    f = lambda name, res, args: (setattr(getattr(lib, name), "restype", res), setattr(getattr(lib, name), "argtypes", args), getattr(lib, name))[2]
    make_reader = f("_Z11make_readerPKcS0_i", C.c_void_p, [C.c_char_p, C.c_char_p, C.c_int32])
    free_reader = f("_Z11free_readerPv", None, [C.c_void_p])
    get_id = f("_Z13reader_get_idPvj", C.c_int64, [C.c_void_p, C.c_uint32])
    get_data = f("_Z15reader_get_dataPvl", C.c_void_p, [C.c_void_p, C.c_int64])
    get_key = f("_Z14reader_get_keyPvl", C.c_uint32, [C.c_void_p, C.c_int64])
    get_len = f("_Z17reader_get_lengthPvl", C.c_int64, [C.c_void_p, C.c_int64])
    get_off = f("_Z17reader_get_offsetPvl", C.c_int64, [C.c_void_p, C.c_int64])
    get_size = f("_Z15reader_get_sizePv", C.c_int64, [C.c_void_p])
    lookup_entry = f("_Z19reader_lookup_entryPvPKc", C.c_uint32, [C.c_void_p, C.c_char_p])
    lookup_name = f("_Z24reader_lookup_name_allocPvj", C.c_char_p, [C.c_void_p, C.c_uint32])
    make_writer = f("_Z11make_writerPKcS0_", C.c_void_p, [C.c_char_p, C.c_char_p])
    free_writer = f("_Z11free_writerPv", None, [C.c_void_p])
    append = f("_Z13writer_appendPvPKcmjS1_", C.c_bool, [C.c_void_p, C.c_char_p, C.c_size_t, C.c_uint32, C.c_char_p])

    db = str(tmp_path / "db")
    w = make_writer(db.encode(), (db + ".index").encode())
    entries = [(7, "gamma", b"third entry\0"), (2, "alpha", b"first\0"), (5, "beta", b"second one\0")]  # keys out of order
    for k, n, d in entries:
        assert append(w, d, len(d), k, n.encode())
    free_writer(w)
    assert open(db + ".dbtype", "rb").read() == struct.pack("<i", 12)
    assert [l.split("\t")[0] for l in open(db + ".index").read().splitlines()] == ["2", "5", "7"]
    assert open(db + ".lookup").read().splitlines() == ["2\talpha\t0", "5\tbeta\t0", "7\tgamma\t0"]
    assert open(db, "rb").read() == b"".join(d for _, _, d in entries)  # data in append order, nothing added

    r = make_reader(db.encode(), (db + ".index").encode(), 1 | 4 | 8)
    assert r and get_size(r) == 3
    by_key = {k: (n, d) for k, n, d in entries}
    for pos, k in enumerate((2, 5, 7)):
        assert get_id(r, k) == pos and get_key(r, pos) == k
        n, d = by_key[k]
        assert get_len(r, pos) == len(d)
        assert C.string_at(get_data(r, pos), get_len(r, pos)) == d
        assert open(db, "rb").read()[get_off(r, pos):get_off(r, pos) + len(d)] == d
        assert lookup_entry(r, n.encode()) == k and lookup_name(r, k) == n.encode()
    assert get_id(r, 3) == -1 and get_data(r, 9) is None and get_len(r, -1) == -1
    assert lookup_entry(r, b"delta") == 0xFFFFFFFF and lookup_name(r, 3) == b""
    free_reader(r)
    r = make_reader(db.encode(), (db + ".index").encode(), 0)  # no data, no lookup
    assert get_size(r) == 3 and get_data(r, 0) is None and lookup_entry(r, b"alpha") == 0xFFFFFFFF
    free_reader(r)
    assert make_reader((db + "_missing").encode(), (db + "_missing.index").encode(), 1) is None
    # the reference's own reader (its CPython module) opens what the handle API wrote
    ref = dbutil.reference_module()
    if ref is not None:
        with ref.open(db, decompress=False) as d2:
            assert len(d2) == 3 and [bytes(x).rstrip(b"\0") if not isinstance(x, str) else x.rstrip("\0") for x in d2] in (
                [b"first", b"second one", b"third entry"], ["first", "second one", "third entry"])


def test_writer_batch_append_both_modes(tmp_path, monkeypatch):
    """DbWriter::appendBatch (the text side of decompress-db): entries of one slab written by all host threads, through a
    shared mapping or by pwrite, between ordinary appends -- every entry NUL-terminated, index in key order, skipped
    entries absent."""
    import ctypes as C

    lib = dbutil.gpu_host_lib()
    lib.fczgpu_db_write_batch.restype = C.c_int
    lib.fczgpu_db_write_batch.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_uint32, C.c_void_p]
    rng = np.random.default_rng(4)
    lens = rng.integers(0, 70000, 300)
    lens[:4] = [0, 1, 4095, 4096]
    off = np.zeros(len(lens) + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    slab = rng.integers(1, 255, int(off[-1]), dtype=np.uint8).tobytes()
    skip = (rng.random(len(lens)) < 0.1).astype(np.uint8)
    for mode in ("mmap", "pwrite"):
        monkeypatch.setenv("FCZ_DB_WRITE", mode)  # mmap is the opt-in
        path = str(tmp_path / f"db_{mode}")
        assert lib.fczgpu_db_write_batch(path.encode(), slab, off.ctypes.data, len(lens), skip.ctypes.data) == 0
        got = dbutil.read_db(path)
        want = [(1, "head", b"head"), (2, "tail", b"tail")] + [
            (100 + c, f"e{c}", slab[int(off[c]) : int(off[c + 1])]) for c in range(len(lens)) if not skip[c]
        ]
        assert got == want, mode
        assert os.path.getsize(path) == sum(len(b) + 1 for _, _, b in want)


def test_python_database_hands_full_entries_to_the_decoder(golden, tmp_path, monkeypatch):
    """decompress=True: every entry reaches the engine at its full indexed length -- with its terminator when the database
    has them (the decoder sizes a blob from its header), complete when it has none (`foldcomp compress --db` and
    `fcz_cli compress-db` write none); raw mode keeps the reference's one-byte strip.  The engine is a stub here: the
    decoding itself is tests/test_gpu_db.py's business."""
    import foldcomp_b200
    from foldcomp_b200 import abi, database

    seen = []

    class StubEngine:
        def decode_to_pdb_host(self, blobs):
            seen.append(blobs.blobs())
            n = blobs.n_chains
            return abi.HostTextBatch(np.arange(n + 1, dtype=np.uint64), np.frombuffer(b"x" * n, np.uint8).copy(), np.zeros(n, np.int32))

    monkeypatch.setattr(foldcomp_b200, "_get_engine", lambda: StubEngine())
    blobs = list(golden.db_blobs[:3])
    for nul in (True, False):
        path = str(tmp_path / f"db_{int(nul)}")
        dbutil.write_db(path, [(k, f"n{k}", b) for k, b in enumerate(blobs)], nul=nul)
        seen.clear()
        with database.FoldcompDatabase(path) as db:
            name, text = db[1]
            assert text == "x" and isinstance(name, str)
        assert seen and seen[0] == [b + (b"\0" if nul else b"") for b in blobs[1:]]  # the window opens at the entry asked for
        with foldcomp_b200.open(path, decompress=False) as db:
            assert [db[i] for i in range(3)] == [(b if nul else b[:-1]) for b in blobs]
