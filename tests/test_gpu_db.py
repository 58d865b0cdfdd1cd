"""Whole-database passes (foldcomp_b200/csrc/fcz_db.cpp: compressDb / decompressDb, the batched form of the
per-entry lambdas of src/main.cpp:438-536 / 612-689) on the GPU, checked entry by entry against the oracle and --
when oracle/_ref/pyref is present -- against the reference's own CPython module and database reader."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import dbutil
import helpers as H
from foldcomp_b200 import abi, pdbio, synth
from foldcomp_b200.abi import HostBlobBatch

pytestmark = pytest.mark.gpu
CLI = os.path.join(H.ROOT, "foldcomp_b200", "csrc", "fcz_cli")


def test_decompress_db_matches_engine_and_oracle(engine, golden, tmp_path):
    lib = dbutil.gpu_host_lib()
    blobs = list(golden.db_blobs) + golden.blobs(25) + [b"garbage entry"]
    keys = list(range(100, 100 + len(blobs)))[::-1]  # written in descending key order: the writer must sort
    entries = [(k, f"e{k}.fcz", b) for k, b in zip(keys, blobs)]
    src, dst = str(tmp_path / "fcz_db"), str(tmp_path / "pdb_db")
    dbutil.write_db(src, entries)
    stats = (C.c_double * 7)()
    assert lib.fczgpu_decompress_db(0, src.encode(), dst.encode(), 0, stats) == 0
    assert int(stats[0]) == len(blobs) and int(stats[1]) == 1
    got = dbutil.read_db(dst)
    assert [g[0] for g in got] == sorted(keys)[1:]  # the garbage entry (largest key... first written) is dropped
    dec = engine.decode_host(HostBlobBatch.from_blobs(blobs))
    by_key = {k: c for c, k in enumerate(keys)}
    ref = dbutil.reference_module()
    for k, name, text in got:
        c = by_key[k]
        assert name == f"e{k}.pdb"
        assert text == H.oracle_format_pdb(dec, c), k  # byte-identical to formatting the engine's decode
        if ref is not None:  # the reference's text of its own decode: same records, coordinates within tolerance
            theirs = ref.decompress(blobs[c])[1].encode("latin-1").split(b"\n")
            ours = text.split(b"\n")
            assert len(ours) == len(theirs)
            for a, b in zip(ours, theirs):
                assert a[:30] == b[:30] and a[54:] == b[54:], (k, a, b)
                if a.startswith(b"ATOM"):
                    assert max(abs(float(a[30 + 8 * j : 38 + 8 * j]) - float(b[30 + 8 * j : 38 + 8 * j])) for j in range(3)) <= 0.051


def test_compress_db_matches_oracle_and_reference_module(engine, tmp_path):
    lib = dbutil.gpu_host_lib()
    lens = np.array([2, 5, 16, 33, 120, 350, 350, 900, 1400, 2100])
    batch = synth.generate(len(lens), lens, seed=44)
    texts = [H.oracle_format_pdb(batch, c) for c in range(batch.n_chains)]
    entries = [(c, f"prot_{c}.pdb", t) for c, t in enumerate(texts)] + [(len(texts), "empty.pdb", b"HEADER nothing\n")]
    src, dst = str(tmp_path / "pdb_db"), str(tmp_path / "fcz_db")
    dbutil.write_db(src, entries)
    stats = (C.c_double * 7)()
    assert lib.fczgpu_compress_db(0, src.encode(), dst.encode(), 25, stats) == 0
    assert int(stats[0]) == len(entries) and int(stats[1]) == 1 and int(stats[2]) == int(lens.sum())
    got = dbutil.read_db(dst)
    assert len(got) == len(texts)
    ref = dbutil.reference_module()
    for (k, name, blob), text in zip(got, texts):
        assert name == f"prot_{k}"  # the entry's base name without extension, like the reference CLI (src/main.cpp:448-449)
        # the title is the text's TITLE record, as in the reference CLI (pdbTitle; src/structure_reader.cpp:31-46)
        title = batch.title(k)
        assert f"TITLE     {title}".encode() in text
        parsed = pdbio.parse_pdb_chain(text.decode(), title)
        assert blob == H.oracle_encode(parsed, 0, 25), k
        if ref is not None:
            assert H.masked(blob) == H.masked(ref.compress(title, text.decode())), k
    if ref is not None:  # the reference's reader opens the database we wrote
        with ref.open(dst) as db:
            assert len(db) == len(texts)
            n0, pdb0 = db[5]
            back = pdbio.parse_pdb_chain(pdb0, "x")
            assert np.abs(back.xyz - batch.chain(5).xyz).max() < 0.5


def test_cli_db_subcommands(tmp_path, golden):
    src = str(tmp_path / "in_db")
    dbutil.write_db(src, [(i, f"d{i}.fcz", b) for i, b in enumerate(golden.db_blobs[:6])])
    mid, back = str(tmp_path / "pdb_db"), str(tmp_path / "fcz_db")
    subprocess.check_call([CLI, "decompress-db", src, mid])
    subprocess.check_call([CLI, "compress-db", mid, back])
    a, b = dbutil.read_db(src), dbutil.read_db(back)
    assert [x[0] for x in a] == [x[0] for x in b] and [x[1].rsplit(".", 1)[0] for x in a] == [x[1] for x in b]  # base names, like the reference CLI
    for (_, _, x), (_, _, y) in zip(a, b):
        dx, dy = H.oracle_decode(x), H.oracle_decode(y)
        assert np.array_equal(dx.res_type, dy.res_type) and H.rmsd(dx.xyz, dy.xyz) < 0.2  # a second lossy generation


def test_python_open_mirrors_reference_database(golden, tmp_path):
    """foldcomp_b200.open() against the reference's FoldcompDatabase (foldcomp.cxx:36-180, 333-433) on the same files."""
    import foldcomp_b200

    names = [f"d{i}x_" for i in range(len(golden.db_blobs))]
    entries = [(i, names[i], b) for i, b in enumerate(golden.db_blobs)]
    path = str(tmp_path / "db")
    dbutil.write_db(path, entries)
    ref = dbutil.reference_module()
    with foldcomp_b200.open(path) as db:
        assert len(db) == len(entries)
        ours = [db[i] for i in range(len(db))]
        assert db[-1] == ours[-1]
        with pytest.raises(IndexError):
            db[len(entries)]
    for (title, pdb), blob in zip(ours, golden.db_blobs):
        want = H.oracle_decode(blob)
        assert title == want.title.decode()
        back = pdbio.parse_pdb_chain(pdb, title)
        assert np.array_equal(back.res_type, want.res_type) and np.abs(back.xyz - want.xyz).max() <= 0.05 + 0.0006
    if ref is not None:
        with ref.open(path) as rdb:
            assert len(rdb) == len(ours)
            for i in (0, 7, len(ours) - 1):
                rname, rpdb = rdb[i]
                assert rname == ours[i][0] and len(rpdb.splitlines()) == len(ours[i][1].splitlines())
    sel = [names[5], "missing_one", names[2]]
    with foldcomp_b200.open(path, ids=sel) as db:
        assert len(db) == 2 and db[0] == ours[5] and db[1] == ours[2]
    with pytest.raises(KeyError):
        foldcomp_b200.open(path, ids=sel, err_on_missing=True)
    with foldcomp_b200.open(path, decompress=False) as db:
        assert db[3] == golden.db_blobs[3]
    with pytest.raises(TypeError):
        foldcomp_b200.open(path, ids="d1asha_")


def test_python_get_data_matches_oracle_and_reference(golden):
    """foldcomp_b200.get_data(fcz): angles, residues and B-factors exact, coordinates within the decode tolerance."""
    import foldcomp_b200

    ref = dbutil.reference_module()
    for blob in golden.blobs(25)[:3] + list(golden.db_blobs[:3]):
        d = foldcomp_b200.get_data(blob)
        ang = H.oracle_unpack_angles(blob)
        L = len(ang)
        f32 = lambda x: np.asarray(x, np.float32)
        assert np.array_equal(f32(d["phi"]), ang[:, 0]) and np.array_equal(f32(d["psi"]), ang[:, 1]) and np.array_equal(f32(d["omega"]), ang[:, 2])
        assert np.array_equal(f32(d["torsion_angles"]), ang[: L - 1][:, [1, 2, 0]].reshape(-1))
        assert np.array_equal(f32(d["bond_angles"]), ang[:, [4, 5, 3]].reshape(-1))
        dec = H.oracle_decode(blob)
        assert np.array_equal(f32(d["b_factors"]), dec.bfactor) and len(d["residues"]) == L
        n_at = len(dec.xyz) + int(dec.meta["has_oxt"])
        assert len(d["coordinates"]) == n_at
        assert np.abs(f32(d["coordinates"])[: len(dec.xyz)] - dec.xyz).max() <= 0.05
        if ref is not None:
            r = ref.get_data(blob)
            assert sorted(r.keys()) == sorted(d.keys())
            for k in ("phi", "psi", "omega", "torsion_angles", "bond_angles", "b_factors", "residues"):
                assert r[k] == d[k], k
            assert len(r["coordinates"]) == len(d["coordinates"])
            assert np.abs(f32(r["coordinates"]) - f32(d["coordinates"])).max() <= 0.05
    with pytest.raises(ValueError):
        foldcomp_b200.get_data("ATOM      1  N   GLY A   1 ...")


def test_python_compress_matches_oracle_and_reference_module(golden):
    """foldcomp_b200.compress(name, pdb_text) -- the CPython module's compress (foldcomp.cxx:253-328) -- byte for byte."""
    import foldcomp_b200

    ref = dbutil.reference_module()
    for c in (golden.names.index("test_af.pdb"), golden.names.index("test.pdb")):
        text = H.oracle_format_pdb(golden.batch, c).decode("latin-1")
        for b in (25, 50):
            blob = foldcomp_b200.compress("entry", text, anchor_residue_threshold=b)
            assert blob == H.oracle_encode(pdbio.parse_pdb_chain(text, "entry"), 0, b)
            if ref is not None:
                assert H.masked(blob) == H.masked(ref.compress("entry", text, anchor_residue_threshold=b))
        name, pdb = foldcomp_b200.decompress(blob)
        assert name == "entry" and pdb.count("\nATOM") + pdb.startswith("ATOM") == text.count("\nATOM") + text.startswith("ATOM")
    with pytest.raises(foldcomp_b200.error):
        foldcomp_b200.compress("x", "HEADER nothing here\n")
    with pytest.raises(TypeError):
        foldcomp_b200.compress("x", text, anchor_residue_threshold="25")


def test_compress_db_splits_chains_and_fragments_like_the_reference_cli(engine, tmp_path):
    """Entries with several chains or breaks in the residue numbering: `foldcomp compress --db` writes one FCZ entry per chain
    / fragment under the entry's name (src/main.cpp:466-517).  compress-db does the same (the GPU parser flags such entries,
    the host cuts them: parsePdbUnits); the two databases hold the same blobs per name."""
    import subprocess

    from test_parse import _host_units

    ref_cli = os.path.join(H.ROOT, "integration", "_build", "foldcomp_ref")
    batch = synth.generate(6, [40, 60, 80, 30, 50, 70], seed=91)
    t = [pdbio.format_pdb(batch.chain(c), 0) for c in range(6)]
    atoms = lambda c, ch=None, shift=0, cut=None: [
        l[:21] + (ch or l[21]) + "%4d" % (int(l[22:26]) + (shift if cut is not None and int(l[22:26]) > cut else 0)) + l[26:]
        for l in t[c].splitlines() if l.startswith("ATOM")]
    texts = {
        "single": t[0],
        "two_chains": "\n".join(atoms(1, "A") + atoms(2, "B")) + "\nEND\n",
        "gap": "\n".join(atoms(3, None, 4, 12)) + "\nEND\n",
        "chains_and_gaps": "\n".join(atoms(4, "X", 9, 20) + atoms(5, "Y", 3, 33)) + "\nEND\n",
        "plain_again": t[1],
    }
    src, dst = str(tmp_path / "pdb_db"), str(tmp_path / "fcz_db")
    dbutil.write_db(src, [(i, n + ".pdb", x.encode()) for i, (n, x) in enumerate(texts.items())])
    stats = (C.c_double * 7)()
    assert dbutil.gpu_host_lib().fczgpu_compress_db(0, src.encode(), dst.encode(), 25, stats) == 0
    got = dbutil.read_db(dst)
    assert [k for k, _, _ in got] == list(range(len(got)))  # a running key per output
    lib = dbutil.gpu_host_lib()
    lib.fczgpu_pdb_title.restype = C.c_int
    lib.fczgpu_pdb_title.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_char_p, C.c_size_t]
    want = []
    for n, x in texts.items():
        buf = C.create_string_buffer(1024)  # the CLI's title rule (TITLE record of the two whole texts, else the entry name)
        tl = lib.fczgpu_pdb_title(x.encode(), len(x.encode()), (n + ".pdb").encode(), buf, 1024)
        title = buf.raw[:tl]
        for rt, bf, xyz, meta in _host_units(x.encode()):
            one = abi.concat_chains([(rt, bf, xyz, np.frombuffer(title, np.uint8), np.array([meta]))])
            want.append((n, H.oracle_encode(one, 0, 25)))
    assert [(n, b) for _, n, b in got] == want
    assert len(want) == 1 + 2 + 2 + 4 + 1 and int(stats[1]) == 0
    if os.path.exists(ref_cli):
        rdst = str(tmp_path / "fcz_ref")
        r = subprocess.run([ref_cli, "compress", "-t", "1", "-y", "--db", src, rdst], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, (r.stdout[-300:], r.stderr[-300:])
        theirs = sorted((n, H.masked(b)) for _, n, b in dbutil.read_db(rdst))
        assert theirs == sorted((n, H.masked(b)) for n, b in want)
