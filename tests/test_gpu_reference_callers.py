"""the reference's OWN callers -- src/main.cpp (the `foldcomp` CLI) and foldcomp/foldcomp.cxx (the
CPython module), both unmodified -- linked against the B200 engine through integration/foldcomp_on_engine.cpp (class
Foldcomp implemented on include/fcz_engine.h), run side by side with the unmodified reference builds of the same
sources (integration/_build/foldcomp_ref, oracle/_ref/pyref).  Mirrors /root/reference/build.sh:25-38 (minimal_test:
compress, decompress, rmsd goldens), the db / extract / check modes of src/main.cpp and
/root/reference/test/test_foldcomp.py:1-40.  The binaries are built here by integration/Makefile (__graft_entry__.build)
and travel to the GPU box; /root/reference itself is never read at test time."""
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers as H
from foldcomp_b200 import pdbio, synth

pytestmark = pytest.mark.gpu
BUILD = os.path.join(H.ROOT, "integration", "_build")
GPU_CLI, REF_CLI = os.path.join(BUILD, "foldcomp_gpu"), os.path.join(BUILD, "foldcomp_ref")
PYGPU, PYREF = os.path.join(BUILD, "pygpu"), os.path.join(H.ROOT, "oracle", "_ref", "pyref")

needs_cli = pytest.mark.skipif(not (os.path.exists(GPU_CLI) and os.path.exists(REF_CLI)), reason="integration/_build not built")
needs_py = pytest.mark.skipif(not (os.path.exists(os.path.join(PYGPU, "foldcomp.so")) and os.path.exists(os.path.join(PYREF, "foldcomp.so"))),
                              reason="integration/_build/pygpu or oracle/_ref/pyref not built")


def run(cli, *args, cwd=None):
    r = subprocess.run([cli, *map(str, args)], capture_output=True, text=True, cwd=cwd, timeout=600)
    assert r.returncode == 0, (cli, args, r.stdout[-400:], r.stderr[-400:])
    return r


def xyz_of(pdb_text):
    return np.array([[float(l[30:38]), float(l[38:46]), float(l[46:54])] for l in pdb_text.splitlines() if l.startswith("ATOM")])


@needs_cli
@pytest.mark.parametrize("name,want_rmsd", [("test.pdb", 0.0826751), ("test_af.pdb", None)])
def test_cli_minimal_test_like_build_sh(golden, tmp_path, name, want_rmsd):
    """build.sh:25-38: compress -> decompress -> rmsd, golden 0.0826751 +- 0.001 for test/test.pdb."""
    ch = golden.batch.chain(golden.names.index(name))
    (tmp_path / "in.pdb").write_text(pdbio.format_pdb(ch, 0))
    for cli, tag in ((GPU_CLI, "gpu"), (REF_CLI, "ref")):
        run(cli, "compress", "-y", tmp_path / "in.pdb", tmp_path / f"{tag}.fcz")
        run(cli, "decompress", "-y", tmp_path / f"{tag}.fcz", tmp_path / f"{tag}.pdb")
    # encode: same bytes as the unmodified reference CLI (its four uninitialised padding bytes masked)
    assert H.masked((tmp_path / "gpu.fcz").read_bytes()) == H.masked((tmp_path / "ref.fcz").read_bytes())
    # decode: same text layout, coordinates within tolerance (+ 3-decimal text rounding)
    g, r = (tmp_path / "gpu.pdb").read_text(), (tmp_path / "ref.pdb").read_text()
    assert [l[:30] + l[54:] for l in g.splitlines()] == [l[:30] + l[54:] for l in r.splitlines()]
    d = np.sqrt(((xyz_of(g) - xyz_of(r)) ** 2).sum(axis=1))
    assert d.max() <= 0.05 + 0.002 and np.sqrt((d ** 2).mean()) <= 0.02
    # the reference's own rmsd tool on the GPU build's output (column 6 = all-atom RMSD)
    cols = run(REF_CLI, "rmsd", tmp_path / "in.pdb", tmp_path / "gpu.pdb").stdout.strip().split("\t")
    cols_ref = run(REF_CLI, "rmsd", tmp_path / "in.pdb", tmp_path / "ref.pdb").stdout.strip().split("\t")
    assert cols[2:4] == cols_ref[2:4]  # residues, atoms
    assert abs(float(cols[5]) - float(cols_ref[5])) <= 1e-3 and abs(float(cols[4]) - float(cols_ref[4])) <= 1e-3
    if want_rmsd is not None:
        assert abs(float(cols[5]) - want_rmsd) <= 1e-3, cols


@needs_cli
def test_cli_directory_db_extract_check(tmp_path):
    """A directory of PDB files through `compress` (directory -> directory and directory -> --db), `decompress`,
    `extract --plddt / --fasta` and `check`, GPU build against reference build: identical FCZ bytes, identical extract
    output, identical check verdicts, decompressed coordinates within tolerance."""
    batch = synth.generate(12, synth.mixed_lengths(np.random.default_rng(5), 12, 40, 400), seed=77)
    src = tmp_path / "pdbs"
    src.mkdir()
    for c in range(batch.n_chains):
        (src / f"chain{c:02d}.pdb").write_text(pdbio.format_pdb(batch.chain(c), 0))
    out = {}
    for cli, tag in ((GPU_CLI, "gpu"), (REF_CLI, "ref")):
        d = tmp_path / tag
        d.mkdir()
        # relative paths from inside the build's own directory: extract prints the input path as the entry's title
        run(cli, "compress", "-t", "4", "-y", src, "fcz", cwd=d)
        run(cli, "compress", "-t", "4", "-y", "--db", src, "db", cwd=d)
        run(cli, "decompress", "-t", "4", "-y", "fcz", "pdb", cwd=d)
        run(cli, "extract", "-t", "2", "-y", "--plddt", "fcz", "plddt", cwd=d)
        run(cli, "extract", "-t", "2", "-y", "--fasta", "fcz", "fasta", cwd=d)
        out[tag] = d
        r = run(cli, "check", "-t", "2", "fcz", cwd=d)
        out[tag + "_check"] = sorted(r.stderr.splitlines())
    names = sorted(os.listdir(out["ref"] / "fcz"))
    assert names == sorted(os.listdir(out["gpu"] / "fcz")) and len(names) == 12
    for n in names:
        assert H.masked((out["gpu"] / "fcz" / n).read_bytes()) == H.masked((out["ref"] / "fcz" / n).read_bytes()), n
    for kind in ("plddt", "fasta"):  # one FASTA-like file each; entry order depends on the thread schedule
        def records(p):
            t = p.read_text().split(">")
            return sorted(x for x in t if x)
        assert records(out["gpu"] / kind) == records(out["ref"] / kind) and len(records(out["ref"] / kind)) == 12, kind
    for n in sorted(os.listdir(out["ref"] / "pdb")):
        g, r = (out["gpu"] / "pdb" / n).read_text(), (out["ref"] / "pdb" / n).read_text()
        assert np.abs(xyz_of(g) - xyz_of(r)).max() <= 0.05 + 0.002, n
    assert out["gpu_check"] == out["ref_check"]
    # the two databases hold the same entries (the reference's writer, unmodified, in both builds)
    from dbutil import read_db

    dg = {n: b for _, n, b in read_db(str(out["gpu"] / "db"))}
    dr = {n: b for _, n, b in read_db(str(out["ref"] / "db"))}
    assert sorted(dg) == sorted(dr) and len(dg) == 12
    for k in dr:
        assert H.masked(dg[k]) == H.masked(dr[k]), k


PY_SNIPPET = r"""
import sys, json, hashlib
sys.path.insert(0, sys.argv[1])
import foldcomp
pdb = open(sys.argv[2]).read()
fcz = foldcomp.compress("test", pdb)
name, text = foldcomp.decompress(fcz)
d = foldcomp.get_data(fcz)
out = {"fcz": fcz.hex(), "name": name, "text": text, "keys": sorted(d.keys()), "phi": d["phi"], "psi": d["psi"], "omega": d["omega"],
       "residues": d["residues"], "b_factors": d["b_factors"], "n_coords": len(d["coordinates"])}
with foldcomp.open(sys.argv[3]) as db:
    out["db"] = [(n, hashlib.sha256(t.encode()).hexdigest() if False else t[:200]) for n, t in db]
    out["db_len"] = len(db)
with foldcomp.open(sys.argv[3], ids=["chain03", "chain07"]) as db:
    out["db_ids"] = [n for n, _ in db]
print(json.dumps(out))
"""


@needs_py
def test_python_module_of_the_reference_on_the_engine(tmp_path):
    """/root/reference/test/test_foldcomp.py:1-40 (compress, decompress, open all / ids / str) plus get_data, through the
    reference's CPython module linked against the engine, compared with the same module linked against the reference."""
    import json

    batch = synth.generate(10, 120, seed=31)
    (tmp_path / "in.pdb").write_text(pdbio.format_pdb(batch.chain(0), 0))
    src = tmp_path / "pdbs"
    src.mkdir()
    for c in range(batch.n_chains):
        (src / f"chain{c:02d}.pdb").write_text(pdbio.format_pdb(batch.chain(c), 0))
    run(REF_CLI, "compress", "-y", "--db", src, tmp_path / "db")
    res = {}
    for tag, path in (("gpu", PYGPU), ("ref", PYREF)):
        r = subprocess.run([sys.executable, "-c", PY_SNIPPET, path, str(tmp_path / "in.pdb"), str(tmp_path / "db")],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, (tag, r.stderr[-800:])
        res[tag] = json.loads(r.stdout.strip().splitlines()[-1])
    g, r = res["gpu"], res["ref"]
    assert H.masked(bytes.fromhex(g["fcz"])) == H.masked(bytes.fromhex(r["fcz"]))
    assert g["name"] == r["name"] == "test" and g["keys"] == r["keys"] and g["residues"] == r["residues"]
    for k in ("phi", "psi", "omega", "b_factors"):
        assert np.array_equal(np.float32(g[k]), np.float32(r[k])), k
    assert g["n_coords"] == r["n_coords"]
    assert np.abs(xyz_of(g["text"]) - xyz_of(r["text"])).max() <= 0.05 + 0.002
    # entries asked for by name ("chain03", "chain07"); what comes back is each blob's TITLE, which the reference CLI took
    # from the file's TITLE record (src/structure_reader.cpp:31-46)
    assert g["db_len"] == r["db_len"] == 10 and g["db_ids"] == r["db_ids"] == [batch.title(3), batch.title(7)]
    assert [n for n, _ in g["db"]] == [n for n, _ in r["db"]]
    for (n, tg), (_, tr) in zip(g["db"], r["db"]):
        assert tg.splitlines()[0] == tr.splitlines()[0], n
