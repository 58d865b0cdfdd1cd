"""The C++ host adapter (foldcomp_b200/csrc/foldcomp_gpu.{h,cpp}: class FoldcompGpu mirroring the
reference's class Foldcomp) driven through the minimal CLI fcz_cli: PDB text -> .fcz -> PDB text."""
import os
import subprocess

import numpy as np
import pytest

import helpers as H
from foldcomp_b200 import pdbio

pytestmark = pytest.mark.gpu
CLI = os.path.join(H.ROOT, "foldcomp_b200", "csrc", "fcz_cli")


@pytest.mark.parametrize("name", ["test.pdb", "test_af.pdb"])
def test_cli_roundtrip_matches_oracle(golden, tmp_path, name):
    c = golden.names.index(name)
    ch = golden.batch.chain(c)
    pdb_in = tmp_path / "in.pdb"
    pdb_in.write_text(pdbio.format_pdb(ch, 0))
    fcz = tmp_path / "out.fcz"
    subprocess.check_call([CLI, "compress", "in.pdb", "out.fcz"], cwd=tmp_path)
    # what the reference CLI would produce for this text input: the title is the text's TITLE record; a text without HEADER id /
    # TITLE record is named after the output path as given, without its extension (main.cpp:451-467, getFileParts)
    title = ch.title(0) or "out"
    parsed = pdbio.parse_pdb_chain(pdb_in.read_text(), title)
    want = H.oracle_encode(parsed, 0, 25)
    assert fcz.read_bytes() == want
    pdb_out = tmp_path / "back.pdb"
    subprocess.check_call([CLI, "decompress", str(fcz), str(pdb_out)])
    back = pdbio.parse_pdb_chain(pdb_out.read_text(), title)
    ref = H.oracle_decode(want)
    assert np.array_equal(back.res_type, ref.res_type)
    assert np.abs(back.xyz - ref.xyz).max() <= 0.05 + 0.0006  # tolerance + 3-decimal text rounding
    assert H.rmsd(back.xyz, ref.xyz) <= 0.01 + 0.0006
    assert pdb_out.read_text().startswith(f"TITLE     {title}\nATOM ")


def test_cli_rejects_garbage(tmp_path):
    bad = tmp_path / "bad.fcz"
    bad.write_bytes(b"not an fcz file")
    r = subprocess.run([CLI, "decompress", str(bad), str(tmp_path / "x.pdb")], capture_output=True)
    assert r.returncode == 1 and b"not an FCZ" in r.stderr


def test_cli_tar_extract_check(golden, tmp_path):
    """FoldcompGpu::writeTar / extract / checkValidity (mirrors of src/foldcomp.h:382, 394, 401) through fcz_cli: the
    archive is a valid tar whose member is the oracle's blob; extract and check agree with the oracle."""
    import tarfile

    ch = golden.batch.chain(golden.names.index("test_af.pdb"))
    pdb_in = tmp_path / "in.pdb"
    pdb_in.write_text(pdbio.format_pdb(ch, 0))
    subprocess.check_call([CLI, "compress-tar", str(pdb_in), str(tmp_path / "out.tar")])
    want = H.oracle_encode(pdbio.parse_pdb_chain(pdb_in.read_text(), "in"), 0, 25)
    with tarfile.open(tmp_path / "out.tar") as t:
        members = t.getmembers()
        assert [m.name for m in members] == ["in.fcz"] and members[0].size == len(want)
        assert t.extractfile(members[0]).read() == want
    (tmp_path / "in.fcz").write_bytes(want)
    for mode, type_ in (("extract-plddt", 0), ("extract-fasta", 1)):
        subprocess.check_call([CLI, mode, str(tmp_path / "in.fcz"), str(tmp_path / "x.txt")])
        lines = (tmp_path / "x.txt").read_bytes().split(b"\n")
        assert lines[1] == H.oracle_extract(want, type_, 1)
    subprocess.check_call([CLI, "check", str(tmp_path / "in.fcz"), str(tmp_path / "c.txt")])
    assert int((tmp_path / "c.txt").read_text()) == H.oracle_check(want)[1] == 0


def test_cli_compress_like_the_reference_cli_on_chains_fragments_and_titles(tmp_path):
    """`fcz_cli compress` on one file against the reference's own CLI (integration/_build/foldcomp_ref): a TITLE record
    becomes the title, several chains and numbering gaps give one .fcz per chain and fragment with the reference's file
    names (src/main.cpp:466-509)."""
    from foldcomp_b200 import synth

    ref_cli = os.path.join(H.ROOT, "integration", "_build", "foldcomp_ref")
    if not os.path.exists(ref_cli):
        pytest.skip("integration/_build/foldcomp_ref not built")
    batch = synth.generate(3, [40, 55, 30], seed=17)
    t = [pdbio.format_pdb(batch.chain(c), 0) for c in range(3)]
    atoms = lambda c, ch, shift=0, cut=None: [
        l[:21] + ch + "%4d" % (int(l[22:26]) + (shift if cut is not None and int(l[22:26]) > cut else 0)) + l[26:]
        for l in t[c].splitlines() if l.startswith("ATOM")]
    cases = {
        "titled": t[0],                                                                      # TITLE record present
        "plain": "\n".join(atoms(0, "A")) + "\nEND\n",                                      # no TITLE: named after the output
        "multi": "TITLE     TWO CHAINS AND A GAP\n" + "\n".join(atoms(1, "A", 6, 20) + atoms(2, "B")) + "\nEND\n",
    }
    for name, text in cases.items():
        (tmp_path / f"{name}.pdb").write_text(text)
        for tag, cli in (("ours", CLI), ("ref", ref_cli)):
            d = tmp_path / f"{tag}_{name}"
            d.mkdir()
            args = [cli, "compress"] + (["-y"] if tag == "ref" else []) + [f"../{name}.pdb", "o.fcz"]
            r = subprocess.run(args, cwd=d, capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, (tag, name, r.stdout[-300:], r.stderr[-300:])
        ours = {f: H.masked((tmp_path / f"ours_{name}" / f).read_bytes()) for f in sorted(os.listdir(tmp_path / f"ours_{name}"))}
        ref = {f: H.masked((tmp_path / f"ref_{name}" / f).read_bytes()) for f in sorted(os.listdir(tmp_path / f"ref_{name}"))}
        assert ours == ref, (name, sorted(ours), sorted(ref))
    assert len(os.listdir(tmp_path / "ours_multi")) == 3  # oA_0.fcz, oA_1.fcz, oB.fcz
