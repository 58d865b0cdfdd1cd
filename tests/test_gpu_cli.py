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
    subprocess.check_call([CLI, "compress", str(pdb_in), str(fcz)])
    # what the reference would produce for this text input (title = output basename, main.cpp:451-465)
    parsed = pdbio.parse_pdb_chain(pdb_in.read_text(), "out")
    want = H.oracle_encode(parsed, 0, 25)
    assert fcz.read_bytes() == want
    pdb_out = tmp_path / "back.pdb"
    subprocess.check_call([CLI, "decompress", str(fcz), str(pdb_out)])
    back = pdbio.parse_pdb_chain(pdb_out.read_text(), "out")
    ref = H.oracle_decode(want)
    assert np.array_equal(back.res_type, ref.res_type)
    assert np.abs(back.xyz - ref.xyz).max() <= 0.05 + 0.0006  # tolerance + 3-decimal text rounding
    assert H.rmsd(back.xyz, ref.xyz) <= 0.01 + 0.0006
    assert pdb_out.read_text().startswith("TITLE     out\nATOM ")


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
