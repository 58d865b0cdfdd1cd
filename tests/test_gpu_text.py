"""The CUDA text kernels (k_pdb_plan / k_pdb_emit / k_extract) through the C ABI against the oracle's restatement of
writeAtomCoordinatesToPDB (src/atom_coordinate.cpp:220-291) and Foldcomp::extract (src/foldcomp.cpp:1260-1336):
byte-identical, including over-long columns, continuation TITLE lines, the -a atom order and empty chains."""
import hashlib
import os

import numpy as np
import pytest

import helpers as H
from foldcomp_b200 import abi, synth
from foldcomp_b200.abi import HostBlobBatch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _check_text(engine, batch, use_alt=False):
    got = engine.pdb_text_host(batch)
    assert got.n_chains == batch.n_chains
    for c in range(batch.n_chains):
        want = H.oracle_format_pdb(batch, c, use_alt) if batch.res_off[c + 1] > batch.res_off[c] else b""
        assert got.text(c) == want, (c, len(got.text(c)), len(want))
    return got


def test_pdb_text_of_decoded_goldens_matches_committed_reference_output(engine, golden):
    """engine decode -> engine text: the coordinates differ from the reference's decode within tolerance, so the text
    is compared through the oracle on the SAME coordinates; the reference's own text of its own decode is pinned by
    sha256 in tests/golden/text_golden.npz (checked here by formatting the committed reference decode on the GPU)."""
    tg = np.load(os.path.join(HERE, "golden", "text_golden.npz"))
    blobs = golden.blobs(25)
    ref_dec = abi.concat_batches([H.decoded_as_batch(H.oracle_decode(b)) for b in blobs])
    got = engine.pdb_text_host(ref_dec)
    for c, name in enumerate(golden.names):
        assert hashlib.sha256(got.text(c)).hexdigest() == str(tg[f"pdb_sha256_{name}"]), name
    dec = engine.decode_host(HostBlobBatch.from_blobs(blobs))
    _check_text(engine, dec)
    engine.set_opts(use_alt_atom_order=True)
    try:
        dec_alt = engine.decode_host(HostBlobBatch.from_blobs(blobs))
        _check_text(engine, dec_alt, use_alt=True)
    finally:
        engine.set_opts(use_alt_atom_order=False)


def test_pdb_text_mixed_lengths_and_extreme_fields(engine):
    rng = np.random.default_rng(8)
    lens = synth.mixed_lengths(rng, 300, lo=2, hi=1500)
    lens[:8] = [2, 15, 16, 31, 32, 33, 64, 65]  # unit boundaries (32 residues per emit unit)
    parts = [synth.generate(len(lens), lens, seed=81)] + [H.extreme_text_chain(s) for s in range(3)] + [H.long_chain(9000)]
    batch = abi.concat_batches(parts)
    # titles of every length around the continuation boundary, and an empty one
    titles = [b"x" * ((7 * c) % 160) for c in range(batch.n_chains)]
    batch.titles = np.frombuffer(b"".join(titles), np.uint8).copy()
    batch.title_off = np.cumsum([0] + [len(t) for t in titles]).astype(np.uint32)
    _check_text(engine, batch)


def test_pdb_text_device_api_and_failed_chains(engine):
    import torch

    from foldcomp_b200.engine import DeviceBlobBatch, DeviceChainBatch, DeviceTextBatch

    batch = synth.generate(64, 120, seed=5)
    blobs = engine.encode_host(batch)
    bl = blobs.blobs()
    bl[3] = b"NOPE" + bl[3][4:]  # fails decode: empty chain, empty text
    hb = HostBlobBatch.from_blobs(bl)
    dev = torch.device("cuda:0")
    dblob = DeviceBlobBatch.from_host(hb, dev)
    dout = DeviceChainBatch(hb.n_chains, batch.n_res, batch.n_atoms, len(batch.titles), dev)
    torch.cuda.synchronize()
    engine.decode_plan_device(dblob, dout)
    engine.decode_device(dblob, dout)
    dtext = DeviceTextBatch(hb.n_chains, 16, dev)
    total = engine.pdb_text_plan_device(dout, dtext)
    dtext.bytes = torch.zeros(total, dtype=torch.uint8, device=dev)
    engine.pdb_text_device(dout, dtext)
    engine.sync()
    got, dec = dtext.to_host(), dout.to_host()
    assert dec.status[3] == abi.FCZ_E_MAGIC and got.text(3) == b""
    for c in range(hb.n_chains):
        if c != 3:
            assert got.text(c) == H.oracle_format_pdb(dec, c), c


def test_extract_plddt_and_sequence(engine, golden):
    blobs = golden.blobs(25) + list(golden.db_blobs) + [b"NOPE", b""]
    b01 = synth.generate(3, 77, seed=6)
    b01.bfactor = (b01.bfactor / np.float32(100.0)).astype(np.float32)  # 0..1 scale (src/foldcomp.cpp:1290-1296)
    blobs += H.oracle_encode_batch(b01, 25).blobs()
    hb = HostBlobBatch.from_blobs(blobs)
    for t, d in ((0, 1), (0, 2), (0, 3), (0, 4), (1, 0)):
        got = engine.extract_host(hb, t, d)
        for c, blob in enumerate(blobs):
            want = H.oracle_extract(blob, t, d) if len(blob) > 4 else b""
            assert got.text(c) == want, (c, t, d)


def test_decode_to_pdb_fused_and_python_decompress(engine, golden):
    """fcz_decode_to_pdb_*: blobs in, text out; equals format(decode) of the separate calls, bad blobs give a status and
    no text.  foldcomp_b200.decompress() -- the CPython module's decompress(bytes) -> (name, pdb) -- goes through it."""
    import foldcomp_b200

    blobs = golden.blobs(25) + [b"NOPE" + golden.blobs(25)[0][4:]] + list(golden.db_blobs[:5])
    hb = HostBlobBatch.from_blobs(blobs)
    got = engine.decode_to_pdb_host(hb)
    dec = engine.decode_host(hb)
    assert list(got.status) == list(dec.status) and got.status[len(golden.names)] == abi.FCZ_E_MAGIC
    for c in range(hb.n_chains):
        want = H.oracle_format_pdb(dec, c) if dec.status[c] == 0 else b""
        assert got.text(c) == want, c
    c = golden.names.index("test.pdb")
    name, pdb = foldcomp_b200.decompress(blobs[c])
    assert name == golden.batch.title(c) and pdb.encode("latin-1") == got.text(c)
    # against the reference's own text of its own decode: same lines, coordinates within the decode tolerance
    want = H.ref_decompress_to_pdb(blobs[c]).decode() if H.have_ref() else None
    if want is not None:
        gl, wl = pdb.splitlines(), want.splitlines()
        assert len(gl) == len(wl)
        for a, b in zip(gl, wl):
            assert a[:30] == b[:30] and a[54:] == b[54:]
            if a.startswith("ATOM"):
                assert max(abs(float(a[30 + 8 * k : 38 + 8 * k]) - float(b[30 + 8 * k : 38 + 8 * k])) for k in range(3)) <= 0.051


def test_text_entry_points_on_empty_batches(engine):
    from foldcomp_b200.abi import HostChainBatch

    assert engine.pdb_text_host(HostChainBatch.empty(0)).n_chains == 0
    assert engine.decode_to_pdb_host(HostBlobBatch.from_blobs([])).n_chains == 0
    assert engine.extract_host(HostBlobBatch.from_blobs([]), 1).n_chains == 0
    off, ang = engine.unpack_angles_host(HostBlobBatch.from_blobs([]))
    assert len(off) == 1 and ang.shape == (0, 6)


def test_extract_matches_the_references_committed_fixtures(engine):
    """The reference repository's own extract goldens (test/test_af.plddt, test/test_af.plddt.tsv) for its upstream-encoded
    test/test_af.fcz, committed verbatim in tests/golden/text_golden.npz."""
    tg = np.load(os.path.join(HERE, "golden", "text_golden.npz"))
    blob = bytes(tg["upstream_test_af_fcz"])
    digits1 = bytes(tg["upstream_test_af_plddt"]).split(b"\n")[1]
    _, n_res, digits4 = bytes(tg["upstream_test_af_plddt_tsv"]).rstrip(b"\n").split(b"\t")
    hb = HostBlobBatch.from_blobs([blob])
    assert engine.extract_host(hb, 0, 1).text(0) == digits1
    assert engine.extract_host(hb, 0, 4).text(0) == digits4
    assert len(engine.extract_host(hb, 1).text(0)) == int(n_res)
    # and the blob decodes (an upstream-encoded file, header floats differ in the last bit from a local encode)
    dec = engine.decode_host(hb)
    assert dec.status[0] == 0 and dec.n_res == int(n_res)
    want = H.oracle_decode(blob)
    assert np.array_equal(dec.res_type, want.res_type) and np.array_equal(dec.bfactor, want.bfactor)
    assert H.max_dev(dec.xyz, want.xyz) <= 0.05
