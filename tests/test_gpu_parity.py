"""Parity of the CUDA engine (through the C ABI, include/fcz_engine.h) with the oracle and with the
committed reference outputs.  Encode: byte-identical.  Decode: BASELINE.md section 3 tolerances --
backbone (N,CA,C) RMSD <= 0.01 A, max per-atom deviation <= 0.05 A, all-atom RMSD <= 0.02 A versus
the reference's own decode of the same blob (measured: ~1e-4 A); residue types, B-factors, titles and
per-chain metadata exact.
"""
import numpy as np
import pytest

import helpers as H
from foldcomp_b200 import abi, synth
from foldcomp_b200.abi import HostBlobBatch, HostChainBatch

pytestmark = pytest.mark.gpu

TOL_BB_RMSD, TOL_MAX, TOL_ALL_RMSD = 0.01, 0.05, 0.02


def _assert_decoded_close(got: HostChainBatch, want: HostChainBatch, exact_meta=True):
    assert np.array_equal(got.res_off, want.res_off)
    assert np.array_equal(got.atom_off, want.atom_off)
    assert np.array_equal(got.res_type, want.res_type)
    assert np.array_equal(got.bfactor, want.bfactor)
    if exact_meta:
        assert np.array_equal(got.title_off, want.title_off)
        assert np.array_equal(got.titles, want.titles)
        assert got.meta.tobytes() == want.meta.tobytes()
    bb, allr, mx = H.per_chain_deviation(got, want)
    assert bb <= TOL_BB_RMSD and allr <= TOL_ALL_RMSD and mx <= TOL_MAX, (bb, allr, mx)
    return bb, allr, mx


# ------------------------------------------------------------------------------ reference goldens


@pytest.mark.parametrize("b", [25, 10, 50, 200])
def test_encode_golden_bytes(engine, golden, b):
    engine.set_opts(anchor_threshold=b)
    out = engine.encode_host(golden.batch)
    want = golden.blobs(b)
    assert list(out.status) == [0] * golden.batch.n_chains
    for c, name in enumerate(golden.names):
        assert out.blob(c) == want[c], (name, b)


@pytest.mark.parametrize("b", [25, 10, 50, 200])
def test_decode_golden_blobs(engine, golden, b):
    blobs = HostBlobBatch.from_blobs(golden.blobs(b))
    got = engine.decode_host(blobs)
    assert list(got.status) == [0] * blobs.n_chains
    for c, name in enumerate(golden.names):
        xyz, bf = golden.decoded(b, c)
        ch = got.chain(c)
        bbm = H.backbone_mask(ch.res_type)
        assert np.array_equal(ch.bfactor, bf), name
        assert H.rmsd(ch.xyz[bbm], xyz[bbm]) <= TOL_BB_RMSD, name
        assert H.max_dev(ch.xyz, xyz) <= TOL_MAX, name
        assert H.rmsd(ch.xyz, xyz) <= TOL_ALL_RMSD, name
        assert ch.title(0) == golden.batch.title(c)
        assert ch.meta.tobytes() == golden.batch.meta[c : c + 1].tobytes()


def test_decode_upstream_example_db(engine, golden):
    got = engine.decode_host(HostBlobBatch.from_blobs(golden.db_blobs))
    assert list(got.status) == [0] * len(golden.db_blobs)
    assert H.max_dev(got.xyz, golden.db_xyz) <= TOL_MAX
    assert H.rmsd(got.xyz, golden.db_xyz) <= TOL_BB_RMSD


def test_decode_alt_atom_order(engine, golden):
    c = golden.names.index("test.pdb")
    blob = golden.blobs(25)[c]
    engine.set_opts(use_alt_atom_order=True)
    try:
        got = engine.decode_host(HostBlobBatch.from_blobs([blob]))
    finally:
        engine.set_opts(use_alt_atom_order=False)
    want = H.oracle_decode(blob, use_alt=True)
    assert H.max_dev(got.xyz, want.xyz) <= TOL_MAX


# ------------------------------------------------------------------------------ oracle, synthetic


def test_roundtrip_uniform_350_host_api(engine):
    engine.set_opts(anchor_threshold=25)
    batch = synth.generate(512, 350, seed=1234)
    want = H.oracle_encode_batch(batch, 25)
    got = engine.encode_host(batch)
    assert np.array_equal(got.blob_off, want.blob_off)
    assert np.array_equal(got.bytes[: len(want.bytes)], want.bytes)
    dec = engine.decode_host(HostBlobBatch(want.blob_off, want.bytes))
    _assert_decoded_close(dec, H.oracle_decode_batch(want))


def test_roundtrip_uniform_350_device_api(engine):
    import torch

    from foldcomp_b200.engine import DeviceBlobBatch, DeviceChainBatch

    engine.set_opts(anchor_threshold=25)
    batch = synth.generate(300, 350, seed=77)
    want = H.oracle_encode_batch(batch, 25)
    dev = torch.device("cuda:0")
    dbatch = DeviceChainBatch.from_host(batch, dev)
    dblob = DeviceBlobBatch(batch.n_chains, abi.encode_bound(batch.n_chains, batch.n_res, batch.n_atoms, len(batch.titles), 25), dev)
    torch.cuda.synchronize()
    engine.encode_device(dbatch, dblob)
    engine.sync()
    got = dblob.to_host()
    assert np.array_equal(got.blob_off, want.blob_off)
    assert np.array_equal(got.bytes, want.bytes)
    dout = DeviceChainBatch(batch.n_chains, batch.n_res, batch.n_atoms, len(batch.titles), dev)
    sizes = engine.decode_plan_device(dblob, dout)
    assert (sizes.n_res, sizes.n_atoms, sizes.n_title_bytes) == (batch.n_res, batch.n_atoms, len(batch.titles))
    engine.decode_device(dblob, dout)
    engine.sync()
    _assert_decoded_close(dout.to_host(), H.oracle_decode_batch(want))


@pytest.mark.parametrize("b", [10, 25, 50, 200])
def test_mixed_lengths_anchor_sweep(engine, b):
    """BASELINE.json config 5 shape: 50-2000 residues, -b sweep.  Exercises every tier incl. the large one."""
    rng = np.random.default_rng(5 + b)
    lens = synth.mixed_lengths(rng, 160)
    lens[:10] = [50, 64, 65, 128, 129, 384, 385, 1280, 1281, 2000]
    batch = synth.generate(len(lens), lens, seed=1000 + b)
    engine.set_opts(anchor_threshold=b)
    try:
        want = H.oracle_encode_batch(batch, b)
        got = engine.encode_host(batch)
        assert list(got.status) == [0] * batch.n_chains
        assert np.array_equal(got.blob_off, want.blob_off)
        bad = [c for c in range(batch.n_chains) if got.blob(c) != want.blob(c)]
        assert not bad, (bad[:5], [int(lens[c]) for c in bad[:5]])
        dec = engine.decode_host(HostBlobBatch(want.blob_off, want.bytes))
        _assert_decoded_close(dec, H.oracle_decode_batch(want))
    finally:
        engine.set_opts(anchor_threshold=25)


def test_edge_cases(engine):
    engine.set_opts(anchor_threshold=25)
    # empty batch
    e = HostChainBatch.empty(0)
    out = engine.encode_host(e)
    assert out.n_chains == 0 and int(out.blob_off[0]) == 0
    dec = engine.decode_host(HostBlobBatch.from_blobs([]))
    assert dec.n_chains == 0
    # shortest chains, ragged batch with one invalid residue code and one chain that is too long
    lens = np.array([2, 3, 5, 40, 3100])
    batch = synth.generate(len(lens), lens, seed=9)
    r = int(batch.res_off[3]) + 7
    batch.res_type[r] = 20  # ASX has no table entry: the reference throws (AAS.at), we report FCZ_E_RESIDUE
    got = engine.encode_host(batch)
    assert list(got.status) == [0, 0, 0, abi.FCZ_E_RESIDUE, abi.FCZ_E_LIMIT]
    assert got.blob(3) == b"" and got.blob(4) == b""
    for c in range(3):
        assert got.blob(c) == H.oracle_encode(batch, c, 25)
    # decode: bad magic, truncated blob, good blob in one batch
    good = got.blob(2)
    blobs = HostBlobBatch.from_blobs([b"NOPE" + good[4:], good[: len(good) - 3], good, b""])
    dec = engine.decode_host(blobs)
    assert list(dec.status) == [abi.FCZ_E_MAGIC, abi.FCZ_E_TRUNCATED, 0, abi.FCZ_E_MAGIC]
    assert dec.n_res == 5
    want = H.oracle_decode(good)
    assert H.max_dev(dec.chain(2).xyz, want.xyz) <= TOL_MAX


def test_unk_residue_and_missing_atoms(engine):
    """UNK carries only N,CA,C; an atom given as (0,0,0) is encoded like the reference's missing atom."""
    batch = synth.generate(1, 30, seed=21)
    ch = batch.chain(0)
    # replace residue 4 by UNK: drop its side-chain atoms
    from foldcomp_b200.tables import tables

    tb = tables()
    nat = tb.natoms[ch.res_type]
    starts = np.concatenate([[0], np.cumsum(nat)])
    keep = np.ones(len(ch.xyz), bool)
    keep[starts[4] + 3 : starts[5]] = False
    rt = ch.res_type.copy()
    rt[4] = 23
    xyz = ch.xyz[keep].copy()
    r_cb = next(r for r in range(6, 30) if nat[r] >= 5)
    xyz[starts[r_cb] + 4 - (nat[4] - 3)] = 0.0  # a "missing" CB further down the chain
    b2 = abi.concat_chains([(rt, ch.bfactor, xyz, ch.titles, ch.meta)])
    got = engine.encode_host(b2)
    assert got.status[0] == 0
    assert got.blob(0) == H.oracle_encode(b2, 0, 25)
    dec = engine.decode_host(got)
    want = H.oracle_decode(got.blob(0))
    assert np.array_equal(dec.res_type, want.res_type) and dec.n_atoms == len(want.xyz)
    assert H.max_dev(dec.xyz, want.xyz) <= TOL_MAX


def test_long_titles_and_unaligned_offsets(engine):
    """Titles of every length 0..40 shift all section and blob offsets through every 16-byte phase."""
    n = 41
    batch = synth.generate(n, 60, seed=31)
    titles = [b"t" * i for i in range(n)]
    batch.titles = np.frombuffer(b"".join(titles), np.uint8).copy()
    batch.title_off = np.cumsum([0] + [len(t) for t in titles]).astype(np.uint32)
    got = engine.encode_host(batch)
    want = H.oracle_encode_batch(batch, 25)
    assert np.array_equal(got.blob_off, want.blob_off)
    assert np.array_equal(got.bytes[: len(want.bytes)], want.bytes)
    dec = engine.decode_host(HostBlobBatch(want.blob_off, want.bytes))
    _assert_decoded_close(dec, H.oracle_decode_batch(want))


# ------------------------------------------------------------------------------ full-size config 2


def test_config2_10k_chains_350(engine):
    """BASELINE.json configs[1]: 10k synthetic 350-residue chains, FCZ byte-identical to the oracle for
    EVERY chain; decode checked against the oracle on every chain as well, plus size-independent
    round-trip properties."""
    engine.set_opts(anchor_threshold=25)
    batch = synth.generate(10000, 350, seed=synth.SEED)
    want = H.oracle_encode_batch(batch, 25)
    got = engine.encode_host(batch)
    assert not got.status.any()
    assert np.array_equal(got.blob_off, want.blob_off)
    if not np.array_equal(got.bytes[: len(want.bytes)], want.bytes):
        bad = [c for c in range(batch.n_chains) if got.blob(c) != want.blob(c)]
        raise AssertionError(f"{len(bad)} of {batch.n_chains} blobs differ, first {bad[:5]}")
    dec = engine.decode_host(HostBlobBatch(got.blob_off, got.bytes))
    bb, allr, mx = _assert_decoded_close(dec, H.oracle_decode_batch(want))
    # properties: sizes follow the format's formula; residue types and titles survive; the round trip
    # stays within the reference's own loss (backbone RMSD vs the ORIGINAL coordinates)
    L, A = 350, np.diff(batch.atom_off.astype(np.int64))
    assert np.array_equal(np.diff(got.blob_off.astype(np.int64)), 97 + 40 * 16 + 11 + 8 * L + (A - 3 * L) + L)
    assert np.array_equal(dec.res_type, batch.res_type) and np.array_equal(dec.titles, batch.titles)
    rb, ra, _ = H.per_chain_deviation(dec, HostChainBatch(batch.res_off, batch.atom_off, batch.title_off, batch.res_type,
                                                          dec.bfactor, batch.xyz, batch.titles, batch.meta))
    assert rb <= 0.1, rb  # north-star ceiling; typical 0.03-0.05 at -b 25
    print(f"config2: decode vs oracle bb_rmsd={bb:.2e} all={allr:.2e} max={mx:.2e}; round trip bb={rb:.3f} all={ra:.3f}")
