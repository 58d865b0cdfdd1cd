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
    # shortest chains, ragged batch with one invalid residue code and one chain with more anchors than the
    # header's 8-bit nAnchor can hold (3100/10 + 2 = 312 > 255, src/foldcomp.h:118-131)
    lens = np.array([2, 3, 5, 40, 3100])
    batch = synth.generate(len(lens), lens, seed=9)
    r = int(batch.res_off[3]) + 7
    batch.res_type[r] = 20  # ASX has no table entry: the reference throws (AAS.at), we report FCZ_E_RESIDUE
    engine.set_opts(anchor_threshold=10)
    try:
        got = engine.encode_host(batch)
        assert list(got.status) == [0, 0, 0, abi.FCZ_E_RESIDUE, abi.FCZ_E_LIMIT]
        assert got.blob(3) == b"" and got.blob(4) == b""
    finally:
        engine.set_opts(anchor_threshold=25)
    got = engine.encode_host(batch)
    assert list(got.status) == [0, 0, 0, abi.FCZ_E_RESIDUE, 0]
    assert got.blob(3) == b"" and got.blob(4) == H.oracle_encode(batch, 4, 25)
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


def test_unknown_residue_codes_24_to_31_decode_as_unk(engine):
    """A record's 5-bit residue field can hold 24..31; the reference's switches map those to UNK
    (convertIntToOneLetterCode / convertIntToThreeLetterCode default branch, src/utility.cpp:297-377, 461-).  Such a
    blob must decode like the oracle (= the reference: backbone only, residue type UNK), through every path that
    indexes the residue tables with a code taken from the blob: decode (canonical and alt order), extract, the fused
    decode -> PDB text, get_data's angles.  Codes 20..22 (ASX, GLX, STP: no table entry, the reference throws) stay
    FCZ_E_RESIDUE."""
    blob = bytearray(H.oracle_encode(H.unk_chain(40, seed=9), 0, 25))
    o_rec = H.blob_sections(bytes(blob))[0]
    for r, code in ((0, 25), (3, 24), (7, 31), (20, 30), (39, 27)):
        blob[o_rec + 8 * r] = (blob[o_rec + 8 * r] & 7) | (code << 3)
    blob = bytes(blob)
    bad = bytearray(blob)
    bad[o_rec + 8 * 5] = (bad[o_rec + 8 * 5] & 7) | (21 << 3)
    good = H.oracle_encode(synth.generate(1, 33, seed=4), 0, 25)
    blobs = HostBlobBatch.from_blobs([good, blob, bytes(bad), good])
    for alt in (False, True):
        engine.set_opts(use_alt_atom_order=alt)
        try:
            dec = engine.decode_host(blobs)
            assert list(dec.status) == [0, 0, abi.FCZ_E_RESIDUE, 0]
            for c in (0, 1, 3):
                want = H.oracle_decode(blobs.blob(c), use_alt=alt)
                got = dec.chain(c)
                assert np.array_equal(got.res_type, want.res_type) and np.array_equal(got.bfactor, want.bfactor)
                assert got.xyz.shape == want.xyz.shape and H.max_dev(got.xyz, want.xyz) <= TOL_MAX
            assert set(dec.chain(1).res_type) == {23}
            txt = engine.decode_to_pdb_host(blobs)
            assert list(txt.status) == [0, 0, abi.FCZ_E_RESIDUE, 0] and txt.text(2) == b""
        finally:
            engine.set_opts(use_alt_atom_order=False)
    seq = engine.extract_host(blobs, 1)
    assert seq.text(1) == H.oracle_extract(blob, 1) == b"X" * 40
    res_off, ang = engine.unpack_angles_host(blobs)
    assert np.array_equal(ang[int(res_off[1]):int(res_off[2])], H.oracle_unpack_angles(blob))


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


@pytest.mark.parametrize("device_api", [False, True])
def test_long_chains_up_to_format_maximum(engine, device_api):
    """Chains beyond the shared-memory tiers (k_encode_long, the global-workspace decode kernels), up to the
    format's 16-bit residue count; mixed into one batch with short chains so every tier launches side by side."""
    import torch

    from foldcomp_b200.engine import DeviceBlobBatch, DeviceChainBatch

    for b, lens in ((25, (2721, 3400, 6200)), (200, (20000, 50000)), (300, (65535,))):
        parts = [H.long_chain(L, seed=i) for i, L in enumerate(lens)] + [synth.generate(12, np.array([2, 60, 350] * 4), seed=3)]
        batch = abi.concat_batches(parts)
        engine.set_opts(anchor_threshold=b)
        try:
            want = H.oracle_encode_batch(batch, b)
            if device_api:
                dev = torch.device("cuda:0")
                dbatch = DeviceChainBatch.from_host(batch, dev)
                dblob = DeviceBlobBatch(batch.n_chains, abi.encode_bound(batch.n_chains, batch.n_res, batch.n_atoms, len(batch.titles), b), dev)
                torch.cuda.synchronize()
                engine.encode_device(dbatch, dblob)
                engine.sync()
                got = dblob.to_host()
                dout = DeviceChainBatch(batch.n_chains, batch.n_res, batch.n_atoms, len(batch.titles), dev)
                engine.decode_plan_device(dblob, dout)
                engine.decode_device(dblob, dout)
                engine.sync()
                dec = dout.to_host()
            else:
                got = engine.encode_host(batch)
                dec = engine.decode_host(HostBlobBatch(want.blob_off, want.bytes))
            assert not got.status.any(), (b, list(got.status))
            assert np.array_equal(got.blob_off, want.blob_off)
            bad = [c for c in range(batch.n_chains) if got.blob(c) != want.blob(c)]
            assert not bad, (b, bad)
            _assert_decoded_close(dec, H.oracle_decode_batch(want))
        finally:
            engine.set_opts(anchor_threshold=25)


def test_degenerate_inputs(engine):
    """Constant B-factors (the reference's discretiser divides by zero and casts NaN), coincident atoms (NaN cosines),
    missing atoms at (0,0,0), collinear triples, NaN / inf coordinates, integer-lattice coordinates, two-valued
    B-factors, and NaN / inf in the FIRST element of every array (whose NaN reaches the header floats, payload and
    sign as an x86-64 build of the reference leaves them): byte-identical to the oracle for EVERY chain (the oracle is
    pinned to the unmodified reference on the same inputs by test_codec_model.py)."""
    kinds, parts = zip(*H.degenerate_chains())
    big = abi.concat_batches(list(parts))
    try:
        for b in (25, 10):
            engine.set_opts(anchor_threshold=b)
            got = engine.encode_host(big)
            want = H.oracle_encode_batch(big, b)
            assert not got.status.any()
            assert np.array_equal(got.blob_off, want.blob_off)
            bad = [(c, kinds[c]) for c in range(big.n_chains) if got.blob(c) != want.blob(c)]
            assert not bad, (b, bad)
    finally:
        engine.set_opts(anchor_threshold=25)


def test_degenerate_decode(engine):
    """Decode of the blobs encoded from the degenerate chains, against the oracle's decode.  Exact fields (residue
    types, B-factors, titles, metadata) always agree.  Coordinates: where the INPUT has no exactly collinear
    consecutive backbone triple, the NaN masks are identical and every finite atom is within the tolerance.  With a
    collinear triple the reference's own frame normalisation (src/nerf.cpp:52-85) is 0/0 or amplifies the last-bit
    noise of its sincosf by ~1e7 -- its output there is rounding noise (all-NaN or Angstroms away from the input) that
    no implementation without glibc's sincosf reproduces; those chains only have to decode without a fault."""
    kinds, parts = zip(*H.degenerate_chains())
    big = abi.concat_batches(list(parts))
    try:
        for b in (25, 10):
            engine.set_opts(anchor_threshold=b)
            blobs = H.oracle_encode_batch(big, b)
            dec = engine.decode_host(HostBlobBatch(blobs.blob_off, blobs.bytes))
            ref = H.oracle_decode_batch(blobs)
            assert not dec.status.any()
            assert np.array_equal(dec.res_off, ref.res_off) and np.array_equal(dec.atom_off, ref.atom_off)
            assert np.array_equal(dec.res_type, ref.res_type) and np.array_equal(dec.titles, ref.titles)
            nb = np.isnan(ref.bfactor)  # NaN B-factors (NaN header floats) at the same places, everything else bit for bit
            assert np.array_equal(np.isnan(dec.bfactor), nb) and np.array_equal(dec.bfactor[~nb].view(np.uint32), ref.bfactor[~nb].view(np.uint32))
            assert dec.meta.tobytes() == ref.meta.tobytes()
            bad, n_ill = [], 0
            for c in range(big.n_chains):
                a0, a1 = int(ref.atom_off[c]), int(ref.atom_off[c + 1])
                if H.has_collinear_backbone(parts[c]):
                    n_ill += 1
                    continue
                why = H.decode_mismatch(dec.xyz[a0:a1], ref.xyz[a0:a1], ref.res_type[int(ref.res_off[c]):int(ref.res_off[c + 1])],
                                        TOL_BB_RMSD, TOL_MAX)
                if why:
                    bad.append((c, kinds[c], why))
            assert not bad, (b, bad[:10])
            assert n_ill < big.n_chains // 4
    finally:
        engine.set_opts(anchor_threshold=25)


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


# ------------------------------------------------------------------------------ full-size config 3


def test_config3_542k_chain_db_decode(engine):
    """BASELINE.json configs[2]: a 542 378-chain (afdb_swissprot_v4-scale) FCZ database decoded in ONE call on one GPU.
    Every chain is DISTINCT: 542 378 chains of 48 lengths between 50 and 2000 residues (config 5's clipped log-normal
    distribution, lengths snapped to a geometric grid), generated on the device (foldcomp_b200/synth_device.py) and
    encoded by the engine.  Checks: every chain encodes and decodes with status 0 and the planned sizes; EVERY chain's
    round trip against its own input, on the device (pooled and worst-chain all-atom RMSD); the oracle on every 2000th
    chain -- FCZ bytes identical, decoded coordinates within the BASELINE.md tolerances."""
    import torch

    from foldcomp_b200 import synth_device
    from foldcomp_b200.engine import DeviceBlobBatch, DeviceChainBatch

    N_DB = 542378
    engine.set_opts(anchor_threshold=25)
    grid = np.unique(np.geomspace(50, 2000, 48).astype(np.int64))
    lens = synth.mixed_lengths(np.random.default_rng(3), N_DB)
    snapped = grid[np.abs(np.log(lens[:, None] / grid[None, :])).argmin(1)]
    lengths, counts = np.unique(snapped, return_counts=True)
    dev = torch.device("cuda:0")
    g = synth_device.generate_device_mixed(lengths, counts, 303, dev)
    n = N_DB
    d = DeviceChainBatch(n, 1, 1, 1, dev)
    d.res_off, d.atom_off, d.title_off = g["res_off"], g["atom_off"], g["title_off"]
    d.res_type, d.bfactor, d.xyz, d.titles, d.meta = g["res_type"], g["bfactor"], g["xyz"], g["titles"], g["meta"]
    d.status = torch.zeros(n, dtype=torch.int32, device=dev)
    n_res, n_atoms, n_title = int(d.res_off[-1].item()), int(d.atom_off[-1].item()), int(d.title_off[-1].item())
    d.n_res, d.n_atoms, d.n_title = n_res, n_atoms, n_title
    assert n_res > 150_000_000
    dblob = DeviceBlobBatch(n, abi.encode_bound(n, n_res, n_atoms, n_title, 25), dev)
    engine.encode_device(d, dblob)
    engine.sync()
    assert int(dblob.status.count_nonzero().item()) == 0
    dout = DeviceChainBatch(n, n_res, n_atoms, n_title, dev)
    torch.cuda.synchronize()
    sizes = engine.decode_plan_device(dblob, dout)
    assert (sizes.n_res, sizes.n_atoms, sizes.n_title_bytes) == (n_res, n_atoms, n_title)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    engine.sync()
    s = torch.cuda.current_stream()
    ev0.record(s)
    torch.cuda.synchronize()
    engine.decode_device(dblob, dout)
    engine.sync()
    ev1.record(s)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    assert int(dout.status.count_nonzero().item()) == 0
    assert torch.equal(dout.res_type[:n_res], d.res_type) and torch.equal(dout.atom_off.to(torch.int64), d.atom_off.to(torch.int64))
    assert torch.equal(dout.titles[:n_title], d.titles) and torch.equal(dout.meta[:n], d.meta)
    # every chain against its own input, on the device (the coordinates are ~19 GB each way)
    apc = (d.atom_off[1:] - d.atom_off[:-1]).to(torch.int64)
    per_chain = torch.zeros(n, device=dev, dtype=torch.float64)
    pooled, CH = 0.0, 20_000_000
    chain_of_atom = torch.repeat_interleave(torch.arange(n, device=dev), apc)
    for a0 in range(0, n_atoms, CH):
        a1 = min(n_atoms, a0 + CH)
        d2 = ((dout.xyz[a0:a1].double() - d.xyz[a0:a1].double()) ** 2).sum(1)
        pooled += float(d2.sum().item())
        per_chain.index_add_(0, chain_of_atom[a0:a1], d2)
    rt_all = (pooled / n_atoms) ** 0.5
    worst = float(torch.sqrt(per_chain / apc.double()).max().item())
    assert rt_all <= 0.1 and worst <= 0.5, (rt_all, worst)
    # the oracle on every 2000th chain
    sample = list(range(0, n, 2000))
    h_res_off, h_atom_off, h_title_off = d.res_off.cpu().numpy(), d.atom_off.cpu().numpy(), d.title_off.cpu().numpy()
    h_boff = dblob.blob_off.cpu().numpy().view(np.uint64).astype(np.int64)
    bb_max, dev_max = 0.0, 0.0
    for c in sample:
        r0, r1, a0, a1 = int(h_res_off[c]), int(h_res_off[c + 1]), int(h_atom_off[c]), int(h_atom_off[c + 1])
        t0, t1 = int(h_title_off[c]), int(h_title_off[c + 1])
        one = abi.HostChainBatch(
            res_off=np.array([0, r1 - r0], np.uint32), atom_off=np.array([0, a1 - a0], np.uint64), title_off=np.array([0, t1 - t0], np.uint32),
            res_type=d.res_type[r0:r1].cpu().numpy(), bfactor=d.bfactor[r0:r1].cpu().numpy(), xyz=d.xyz[a0:a1].cpu().numpy(),
            titles=d.titles[t0:t1].cpu().numpy(), meta=d.meta[c : c + 1].cpu().numpy().view(abi.META_DTYPE).reshape(-1))
        want = H.oracle_encode(one, 0, 25)
        assert bytes(dblob.bytes[int(h_boff[c]) : int(h_boff[c + 1])].cpu().numpy()) == want, c
        ref = H.oracle_decode(want)
        mine = dout.xyz[a0:a1].cpu().numpy()
        bbm = H.backbone_mask(ref.res_type)
        bb_max, dev_max = max(bb_max, H.rmsd(mine[bbm], ref.xyz[bbm])), max(dev_max, H.max_dev(mine, ref.xyz))
    assert bb_max <= TOL_BB_RMSD and dev_max <= TOL_MAX, (bb_max, dev_max)
    print(f"config3: {N_DB} distinct chains of {len(lengths)} lengths, {n_res} residues decoded in {ms:.1f} ms wall "
          f"({n_res / ms / 1e6:.2f} G res/s incl. launch overheads); all-atom round-trip RMSD vs input {rt_all:.3f} A (worst chain {worst:.3f}); "
          f"oracle on {len(sample)} chains: bytes identical, decode bb RMSD <= {bb_max:.1e}, max {dev_max:.1e}")


def test_terminated_blobs_are_a_db_slab(engine):
    """opts.terminate_blobs: every blob is followed by one NUL that blob_off counts, so the output is a foldcomp-db
    data slab as it stands; decode accepts the terminated entries unchanged."""
    lens = np.array([2, 17, 64, 65, 350, 351, 1300, 3000])
    batch = synth.generate(len(lens), lens, seed=88)
    want = H.oracle_encode_batch(batch, 25)
    engine.set_opts(terminate_blobs=True)
    try:
        got = engine.encode_host(batch)
    finally:
        engine.set_opts(terminate_blobs=False)
    assert not got.status.any()
    assert np.array_equal(np.diff(got.blob_off.astype(np.int64)), np.diff(want.blob_off.astype(np.int64)) + 1)
    for c in range(batch.n_chains):
        assert got.blob(c) == want.blob(c) + b"\0", c
    dec = engine.decode_host(HostBlobBatch(got.blob_off, got.bytes))
    _assert_decoded_close(dec, H.oracle_decode_batch(want))
