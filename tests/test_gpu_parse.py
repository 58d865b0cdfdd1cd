"""SURVEY.md section 8 f3 on the GPU: the CUDA ATOM parser (k_parse_lines / k_parse_plan / k_parse_emit) through the C ABI
(fcz_parse_pdb_plan + fcz_parse_pdb_batch on device text, fcz_encode_pdb_text_batch on host text) against the host parser
parsePdbChain (pinned to the reference's CPython module by tests/test_db_host.py), the oracle's encoder, and -- when
oracle/_ref/pyref is there -- the reference's own compress() (foldcomp/foldcomp.cxx:253-293) on the same text."""
import numpy as np
import pytest

import dbutil
import helpers as H
from foldcomp_b200 import abi, pdbio, synth
from foldcomp_b200.abi import HostTextBatch
from foldcomp_b200.engine import DeviceTextBatch
from test_parse import _host_parse, cli_fragments, messy_variants

pytestmark = pytest.mark.gpu

FLAG_TO_STATUS = {0: 0, 1: abi.FCZ_E_PARSE_NOATOM, 2: abi.FCZ_E_PARSE_CHAINS, 3: abi.FCZ_E_PARSE_RECORD, 4: abi.FCZ_E_PARSE_NUMBER,
                  5: abi.FCZ_E_PARSE_GAPS}


def _want_flag(text: bytes, host_flag: int) -> int:
    """The GPU parser's flag from the single-chain host parser's: the same, plus 5 where `foldcomp compress` would cut the
    chain into fragments (the host side of that is parsePdbUnits)."""
    return 5 if host_flag == 0 and cli_fragments(text) else host_flag


def _texts_batch(texts):
    off = np.zeros(len(texts) + 1, np.uint64)
    off[1:] = np.cumsum([len(t) for t in texts], dtype=np.uint64)
    return HostTextBatch(off, np.frombuffer(b"".join(texts) + b"\0", np.uint8).copy())


def _device_texts(texts, device):
    import torch

    hb = _texts_batch(texts)
    d = DeviceTextBatch(hb.n_chains, len(hb.bytes), device)
    d.text_off.copy_(torch.from_numpy(hb.text_off.view(np.int64)))
    d.bytes[: len(hb.bytes)].copy_(torch.from_numpy(hb.bytes))
    return d


def _texts(golden):
    batch = synth.generate(40, synth.mixed_lengths(np.random.default_rng(12), 40, 2, 900), seed=77)
    texts = [pdbio.format_pdb(batch.chain(c), 0).encode() for c in range(batch.n_chains)]
    names = [f"syn{c}" for c in range(batch.n_chains)]
    for k, v in messy_variants(golden).items():
        texts.append(v.encode())
        names.append(k)
    texts.append(b"ATOM      1  N   MET A   1      1e1     2.000   3.000  1.00 50.00           N  \n" * 3)  # exponent: outside the fixed-point grammar
    names.append("exponent")
    return texts, names


def test_device_parser_matches_host_parser(engine, golden):
    import torch

    texts, names = _texts(golden)
    dt = _device_texts(texts, torch.device("cuda", engine.device))
    got = engine.parse_pdb_device(dt)
    engine.sync()
    hb = got.to_host()
    status = got.status[: len(texts)].cpu().numpy()
    for c, (t, name) in enumerate(zip(texts, names)):
        want = _host_parse(t)
        if name == "exponent":  # strtof reads it, the GPU grammar rejects it -- never mis-parses it
            assert status[c] == abi.FCZ_E_PARSE_NUMBER and want[0] == 0
            continue
        flag = _want_flag(t, want[0])
        assert status[c] == FLAG_TO_STATUS[flag], (name, status[c], flag)
        r0, r1, a0, a1 = int(hb.res_off[c]), int(hb.res_off[c + 1]), int(hb.atom_off[c]), int(hb.atom_off[c + 1])
        if flag:
            assert r1 == r0 and a1 == a0, name
            continue
        assert np.array_equal(hb.res_type[r0:r1], want[1]), name
        assert np.array_equal(hb.bfactor[r0:r1].view(np.uint32), want[2].view(np.uint32)), name
        assert np.array_equal(hb.xyz[a0:a1].view(np.uint32), want[3].view(np.uint32)), name
        assert hb.meta[c].tobytes() == want[4].tobytes(), name


def test_text_to_fcz_in_one_call_matches_oracle_and_reference(engine, golden):
    texts, names = _texts(golden)
    blobs = engine.encode_pdb_text_host(_texts_batch(texts), [n.encode() for n in names])
    ref = dbutil.reference_module()
    n_ref = 0
    for c, (t, name) in enumerate(zip(texts, names)):
        flag, rt, bf, xyz, meta = _host_parse(t)
        if name == "exponent":
            assert blobs.status[c] == abi.FCZ_E_PARSE_NUMBER and blobs.blob(c) == b""
            continue
        flag = _want_flag(t, flag)
        if flag:
            assert blobs.status[c] == FLAG_TO_STATUS[flag] and blobs.blob(c) == b"", name
            continue
        one = abi.concat_chains([(rt, bf, xyz, np.frombuffer(name.encode(), np.uint8), np.array([meta]))])
        if len(rt) < 2:  # the encoder's own limit
            assert blobs.status[c] == abi.FCZ_E_LIMIT and blobs.blob(c) == b"", name
            continue
        assert blobs.status[c] == 0, (name, blobs.status[c])
        assert blobs.blob(c) == H.oracle_encode(one, 0, 25), name
        # "shuffled", "missing_atom" (a CA): the reference's backbone is the atoms named N / CA / C in INPUT order, however many
        # there are (filterBackbone, src/atom_coordinate.cpp:135-143), while the canonical layout stores every atom in its
        # table slot: backbone atoms out of order are encoded as if in order, a missing one reads (0,0,0) like a missing
        # side-chain atom does in the reference too (findFirstAtomCoords) -- a documented property of the slot layout
        # (DESIGN.md section 2); a missing SIDE-CHAIN atom is the same in both ("missing_sidechain_atom")
        # "unknown_residue": the reference's compress() dies on an uncaught std::out_of_range (AAS.at, src/sidechain.cpp:177)
        # and takes the process with it; here such a residue is encoded as UNK, like a decode of codes 24..31
        # "numbering_step_back": the reference looks its anchor atoms up by residue NUMBER (src/foldcomp.cpp:756-760); with
        # numbers that repeat it reads past its vectors (a segmentation fault here)
        if ref is not None and len(rt) >= 3 and name not in ("shuffled", "missing_atom", "unknown_residue", "numbering_step_back"):
            try:
                want = ref.compress(name, t.decode())
            except Exception:
                continue
            assert H.masked(blobs.blob(c)) == H.masked(want), name
            n_ref += 1
    assert ref is None or n_ref >= 40


def test_text_to_fcz_headline_shape(engine):
    """2 000 chains of 350 residues as PDB text in one call: every blob identical to the oracle's on the chain the text came
    from (the writer's %8.3f text IS the float, so the round trip through text is exact)."""
    batch = synth.generate(2000, 350, seed=4)
    out = engine.pdb_text_host(batch)
    titles = [batch.title(c).encode() for c in range(batch.n_chains)]
    blobs = engine.encode_pdb_text_host(out, titles)
    assert not blobs.status.any()
    rt = engine.decode_host(blobs)
    assert np.array_equal(rt.res_type, batch.res_type)
    want = H.oracle_encode_batch(_roundtrip_batch(engine, out, batch), 25)
    assert np.array_equal(blobs.blob_off, want.blob_off)
    assert np.array_equal(blobs.bytes[: int(want.blob_off[-1])], want.bytes[: int(want.blob_off[-1])])


def _roundtrip_batch(engine, texts, batch):
    """The chains as the parser sees them: coordinates and B-factors rounded through the PDB columns."""
    import torch

    d = DeviceTextBatch(texts.n_chains, len(texts.bytes) + 16, torch.device("cuda", engine.device))
    d.text_off.copy_(torch.from_numpy(texts.text_off.view(np.int64)))
    d.bytes[: len(texts.bytes)].copy_(torch.from_numpy(texts.bytes))
    got = engine.parse_pdb_device(d)
    engine.sync()
    hb = got.to_host()
    hb.title_off, hb.titles = batch.title_off, batch.titles
    return hb


def test_text_entry_points_on_empty_and_all_failed_batches(engine):
    """No entries at all, and a batch in which no entry parses: empty blobs, statuses, no fault."""
    import torch

    empty = engine.encode_pdb_text_host(_texts_batch([]), [])
    assert empty.n_chains == 0 and int(empty.blob_off[-1]) == 0
    texts = [b"", b"HEADER only\n", b"ATOM  short\n"]
    out = engine.encode_pdb_text_host(_texts_batch(texts), [b"a", b"b", b"c"])
    assert list(out.status) == [abi.FCZ_E_PARSE_NOATOM, abi.FCZ_E_PARSE_NOATOM, abi.FCZ_E_PARSE_RECORD] and int(out.blob_off[-1]) == 0
    got = engine.parse_pdb_device(_device_texts(texts, torch.device("cuda", engine.device)))
    engine.sync()
    assert int(got.res_off[-1].item()) == 0 and int(got.atom_off[-1].item()) == 0
