"""SURVEY.md section 8 f4, `check`: the oracle's restatement of Foldcomp::read + checkValidity against the unmodified
reference (CPU), and -- on a GPU -- fcz_check_batch against the oracle."""
import numpy as np
import pytest

import helpers as H
from foldcomp_b200 import abi
from foldcomp_b200.abi import HostBlobBatch


def _batch(blobs):
    off = np.zeros(len(blobs) + 1, np.uint64)
    off[1:] = np.cumsum([len(b) for b in blobs])
    data = np.frombuffer(b"".join(blobs), np.uint8).copy() if off[-1] else np.zeros(1, np.uint8)
    return HostBlobBatch(off, data)


def test_oracle_check_matches_reference():
    """Every complete blob: same read code and ValidityError class as the reference.  Truncated blobs are left out on the
    reference side: there it reads past the end of its stream and checks whatever its (partly uninitialised) buffers
    held, i.e. its answer is not a function of the input."""
    seen = set()
    for label, blob in H.check_cases():
        rc, v = H.oracle_check(blob)
        seen.add((rc, v))
        if label.startswith("cut") or label == "empty":
            assert rc in (abi.FCZ_E_TRUNCATED, abi.FCZ_E_MAGIC) and (rc == abi.FCZ_E_MAGIC or 1 <= v <= 3), (label, rc, v)
            continue
        if H.have_ref():
            assert H.ref_check(blob) == (rc, v), label
    # classes 0, 4, 5, 6, the three truncation classes and bad magic all occur
    assert {(0, 0), (0, 4), (0, 5), (0, 6), (abi.FCZ_E_MAGIC, 0)} <= seen
    assert {v for rc, v in seen if rc == abi.FCZ_E_TRUNCATED} == {1, 2, 3}


def test_validity_names():
    assert abi.VALIDITY[4] == "E_EMPTY_BACKBONE_ANGLE" and len(abi.VALIDITY) == len(abi.VALIDITY_MESSAGE) == 7


@pytest.mark.gpu
def test_gpu_check_matches_oracle(engine):
    cases = H.check_cases()
    rs, va = engine.check_host(_batch([b for _, b in cases]))
    for i, (label, blob) in enumerate(cases):
        assert (int(rs[i]), int(va[i])) == H.oracle_check(blob), label


@pytest.mark.gpu
def test_gpu_check_large_batch(engine):
    """10 000 good blobs + a zeroed one in the middle, through the host-memory call."""
    from foldcomp_b200 import synth

    batch = synth.generate(2000, 120, seed=77)
    blobs = engine.encode_host(batch)
    data = blobs.bytes.copy()
    c = 1234
    o_rec, o_sc, o_temp, size, L, n_sc = H.blob_sections(blobs.blob(c))
    b0 = int(blobs.blob_off[c])
    data[b0 + o_sc : b0 + o_sc + n_sc] = 0
    rs, va = engine.check_host(HostBlobBatch(blobs.blob_off, data))
    assert not rs.any()
    want = np.zeros(2000, np.int32)
    want[c] = 5
    assert np.array_equal(va, want)
