"""The certified shortcuts of the encode kernel (fcz_math.h: cos_ref, deg_ref; fcz_codec.h: sc_byte_fast)
must return EXACTLY what the reference's sequence returns -- on random inputs, on inputs sitting on and
next to every side-chain threshold (where the single-precision cosine is ambiguous and the exact path must
take over), and on degenerate inputs.  Runs the product header through the one-thread host model."""
import ctypes as C

import numpy as np
import pytest

import helpers as H


def _lib(perturb=False):
    lib = H.emu(perturb)
    lib.emu_sc_bytes.argtypes = [C.c_void_p] * 3 + [C.c_uint32, C.c_void_p, C.c_void_p]
    lib.emu_thresholds.argtypes = [C.c_void_p, C.c_void_p]
    lib.emu_cos_deg_mismatches.restype = C.c_uint64
    lib.emu_cos_deg_mismatches.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64)]
    return lib


def _sc(lib, inner, p, neg):
    inner = np.ascontiguousarray(inner, np.float32)
    p = np.ascontiguousarray(p, np.float32)
    neg = np.ascontiguousarray(neg, np.uint8)
    fast = np.zeros(len(inner), np.uint8)
    exact = np.zeros(len(inner), np.uint8)
    lib.emu_sc_bytes(inner.ctypes.data, p.ctypes.data, neg.ctypes.data, len(inner), fast.ctypes.data, exact.ctypes.data)
    return fast, exact


PERTURB = pytest.mark.parametrize("perturb", [False, True], ids=["ieee", "perturbed"])


@PERTURB
def test_side_chain_bytes_random(perturb):
    lib = _lib(perturb)
    rng = np.random.default_rng(0)
    n = 2_000_000
    # cosines spread over [-1,1] incl. the ill-conditioned ends; p over many magnitudes
    p = np.exp(rng.uniform(np.log(1e-6), np.log(1e6), n)).astype(np.float32)
    c = np.cos(rng.uniform(0, np.pi, n))
    inner = (c * np.sqrt(p.astype(np.float64))).astype(np.float32)
    neg = rng.integers(0, 2, n)
    fast, exact = _sc(lib, inner, p, neg)
    assert np.array_equal(fast, exact)
    assert len(np.unique(exact)) == 256  # every byte value occurs


@PERTURB
def test_side_chain_bytes_on_every_threshold(perturb):
    lib = _lib(perturb)
    pos = np.zeros(128, np.float32)
    neg_t = np.zeros(128, np.float32)
    lib.emu_thresholds(pos.ctypes.data, neg_t.ctypes.data)
    assert np.all(np.diff(pos) > 0) and np.all(np.diff(neg_t[:127]) > 0)
    pos = -pos  # stored negated (increasing); the thresholds themselves are cosines
    cs, ng = [], []
    for arr, flag in ((pos, 0), (neg_t[:127], 1)):
        for t in arr:
            v = np.float32(t)
            for k in range(-12, 13):  # the threshold and its 12 float neighbours on each side
                x = v
                for _ in range(abs(k)):
                    x = np.nextafter(x, np.float32(2.0 if k > 0 else -2.0), dtype=np.float32)
                cs.append(x)
                ng.append(flag)
    cs = np.array(cs, np.float32)
    for scale in (1.0, 4.0, 0.25, 3.0):  # p = scale^2: inner = c*scale (exact for powers of two, rounded for 3)
        fast, exact = _sc(lib, cs * np.float32(scale), np.full(len(cs), scale * scale, np.float32), ng)
        assert np.array_equal(fast, exact), scale


@PERTURB
def test_side_chain_bytes_degenerate(perturb):
    lib = _lib(perturb)
    inner = np.array([0, 0, 1, -1, 1e-20, 5, -5, np.nan, 1, 1.0000001, -1.0000001], np.float32)
    p = np.array([0, 1, 0, 0, 1e-40, 1, 1, 1, np.inf, 1, 1], np.float32)
    for flag in (0, 1):
        fast, exact = _sc(lib, inner, p, np.full(len(inner), flag))
        assert np.array_equal(fast, exact), (flag, fast, exact)


def test_cosine_and_degree_shortcuts_random():
    lib = _lib()
    rng = np.random.default_rng(1)
    n = 3_000_000
    p = np.exp(rng.uniform(np.log(1e-4), np.log(1e4), n)).astype(np.float32)
    c = np.cos(rng.uniform(0, np.pi, n))
    c[: n // 10] = np.cos(rng.uniform(np.pi - 0.2, np.pi, n // 10))  # omega-like, |c| -> 1
    inner = (c * np.sqrt(p.astype(np.float64))).astype(np.float32)
    fb = C.c_uint64()
    bad = lib.emu_cos_deg_mismatches(inner.ctypes.data, p.ctypes.data, n, C.byref(fb))
    assert bad == 0
    assert fb.value < n // 1000  # the exact fallback is rare (expected ~1e-6 of items)


def test_fast_acos_sampled():
    """Every 97th float of [-1, 1] (44 M inputs): error of acos_deg_fast vs long double, and the certified float
    equals the reference expression wherever it certifies.  (stride 1 = all 2^31 floats, run once by hand:
    max relative error 2^-49.76, 0 wrong, 549 uncertified, see DESIGN.md.)"""
    lib = _lib()
    lib.emu_acos_check.restype = C.c_long
    lib.emu_acos_check.argtypes = [C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    wrong, unc = C.c_uint64(), C.c_uint64()
    e = lib.emu_acos_check(97, C.byref(wrong), C.byref(unc))
    assert e <= -46000, e
    assert wrong.value == 0
    assert unc.value < 2000


def test_float_acos_bound():
    """acosdeg_f (the float-first arccosine of the encoder) against the reference's float over every 97th float of
    [-1, 1], with the reciprocal square root it is handed off by up to +-4 ulp: the deviation must stay below
    FCZ_ACOS_E0 = 8e-5 degrees, the constant every certification in fcz_codec.h builds on.  (stride 1 = all 2^31 floats,
    run once by hand: 3.8147e-05.)"""
    lib = _lib()
    lib.emu_acosdeg_f_check.restype = C.c_double
    lib.emu_acosdeg_f_check.argtypes = [C.c_uint32]
    worst = lib.emu_acosdeg_f_check(97)
    assert 1e-6 < worst <= 4.0e-5, worst


@PERTURB
def test_float_first_backbone_path_is_exercised_and_exact(perturb):
    """The float-first path must (a) carry the synthetic chains -- no fall-back to the all-exact path, a few dozen
    re-evaluated values per 350-residue chain -- and (b) give the oracle's bytes, also when the approximate
    reciprocals it starts from are off by a few ulp (the perturbed build stands in for the GPU's MUFU results)."""
    from foldcomp_b200 import synth

    batch = synth.generate(120, 350, seed=4242)
    want = H.oracle_encode_batch(batch, 25)
    n_list = []
    for c in range(batch.n_chains):
        assert H.emu_encode(batch, c, 25, perturb=perturb) == want.blob(c), c
        n, fell_back, _ = H.emu_last_stats(perturb)
        assert not fell_back, c
        n_list.append(n)
    assert 10 <= np.mean(n_list) <= 120 and max(n_list) <= 256, (np.mean(n_list), max(n_list))
    mixed = synth.generate(60, synth.mixed_lengths(np.random.default_rng(3), 60), seed=99)
    for b in (10, 25, 50, 200):
        want = H.oracle_encode_batch(mixed, b)
        for c in range(mixed.n_chains):
            assert H.emu_encode(mixed, c, b, perturb=perturb) == want.blob(c), (b, c)


def test_float_first_path_on_tight_geometry():
    """Arrays with a tiny range (every residue the same conformation: disc_f is huge, every value sits next to a bin
    edge) and ideal helices: the undecided list overflows or the bounds cross, the chain takes the all-exact path,
    bytes still equal the oracle's."""
    from foldcomp_b200 import synth

    base = synth.generate(1, 80, seed=12)
    # repeat one residue pair's geometry: build by copying the oracle-decoded coordinates of a constant-angle chain is
    # overkill -- scaling the chain towards a line (x *= 1e-3 in two axes) gives near-degenerate, tightly ranged angles
    for scale in (1e-2, 1e-4):
        b = synth.generate(1, 80, seed=13)
        b.xyz = b.xyz.copy()
        b.xyz[:, 1:] *= np.float32(scale)
        for anchor in (25, 10):
            o = H.oracle_encode(b, 0, anchor)
            for pert in (False, True):
                assert H.emu_encode(b, 0, anchor, perturb=pert) == o, (scale, anchor, pert)


def test_decode_cossin_deg():
    """The decoder's degree-domain sincos: absolute error vs double below 1.2e-7 over [-720, 720] degrees
    (20 M points) and sane on the slow path beyond."""
    lib = _lib()
    lib.emu_cossin_check.restype = C.c_long
    lib.emu_cossin_check.argtypes = [C.c_double, C.c_double, C.c_uint64]
    assert lib.emu_cossin_check(-720.0, 720.0, 20_000_001) <= 120
    assert lib.emu_cossin_check(-181.0, 181.0, 20_000_001) <= 120
    assert lib.emu_cossin_check(721.0, 5000.0, 1_000_001) <= 8000  # float radians: half an ulp of 87 rad is 3.8e-6
