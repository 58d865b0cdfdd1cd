"""The certified shortcuts of the encode kernel (fcz_math.h: cos_ref, deg_ref; fcz_codec.h: sc_byte_fast)
must return EXACTLY what the reference's sequence returns -- on random inputs, on inputs sitting on and
next to every side-chain threshold (where the single-precision cosine is ambiguous and the exact path must
take over), and on degenerate inputs.  Runs the product header through the one-thread host model."""
import ctypes as C

import numpy as np

import helpers as H


def _lib():
    lib = H.emu()
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


def test_side_chain_bytes_random():
    lib = _lib()
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


def test_side_chain_bytes_on_every_threshold():
    lib = _lib()
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


def test_side_chain_bytes_degenerate():
    lib = _lib()
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


def test_decode_cossin_deg():
    """The decoder's degree-domain sincos: absolute error vs double below 1.2e-7 over [-720, 720] degrees
    (20 M points) and sane on the slow path beyond."""
    lib = _lib()
    lib.emu_cossin_check.restype = C.c_long
    lib.emu_cossin_check.argtypes = [C.c_double, C.c_double, C.c_uint64]
    assert lib.emu_cossin_check(-720.0, 720.0, 20_000_001) <= 120
    assert lib.emu_cossin_check(-181.0, 181.0, 20_000_001) <= 120
    assert lib.emu_cossin_check(721.0, 5000.0, 1_000_001) <= 8000  # float radians: half an ulp of 87 rad is 3.8e-6
