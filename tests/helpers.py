"""ctypes wrappers of the CHECKERS used by the tests:
  * oracle/libfcz_oracle.so       -- plain-C restatement (fcz_oracle.c)
  * oracle/_ref/libfoldcomp_ref.so -- the unmodified reference compiled in place (ref_shim.cpp)
  * tests/emu/libfcz_emu.so       -- one-thread host instantiation of the product codec header
None of this is product code.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from foldcomp_b200 import abi  # noqa: E402
from foldcomp_b200.abi import HostBlobBatch, HostChainBatch  # noqa: E402
from foldcomp_b200.tables import tables  # noqa: E402

ORACLE_SO = os.path.join(ROOT, "oracle", "libfcz_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libfoldcomp_ref.so")
EMU_SO = os.path.join(ROOT, "tests", "emu", "libfcz_emu.so")
REFERENCE_DIR = "/root/reference"
MASK = (14, 15, 22, 23)  # uninitialised CompressedFileHeader padding in the reference

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C")


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "libfcz_oracle.so"])
    if os.path.isdir(os.path.join(REFERENCE_DIR, "src")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
        subprocess.call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "pyref"])


EMU_PERTURB_SO = os.path.join(ROOT, "tests", "emu", "libfcz_emu_perturb.so")


def build_emu(perturb: bool = False):
    """perturb=True: the variant whose approximate reciprocal (square root) is pushed off by up to +-3 ulp, like a GPU's
    MUFU results may be (fcz_math.h: FCZ_EMU_PERTURB) -- the certified shortcuts must not care."""
    src = os.path.join(ROOT, "tests", "emu", "fcz_emu.cpp")
    so = EMU_PERTURB_SO if perturb else EMU_SO
    deps = [src] + [os.path.join(ROOT, "foldcomp_b200", "csrc", f) for f in ("fcz_codec.h", "fcz_math.h", "fcz_format.h", "fcz_tables.h", "fcz_text.h", "fcz_parse.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-std=c++17"]
                              + (["-DFCZ_EMU_PERTURB"] if perturb else []) + ["-o", so, src])


_cache = {}


def oracle():
    if "oracle" not in _cache:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        lib = C.CDLL(ORACLE_SO)
        lib.fcz_oracle_encode_chain.restype = C.c_int64
        lib.fcz_oracle_encode_chain.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_uint32, C.c_int32, C.c_void_p, C.c_uint64]
        lib.fcz_oracle_peek.restype = C.c_int
        lib.fcz_oracle_peek.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
        lib.fcz_oracle_decode_chain.restype = C.c_int
        lib.fcz_oracle_decode_chain.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.fcz_oracle_encode_batch.restype = C.c_int
        lib.fcz_oracle_encode_batch.argtypes = [C.POINTER(abi.FczChainBatch), C.POINTER(abi.FczBlobBatch), C.c_int32, C.c_int]
        lib.fcz_oracle_decode_plan.restype = C.c_int
        lib.fcz_oracle_decode_plan.argtypes = [C.POINTER(abi.FczBlobBatch), C.POINTER(abi.FczChainBatch), C.POINTER(abi.FczSizes)]
        lib.fcz_oracle_decode_batch.restype = C.c_int
        lib.fcz_oracle_decode_batch.argtypes = [C.POINTER(abi.FczBlobBatch), C.POINTER(abi.FczChainBatch), C.c_int, C.c_int]
        _cache["oracle"] = lib
    return _cache["oracle"]


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref():
    if "ref" not in _cache:
        lib = C.CDLL(REF_SO)
        lib.ref_compress.restype = C.c_int
        lib.ref_compress.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_char, C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        lib.ref_decompress.restype = C.c_int
        lib.ref_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.ref_roundtrip_batch.restype = C.c_int
        lib.ref_roundtrip_batch.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        lib.ref_max_threads.restype = C.c_int
        lib.ref_type_natoms.restype = C.c_int
        lib.ref_type_natoms.argtypes = [C.c_int]
        _cache["ref"] = lib
    return _cache["ref"]


def emu(perturb: bool = False):
    key = "emu_perturb" if perturb else "emu"
    if key not in _cache:
        build_emu(perturb)
        lib = C.CDLL(EMU_PERTURB_SO if perturb else EMU_SO)
        lib.emu_encode_chain.restype = C.c_int64
        lib.emu_encode_chain.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_uint32, C.c_int32, C.c_void_p, C.c_uint64]
        lib.emu_decode_chain.restype = C.c_int
        lib.emu_decode_chain.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.emu_last_stats.argtypes = [C.c_void_p]
        _cache[key] = lib
    return _cache[key]


def emu_last_stats(perturb: bool = False):
    """(values re-evaluated exactly, fell back to the all-exact path) of the last emu_encode on this thread."""
    out = np.zeros(8, np.uint32)
    emu(perturb).emu_last_stats(out.ctypes.data)
    return int(out[0]), bool(out[1]), out[2:].copy()  # ... and the list entries per array (phi psi omega N-CA-C CA-C-N C-N-CA)


# --------------------------------------------------------------------------- per-chain convenience


def _chain_args(b: HostChainBatch, c: int):
    r0, r1 = int(b.res_off[c]), int(b.res_off[c + 1])
    a0, a1 = int(b.atom_off[c]), int(b.atom_off[c + 1])
    t0, t1 = int(b.title_off[c]), int(b.title_off[c + 1])
    rt = np.ascontiguousarray(b.res_type[r0:r1])
    bf = np.ascontiguousarray(b.bfactor[r0:r1])
    xyz = np.ascontiguousarray(b.xyz[a0:a1])
    title = bytes(b.titles[t0:t1])
    meta = np.ascontiguousarray(b.meta[c : c + 1])
    return rt, bf, xyz, title, meta


def _encode_with(fn, b: HostChainBatch, c: int, anchor: int):
    rt, bf, xyz, title, meta = _chain_args(b, c)
    cap = abi.encode_bound(1, len(rt), len(xyz), len(title), anchor) + 64
    out = np.zeros(cap, np.uint8)
    n = fn(rt.ctypes.data, len(rt), xyz.ctypes.data, bf.ctypes.data, meta.ctypes.data, title, len(title), anchor, out.ctypes.data, cap)
    if n < 0:
        return int(n)
    return bytes(out[:n])


def oracle_encode(b: HostChainBatch, c: int = 0, anchor: int = 25):
    return _encode_with(oracle().fcz_oracle_encode_chain, b, c, anchor)


def emu_encode(b: HostChainBatch, c: int = 0, anchor: int = 25, perturb: bool = False):
    return _encode_with(emu(perturb).emu_encode_chain, b, c, anchor)


def ref_encode(b: HostChainBatch, c: int = 0, anchor: int = 25):
    rt, bf, xyz, title, meta = _chain_args(b, c)
    m = meta[0]
    cap = abi.encode_bound(1, len(rt), len(xyz), len(title), anchor) + 64
    out = np.zeros(cap, np.uint8)
    n = C.c_size_t(0)
    oxt = np.ascontiguousarray(m["oxt"], np.float32)
    rc = ref().ref_compress(rt.ctypes.data, len(rt), xyz.ctypes.data, bf.ctypes.data, int(m["has_oxt"]), oxt.ctypes.data,
                            int(m["idx_residue"]), int(m["idx_atom"]), bytes([int(m["chain"])]), title, len(title), anchor,
                            out.ctypes.data, cap, C.byref(n))
    if rc != 0:
        return rc
    return bytes(out[: n.value])


class Decoded:
    def __init__(self, res_type, bfac, xyz, meta, title):
        self.res_type, self.bfactor, self.xyz, self.meta, self.title = res_type, bfac, xyz, meta, title


def _peek(blob: bytes):
    L, na, tl = C.c_uint32(), C.c_uint64(), C.c_uint32()
    buf = np.frombuffer(blob, np.uint8)
    rc = oracle().fcz_oracle_peek(buf.ctypes.data, len(blob), C.byref(L), C.byref(na), C.byref(tl))
    return rc, L.value, na.value, tl.value


def _decode_with(fn, blob: bytes, use_alt: bool = False):
    rc, L, na, tl = _peek(blob)
    if rc:
        return rc
    buf = np.frombuffer(blob, np.uint8)
    rt = np.zeros(L, np.uint8)
    bf = np.zeros(L, np.float32)
    xyz = np.zeros((na, 3), np.float32)
    meta = np.zeros(1, abi.META_DTYPE)
    title = np.zeros(max(tl, 1), np.uint8)
    rc = fn(buf.ctypes.data, len(blob), int(use_alt), rt.ctypes.data, bf.ctypes.data, xyz.ctypes.data, meta.ctypes.data, title.ctypes.data)
    if rc:
        return rc
    return Decoded(rt, bf, xyz, meta[0], bytes(title[:tl]))


def oracle_decode(blob: bytes, use_alt: bool = False):
    return _decode_with(oracle().fcz_oracle_decode_chain, blob, use_alt)


def emu_decode(blob: bytes, use_alt: bool = False):
    return _decode_with(emu().emu_decode_chain, blob, use_alt)


def ref_decode(blob: bytes, use_alt: bool = False):
    rc, L, na, tl = _peek(blob)
    if rc:
        return rc
    buf = np.frombuffer(blob, np.uint8)
    xyz = np.zeros((na + 1, 3), np.float32)
    bf = np.zeros(L, np.float32)
    rt = np.zeros(L, np.uint8)
    n_atoms, n_res, has_oxt = C.c_int(), C.c_int(), C.c_int()
    rc = ref().ref_decompress(buf.ctypes.data, len(blob), int(use_alt), xyz.ctypes.data, na + 1, C.byref(n_atoms), bf.ctypes.data,
                              rt.ctypes.data, L, C.byref(n_res), C.byref(has_oxt))
    if rc:
        return rc
    n = n_atoms.value - (1 if has_oxt.value else 0)
    meta = np.zeros(1, abi.META_DTYPE)[0]
    meta["has_oxt"] = has_oxt.value
    if has_oxt.value:
        meta["oxt"] = xyz[n]
    return Decoded(rt, bf, xyz[:n].copy(), meta, b"")


def masked(blob: bytes) -> bytes:
    b = bytearray(blob)
    for i in MASK:
        if i < len(b):
            b[i] = 0
    return bytes(b)


def rmsd(a: np.ndarray, b: np.ndarray) -> float:
    d = a.astype(np.float64) - b.astype(np.float64)
    return float(np.sqrt((d * d).sum(axis=1).mean()))


def max_dev(a: np.ndarray, b: np.ndarray) -> float:
    d = a.astype(np.float64) - b.astype(np.float64)
    return float(np.sqrt((d * d).sum(axis=1)).max())


def backbone_mask(res_type: np.ndarray) -> np.ndarray:
    tb = tables()
    m = []
    for c in res_type:
        n = int(tb.natoms[int(c)])
        m += [True, True, True] + [False] * (n - 3)
    return np.array(m, bool)


# --------------------------------------------------------------------------- batch (OpenMP) oracle


def oracle_encode_batch(b: HostChainBatch, anchor: int = 25, threads: int = 0) -> HostBlobBatch:
    cap = abi.encode_bound(b.n_chains, b.n_res, b.n_atoms, len(b.titles), anchor) + 64
    out = HostBlobBatch.empty(b.n_chains, cap)
    sin, sout = b.as_struct(), out.as_struct()
    rc = oracle().fcz_oracle_encode_batch(C.byref(sin), C.byref(sout), anchor, threads)
    assert rc == 0, rc
    out.bytes = out.bytes[: int(out.blob_off[-1])]
    return out


def oracle_decode_batch(blobs: HostBlobBatch, use_alt: bool = False, threads: int = 0) -> HostChainBatch:
    n = blobs.n_chains
    plan = HostChainBatch.empty(n)
    sizes = abi.FczSizes()
    sin, sp = blobs.as_struct(), plan.as_struct()
    rc = oracle().fcz_oracle_decode_plan(C.byref(sin), C.byref(sp), C.byref(sizes))
    assert rc == 0
    out = HostChainBatch(
        res_off=plan.res_off, atom_off=plan.atom_off, title_off=plan.title_off,
        res_type=np.zeros(sizes.n_res, np.uint8), bfactor=np.zeros(sizes.n_res, np.float32),
        xyz=np.zeros((sizes.n_atoms, 3), np.float32), titles=np.zeros(sizes.n_title_bytes, np.uint8),
        meta=np.zeros(n, abi.META_DTYPE), status=plan.status,
    )
    so = out.as_struct()
    rc = oracle().fcz_oracle_decode_batch(C.byref(sin), C.byref(so), int(use_alt), threads)
    assert rc == 0
    return out


def per_chain_deviation(a: HostChainBatch, b: HostChainBatch):
    """(worst backbone RMSD, worst all-atom RMSD, worst max deviation) over chains of two decoded batches."""
    assert np.array_equal(a.atom_off, b.atom_off) and np.array_equal(a.res_off, b.res_off)
    d2 = ((a.xyz.astype(np.float64) - b.xyz.astype(np.float64)) ** 2).sum(axis=1)
    bb = backbone_mask_fast(a.res_type)
    aoff = a.atom_off.astype(np.int64)
    worst_bb = worst_all = worst_max = 0.0
    cs_all = np.concatenate([[0.0], np.cumsum(d2)])
    cs_bb = np.concatenate([[0.0], np.cumsum(d2 * bb)])
    cnt_bb = np.concatenate([[0], np.cumsum(bb)])
    for c in range(a.n_chains):
        lo, hi = aoff[c], aoff[c + 1]
        if hi == lo:
            continue
        worst_all = max(worst_all, np.sqrt((cs_all[hi] - cs_all[lo]) / (hi - lo)))
        worst_bb = max(worst_bb, np.sqrt((cs_bb[hi] - cs_bb[lo]) / max(cnt_bb[hi] - cnt_bb[lo], 1)))
    worst_max = float(np.sqrt(d2.max())) if len(d2) else 0.0
    return float(worst_bb), float(worst_all), worst_max


def backbone_mask_fast(res_type: np.ndarray) -> np.ndarray:
    tb = tables()
    nat = tb.natoms[res_type.astype(np.int64)]
    starts = np.concatenate([[0], np.cumsum(nat)])[:-1]
    m = np.zeros(int(nat.sum()), bool)
    for k in range(3):
        m[starts + k] = True
    return m


def long_chain(L: int, seed: int = 0) -> HostChainBatch:
    """One chain of L residues made by joining 350-residue synthetic chains end to end (seconds even for the
    format's maximum of 65535 residues; the generator's serial NeRF walk would take a minute).  The joints have
    arbitrary geometry, which is fine: encode parity is bit-exact on any coordinates and decode parity compares
    two decoders on the same blob."""
    from foldcomp_b200 import synth

    nparts = (L + 349) // 350
    src = synth.generate(nparts, 350, seed=4242 + seed)
    # shift part i so that the chain does not fold back onto itself
    xyz = src.xyz.copy()
    for i in range(nparts):
        a0, a1 = int(src.atom_off[i]), int(src.atom_off[i + 1])
        xyz[a0:a1] += np.float32(3.8) * i
    rt = src.res_type[:L]
    A = int(tables().natoms[rt].sum())
    meta = src.meta[:1].copy()
    meta["n_atom"] = np.uint16((A + int(meta["has_oxt"][0])) & 0xFFFF)  # the header field is 16 bits (src/foldcomp.h:118-131)
    return abi.concat_chains([(rt, src.bfactor[:L], xyz[:A], np.frombuffer(f"long_{L}".encode(), np.uint8), meta)])


# --------------------------------------------------------------------------- text


def _format_with(fn, b: HostChainBatch, c: int, use_alt: bool, via_ref: bool = False) -> bytes:
    rt, bf, xyz, title, meta = b._slice(c)
    rt = np.ascontiguousarray(rt)
    bf = np.ascontiguousarray(bf)
    xyz = np.ascontiguousarray(xyz)
    tb = bytes(title)
    cap = 128 + len(tb) * 2 + 120 * (len(xyz) + 2)
    out = np.zeros(cap, np.uint8)
    m = np.ascontiguousarray(meta)
    if via_ref:
        oxt = np.ascontiguousarray(m["oxt"][0], np.float32)
        n = fn(rt.ctypes.data, len(rt), xyz.ctypes.data, bf.ctypes.data, int(m["has_oxt"][0]), oxt.ctypes.data, int(m["idx_residue"][0]),
               int(m["idx_atom"][0]), bytes([int(m["chain"][0])]), tb, len(tb), int(use_alt), out.ctypes.data, cap)
    else:
        n = fn(rt.ctypes.data, len(rt), xyz.ctypes.data, bf.ctypes.data, m.ctypes.data, tb, len(tb), int(use_alt), out.ctypes.data, cap)
    assert 0 <= n <= cap, n
    return bytes(out[:n])


def _text_protos(lib, prefix):
    f = getattr(lib, prefix + "format_pdb")
    f.restype = C.c_int64
    f.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_uint32, C.c_int, C.c_void_p, C.c_uint64]
    e = getattr(lib, prefix + "extract")
    e.restype = C.c_int64
    e.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_uint64]
    return f, e


def oracle_format_pdb(b: HostChainBatch, c: int = 0, use_alt: bool = False) -> bytes:
    f, _ = _text_protos(oracle(), "fcz_oracle_")
    return _format_with(f, b, c, use_alt)


def emu_format_pdb(b: HostChainBatch, c: int = 0, use_alt: bool = False) -> bytes:
    f, _ = _text_protos(emu(), "emu_")
    return _format_with(f, b, c, use_alt)


def ref_format_pdb(b: HostChainBatch, c: int = 0, use_alt: bool = False) -> bytes:
    lib = ref()
    lib.ref_format_pdb.restype = C.c_int64
    lib.ref_format_pdb.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_char, C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    return _format_with(lib.ref_format_pdb, b, c, use_alt, via_ref=True)


def ref_decompress_to_pdb(blob: bytes, use_alt: bool = False) -> bytes:
    lib = ref()
    lib.ref_decompress_to_pdb.restype = C.c_int64
    lib.ref_decompress_to_pdb.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t]
    cap = 1 << 24
    out = np.zeros(cap, np.uint8)
    n = lib.ref_decompress_to_pdb(blob, len(blob), int(use_alt), out.ctypes.data, cap)
    assert 0 <= n <= cap, n
    return bytes(out[:n])


def _extract_with(fn, blob: bytes, type_: int, digits: int) -> bytes:
    cap = 8 * len(blob) + 64
    out = np.zeros(cap, np.uint8)
    n = fn(blob, len(blob), type_, digits, out.ctypes.data, cap)
    assert 0 <= n <= cap, n
    return bytes(out[:n])


def oracle_extract(blob: bytes, type_: int, digits: int = 2) -> bytes:
    return _extract_with(_text_protos(oracle(), "fcz_oracle_")[1], blob, type_, digits)


def emu_extract(blob: bytes, type_: int, digits: int = 2) -> bytes:
    return _extract_with(_text_protos(emu(), "emu_")[1], blob, type_, digits)


def ref_extract(blob: bytes, type_: int, digits: int = 2) -> bytes:
    lib = ref()
    lib.ref_extract.restype = C.c_int64
    lib.ref_extract.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    return _extract_with(lib.ref_extract, blob, type_, digits)


def decoded_as_batch(dec, title: bytes | None = None) -> HostChainBatch:
    """A `Decoded` (oracle/emu decode of one blob; the OXT lives in meta) as a 1-chain batch in the engine's layout."""
    meta = np.zeros(1, abi.META_DTYPE)
    meta[0] = dec.meta
    return abi.concat_chains([(dec.res_type, dec.bfactor, dec.xyz, np.frombuffer(dec.title if title is None else title, np.uint8), meta)])


def extreme_text_chain(seed: int = 0) -> HostChainBatch:
    """A chain whose fields overflow their PDB columns (std::setw is a minimum width): residue numbers beyond 9999,
    serials beyond 99999, coordinates beyond 9999.999 / below -999.999, B-factors beyond 999.99, values that round
    to -0.000, and a title long enough for continuation lines."""
    from foldcomp_b200 import synth

    b = synth.generate(1, 700, seed=900 + seed)
    rng = np.random.default_rng(seed)
    xyz = b.xyz.copy()
    idx = rng.choice(len(xyz), 60, replace=False)
    vals = np.array([12345.678, -1234.567, 99999.9996, -0.0004, -0.0005, 0.0004999, 9999.9995, -999.9995, 1e6, -1e6, 123456.7, 0.9995, -0.9995,
                     2147480000.0, -2147480000.0], np.float32)
    for j, i in enumerate(idx):
        xyz[i, j % 3] = vals[j % len(vals)]
    b.xyz = xyz
    bf = b.bfactor.copy()
    bf[:8] = [1000.0, 999.995, -100.0, -0.004, 0.005, 99.995, 12345.5, -99.994]
    b.bfactor = bf
    b.meta["idx_residue"] = 9700
    b.meta["idx_atom"] = 65000
    title = (b"a very long title " * 12)[:205]
    b.titles = np.frombuffer(title, np.uint8).copy()
    b.title_off = np.array([0, len(title)], np.uint32)
    return b


def oracle_unpack_angles(blob: bytes) -> np.ndarray:
    """[L, 6] continuised phi, psi, omega, N-CA-C, CA-C-N, C-N-CA of every residue record (src/foldcomp.cpp:122-153)."""
    lib = oracle()
    lib.fcz_oracle_unpack_angles.restype = C.c_int64
    lib.fcz_oracle_unpack_angles.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p]
    L = int.from_bytes(blob[4:6], "little")
    out = np.zeros((L, 6), np.float32)
    n = lib.fcz_oracle_unpack_angles(blob, len(blob), out.ctypes.data)
    assert n == L, n
    return out


def oracle_backbone_angles(b: HostChainBatch, c: int = 0) -> np.ndarray:
    """[L, 6] per residue: psi, omega, next phi, bond angles at N, CA, C -- before quantisation (src/foldcomp.cpp:484-496)."""
    lib = oracle()
    lib.fcz_oracle_backbone_angles.restype = C.c_int
    lib.fcz_oracle_backbone_angles.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    r0, r1, a0, a1 = int(b.res_off[c]), int(b.res_off[c + 1]), int(b.atom_off[c]), int(b.atom_off[c + 1])
    rt, xyz = np.ascontiguousarray(b.res_type[r0:r1]), np.ascontiguousarray(b.xyz[a0:a1])
    out = np.zeros((r1 - r0, 6), np.float32)
    rc = lib.fcz_oracle_backbone_angles(rt.ctypes.data, r1 - r0, xyz.ctypes.data, out.ctypes.data)
    assert rc == 0, rc
    return out


def emu_backbone_angles(b: HostChainBatch, c: int = 0) -> np.ndarray:
    lib = emu()
    lib.emu_backbone_angles.restype = C.c_int
    lib.emu_backbone_angles.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    r0, r1, a0, a1 = int(b.res_off[c]), int(b.res_off[c + 1]), int(b.atom_off[c]), int(b.atom_off[c + 1])
    rt, xyz = np.ascontiguousarray(b.res_type[r0:r1]), np.ascontiguousarray(b.xyz[a0:a1])
    out = np.zeros((r1 - r0, 6), np.float32)
    assert lib.emu_backbone_angles(rt.ctypes.data, r1 - r0, xyz.ctypes.data, out.ctypes.data) == 0
    return out


def get_data_lists(ang: np.ndarray):
    """(torsion_angles, bond_angles, phi, psi, omega) as the reference's get_data(pdb_text) lists them, from the [L, 6] layout."""
    L = len(ang)
    tors = ang[: L - 1, :3].reshape(-1)
    bond = ang[:, 3:].reshape(-1)[1 : 3 * L - 1]
    return tors, bond, tors[2::3], tors[0::3], tors[1::3]


# --------------------------------------------------------------------------- degenerate inputs (shared generator)


def _nanf(bits: int) -> np.float32:
    return np.array([bits], np.uint32).view(np.float32)[0]


def degenerate_chains():
    """[(kind, HostChainBatch of one chain)]: inputs on which the reference's arithmetic leaves the beaten path.
    Kinds 0-9 (120 chains, the round-1 set): constant B-factors (the discretiser divides by zero and casts NaN), coincident
    atoms, missing atoms at (0,0,0), scaled, collinear triple, NaN coordinate, NaN B-factor, integer lattice, inf
    coordinate, two-valued B-factors.  Kinds 10-14 aim at the FIRST element of every array, whose NaN reaches the
    header floats (std::min_element keeps a NaN first element, src/discretizer.cpp:27-28): NaN / inf in each of the
    first six backbone atoms, neighbouring backbone atoms coincident (0/0), NaN first B-factors of several payloads
    and signs, all-infinite B-factors, two different NaN payloads among the first atoms."""
    from foldcomp_b200 import synth

    rng = np.random.default_rng(7)
    out = []
    for trial in range(120):
        L = int(rng.integers(2, 60))
        batch = synth.generate(1, L, seed=1000 + trial)
        x, bf = batch.xyz.copy(), batch.bfactor.copy()
        A, kind = len(x), trial % 10
        if kind == 0:
            bf[:] = 50.0
        elif kind == 1:
            x[rng.integers(0, A)] = x[rng.integers(0, A)]
        elif kind == 2:
            x[rng.integers(0, A, 3)] = 0.0
        elif kind == 3:
            x *= np.float32(100.0)
        elif kind == 4:
            x[2] = x[1] + (x[1] - x[0])
        elif kind == 5:
            x[rng.integers(0, A)] = np.nan
        elif kind == 6:
            bf[rng.integers(0, L)] = np.nan
        elif kind == 7:
            x[:] = np.round(x)
        elif kind == 8:
            x[rng.integers(0, A)] = np.inf
        else:
            bf[:] = rng.choice([0.0, 100.0], L)
        batch.xyz, batch.bfactor = x, bf
        out.append((kind, batch))

    nat = tables().natoms
    seed = [2000]

    def fresh(L=None):
        seed[0] += 1
        b = synth.generate(1, L if L else int(rng.integers(3, 40)), seed=seed[0])
        b.xyz, b.bfactor = b.xyz.copy(), b.bfactor.copy()
        n0 = int(nat[int(b.res_type[0])])
        return b, [0, 1, 2, n0, n0 + 1, n0 + 2]  # atom indices of the first six backbone atoms

    for j in range(6):  # kind 10: NaN in one component / all components of backbone atom j
        for comp in (0, 1, 2, None):
            b, bbi = fresh()
            if comp is None:
                b.xyz[bbi[j]] = np.nan
            else:
                b.xyz[bbi[j], comp] = np.nan
            out.append((10, b))
    for j in range(6):  # kind 11: +inf / -inf
        for v in (np.inf, -np.inf):
            b, bbi = fresh()
            b.xyz[bbi[j], int(rng.integers(0, 3))] = v
            out.append((11, b))
        b, bbi = fresh()
        b.xyz[bbi[j]] = np.inf
        out.append((11, b))
    for j in range(5):  # kind 12: backbone atoms j and j+1 coincide (zero bond vector: 0/0)
        b, bbi = fresh()
        b.xyz[bbi[j + 1]] = b.xyz[bbi[j]]
        out.append((12, b))
    for first in (_nanf(0x7FC00000), _nanf(0xFFC00000), _nanf(0x7F812345), _nanf(0xFFA00001), np.float32(np.inf), np.float32(-np.inf)):
        b, _ = fresh()  # kind 13: first B-factor NaN (quiet, negative, signalling payloads) or infinite
        b.bfactor[0] = first
        out.append((13, b))
    for v in (np.inf, -np.inf):  # ... and every B-factor infinite (max - min = inf - inf)
        b, _ = fresh()
        b.bfactor[:] = v
        out.append((13, b))
    return out


def has_collinear_backbone(b: HostChainBatch, c: int = 0, eps: float = 1e-6) -> bool:
    """True when three consecutive backbone atoms of chain c are collinear to within eps (sine of their angle), or a
    bond vector vanishes: the geometry on which Nerf::place_atom's frame (src/nerf.cpp:52-85) is undefined."""
    a0, a1 = int(b.atom_off[c]), int(b.atom_off[c + 1])
    r0, r1 = int(b.res_off[c]), int(b.res_off[c + 1])
    p = b.xyz[a0:a1][backbone_mask(b.res_type[r0:r1])].astype(np.float64)
    u, v = p[:-2] - p[1:-1], p[2:] - p[1:-1]
    with np.errstate(all="ignore"):
        s = np.linalg.norm(np.cross(u, v), axis=1) / (np.linalg.norm(u, axis=1) * np.linalg.norm(v, axis=1))
    s = s[np.isfinite(s)]
    return bool(len(s) and s.min() < eps)


def decode_mismatch(got: np.ndarray, want: np.ndarray, res_type: np.ndarray, tol_bb_rmsd: float, tol_max: float):
    """None when two decodes of one chain agree: identical NaN / inf masks, finite atoms within the tolerances."""
    ng, nw = np.isnan(got), np.isnan(want)
    if not np.array_equal(ng, nw):
        return f"NaN masks differ ({int(ng.sum())} vs {int(nw.sum())})"
    with np.errstate(all="ignore"):
        if not np.array_equal(np.isinf(got), np.isinf(want)) or not np.array_equal(got[np.isinf(want)], want[np.isinf(want)]):
            return "inf masks differ"
        fin = np.isfinite(want).all(axis=1)
        if not fin.any():
            return None
        d = np.sqrt(((got[fin].astype(np.float64) - want[fin].astype(np.float64)) ** 2).sum(axis=1))
    if d.max() > tol_max:
        return f"max deviation {d.max():.3e}"
    bb = backbone_mask(res_type)[fin]
    if bb.any() and np.sqrt((d[bb] ** 2).mean()) > tol_bb_rmsd:
        return f"backbone rmsd {np.sqrt((d[bb] ** 2).mean()):.3e}"
    return None


# --------------------------------------------------------------------------- check (Foldcomp::checkValidity)


def oracle_check(blob: bytes):
    """(read status, ValidityError class) of the oracle's restatement of read + checkValidity."""
    lib = oracle()
    lib.fcz_oracle_check.restype = C.c_int
    lib.fcz_oracle_check.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_int)]
    buf = np.frombuffer(blob, np.uint8) if len(blob) else np.zeros(1, np.uint8)
    v = C.c_int()
    rc = lib.fcz_oracle_check(buf.ctypes.data, len(blob), C.byref(v))
    return int(rc), int(v.value)


def ref_check(blob: bytes):
    lib = ref()
    lib.ref_check.restype = C.c_int
    lib.ref_check.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]
    buf = np.frombuffer(blob, np.uint8) if len(blob) else np.zeros(1, np.uint8)
    v = C.c_int()
    rc = lib.ref_check(buf.ctypes.data, len(blob), C.byref(v))
    return int(rc), int(v.value)


def blob_sections(blob: bytes):
    """(o_rec, o_sc, o_temp, size, L, n_sc) of an FCZ blob."""
    import struct

    L, = struct.unpack_from("<H", blob, 4)
    na = blob[12]
    n_sc, = struct.unpack_from("<I", blob, 16)
    T, = struct.unpack_from("<I", blob, 24)
    o_rec = 76 + 4 * na + T + 36 * na + 13
    o_sc = o_rec + 8 * L
    o_temp = o_sc + n_sc
    return o_rec, o_sc, o_temp, o_temp + 8 + L, L, n_sc


def unk_chain(L: int = 30, seed: int = 5) -> HostChainBatch:
    """One chain of L residues that are all UNK (backbone only): no side-chain torsions at all."""
    from foldcomp_b200 import synth

    b = synth.generate(1, L, seed=seed)
    xyz = b.xyz[backbone_mask(b.res_type)].copy()
    meta = b.meta.copy()
    meta["n_atom"] = 3 * L
    meta["has_oxt"] = 0
    return HostChainBatch(res_off=np.array([0, L], np.uint32), atom_off=np.array([0, 3 * L], np.uint64), title_off=b.title_off,
                          res_type=np.full(L, 23, np.uint8), bfactor=b.bfactor, xyz=xyz, titles=b.titles, meta=meta,
                          status=np.zeros(1, np.int32))


def check_cases():
    """[(label, blob)] for read + checkValidity: good blobs, each section zeroed (alone and together), a chain without
    side-chain torsions, bad magic, empty input, blobs cut inside every section."""
    from foldcomp_b200 import synth

    out = []
    for i in range(3):
        out.append((f"good{i}", oracle_encode(synth.generate(1, 20 + 40 * i, seed=300 + i), 0, 25)))
    base = out[1][1]
    o_rec, o_sc, o_temp, size, L, n_sc = blob_sections(base)

    def zero_bb(b):
        b = bytearray(b)
        for r in range(L):
            b[o_rec + 8 * r] &= 0xF8
            b[o_rec + 8 * r + 1 : o_rec + 8 * r + 5] = bytes(4)
        return bytes(b)

    def zero_sc(b):
        b = bytearray(b)
        b[o_sc : o_sc + n_sc] = bytes(n_sc)
        return bytes(b)

    def zero_t(b):
        b = bytearray(b)
        b[o_temp + 8 : o_temp + 8 + L] = bytes(L)
        return bytes(b)

    out += [("zero_bb", zero_bb(base)), ("zero_sc", zero_sc(base)), ("zero_t", zero_t(base)),
            ("zero_bb_sc", zero_sc(zero_bb(base))), ("zero_sc_t", zero_t(zero_sc(base))), ("zero_all", zero_t(zero_sc(zero_bb(base))))]
    nz = bytearray(zero_bb(base))
    nz[o_rec + 8 * (L - 1) + 4] = 1  # a single non-zero phi bit in the last record
    out.append(("one_phi_bit", bytes(nz)))
    out.append(("all_unk", oracle_encode(unk_chain(), 0, 25)))
    out.append(("bad_magic", b"FCMQ" + base[4:]))
    out.append(("zeros", bytes(len(base))))
    out.append(("empty", b""))
    for cut in (3, 40, o_rec - 1, o_rec + 5, o_sc - 1, o_sc + 1, o_temp - 1, o_temp + 3, size - 1):
        out.append((f"cut{cut}", base[:cut]))
    return out
