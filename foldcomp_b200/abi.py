"""ctypes mirror of include/fcz_engine.h (structs, constants) plus host-side batch containers.

Pure host logic: nothing here touches the GPU.  `HostChainBatch` / `HostBlobBatch` hold numpy
arrays in the canonical SoA layout and can expose themselves as `fcz_chain_batch` /
`fcz_blob_batch` structs pointing at host memory.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

FCZ_MEM_HOST = 0
FCZ_MEM_DEVICE = 1

FCZ_OK = 0
FCZ_E_MAGIC = -1
FCZ_E_TRUNCATED = -2
FCZ_E_RESIDUE = -3
FCZ_E_LIMIT = -4
FCZ_E_CAPACITY = -5
FCZ_E_CUDA = -6
FCZ_E_ARG = -7
# PDB text parser, per entry (the flags of foldcomp/foldcomp.cxx:262-291 plus the two shapes the reference throws on)
FCZ_E_PARSE_NOATOM = -11
FCZ_E_PARSE_CHAINS = -12
FCZ_E_PARSE_RECORD = -13
FCZ_E_PARSE_NUMBER = -14
FCZ_E_PARSE_GAPS = -15

DEFAULT_ANCHOR_THRESHOLD = 25  # src/foldcomp.h:56

# ValidityError classes (src/foldcomp.h:59-67) as fcz_check_batch reports them, with printValidityError's messages
# (src/foldcomp.cpp:1534-1561)
VALIDITY = (
    "SUCCESS",
    "E_BACKBONE_COUNT_MISMATCH",
    "E_SIDECHAIN_COUNT_MISMATCH",
    "E_TEMP_FACTOR_COUNT_MISMATCH",
    "E_EMPTY_BACKBONE_ANGLE",
    "E_EMPTY_SIDECHAIN_ANGLE",
    "E_EMPTY_TEMP_FACTOR",
)
VALIDITY_MESSAGE = (
    "",
    "[Error] Number of backbone angles does not match header: ",
    "[Error] Number of sidechain angles does not match header: ",
    "[Error] Number of temperature factors does not match header: ",
    "[Error] All backbone angles are empty: ",
    "[Error] All sidechain angles are empty: ",
    "[Error] All temperature factors are empty: ",
)


class FczOpts(C.Structure):
    _fields_ = [("anchor_threshold", C.c_int32), ("use_alt_atom_order", C.c_int32), ("stream", C.c_void_p), ("terminate_blobs", C.c_int32)]


class FczChainMeta(C.Structure):
    _fields_ = [
        ("n_atom", C.c_uint16),
        ("idx_residue", C.c_uint16),
        ("idx_atom", C.c_uint16),
        ("chain", C.c_uint8),
        ("has_oxt", C.c_uint8),
        ("oxt", C.c_float * 3),
    ]


META_DTYPE = np.dtype(
    [
        ("n_atom", "<u2"),
        ("idx_residue", "<u2"),
        ("idx_atom", "<u2"),
        ("chain", "u1"),
        ("has_oxt", "u1"),
        ("oxt", "<f4", (3,)),
    ]
)
assert META_DTYPE.itemsize == C.sizeof(FczChainMeta) == 20


class FczChainBatch(C.Structure):
    _fields_ = [
        ("n_chains", C.c_uint32),
        ("mem", C.c_int32),
        ("res_off", C.c_void_p),
        ("atom_off", C.c_void_p),
        ("title_off", C.c_void_p),
        ("res_type", C.c_void_p),
        ("bfactor", C.c_void_p),
        ("xyz", C.c_void_p),
        ("titles", C.c_void_p),
        ("meta", C.c_void_p),
        ("status", C.c_void_p),
        ("res_cap", C.c_uint64),
        ("atom_cap", C.c_uint64),
        ("title_cap", C.c_uint64),
    ]


class FczBlobBatch(C.Structure):
    _fields_ = [
        ("n_chains", C.c_uint32),
        ("mem", C.c_int32),
        ("blob_off", C.c_void_p),
        ("bytes", C.c_void_p),
        ("status", C.c_void_p),
        ("bytes_cap", C.c_uint64),
    ]


class FczTextBatch(C.Structure):
    _fields_ = [
        ("n_chains", C.c_uint32),
        ("mem", C.c_int32),
        ("text_off", C.c_void_p),
        ("bytes", C.c_void_p),
        ("bytes_cap", C.c_uint64),
        ("status", C.c_void_p),
    ]


class FczSizes(C.Structure):
    _fields_ = [
        ("n_res", C.c_uint64),
        ("n_atoms", C.c_uint64),
        ("n_title_bytes", C.c_uint64),
        ("n_blob_bytes", C.c_uint64),
    ]


class FczProfile(C.Structure):
    _fields_ = [
        ("encode_kernel_ms", C.c_double),
        ("decode_kernel_ms", C.c_double),
        ("encode_launches", C.c_uint64),
        ("decode_launches", C.c_uint64),
        ("kernel_ms", C.c_double * 8),
        ("kernel_launches", C.c_uint64 * 8),
    ]


PROF_KINDS = {"encode_span": 0, "decode_span": 1, "k_encode": 2, "k_dec_front": 3, "k_dec_stitch_t": 4, "k_dec_back": 5}
PROF_TEXT_KINDS = {"k_pdb_plan": 6, "k_pdb_emit": 7}


def _ptr(a: np.ndarray | None):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


@dataclass
class HostChainBatch:
    """Chains in canonical slot order, host memory (numpy)."""

    res_off: np.ndarray  # uint32 [n+1]
    atom_off: np.ndarray  # uint64 [n+1]
    title_off: np.ndarray  # uint32 [n+1]
    res_type: np.ndarray  # uint8 [R]
    bfactor: np.ndarray  # float32 [R]
    xyz: np.ndarray  # float32 [A,3]
    titles: np.ndarray  # uint8 [T]
    meta: np.ndarray  # META_DTYPE [n]
    status: np.ndarray = field(default=None)  # int32 [n]

    def __post_init__(self):
        if self.status is None:
            self.status = np.zeros(self.n_chains, np.int32)

    @property
    def n_chains(self) -> int:
        return len(self.res_off) - 1

    @property
    def n_res(self) -> int:
        return int(self.res_off[-1])

    @property
    def n_atoms(self) -> int:
        return int(self.atom_off[-1])

    def title(self, c: int) -> str:
        return bytes(self.titles[self.title_off[c] : self.title_off[c + 1]]).decode("latin-1")

    def chain(self, c: int) -> "HostChainBatch":
        return self.select([c])

    def select(self, idx) -> "HostChainBatch":
        idx = list(idx)
        parts = [self._slice(c) for c in idx]
        return concat_chains(parts)

    def _slice(self, c: int):
        r0, r1 = int(self.res_off[c]), int(self.res_off[c + 1])
        a0, a1 = int(self.atom_off[c]), int(self.atom_off[c + 1])
        t0, t1 = int(self.title_off[c]), int(self.title_off[c + 1])
        return (
            self.res_type[r0:r1],
            self.bfactor[r0:r1],
            self.xyz[a0:a1],
            self.titles[t0:t1],
            self.meta[c : c + 1],
        )

    def as_struct(self) -> FczChainBatch:
        s = FczChainBatch()
        s.n_chains = self.n_chains
        s.mem = FCZ_MEM_HOST
        s.res_off = _ptr(self.res_off)
        s.atom_off = _ptr(self.atom_off)
        s.title_off = _ptr(self.title_off)
        s.res_type = _ptr(self.res_type)
        s.bfactor = _ptr(self.bfactor)
        s.xyz = _ptr(self.xyz)
        s.titles = _ptr(self.titles)
        s.meta = _ptr(self.meta)
        s.status = _ptr(self.status)
        s.res_cap = len(self.res_type)
        s.atom_cap = len(self.xyz)
        s.title_cap = len(self.titles)
        return s

    @staticmethod
    def empty(n_chains: int, n_res: int = 0, n_atoms: int = 0, n_title: int = 0) -> "HostChainBatch":
        return HostChainBatch(
            res_off=np.zeros(n_chains + 1, np.uint32),
            atom_off=np.zeros(n_chains + 1, np.uint64),
            title_off=np.zeros(n_chains + 1, np.uint32),
            res_type=np.zeros(n_res, np.uint8),
            bfactor=np.zeros(n_res, np.float32),
            xyz=np.zeros((n_atoms, 3), np.float32),
            titles=np.zeros(max(n_title, 1), np.uint8)[:n_title],
            meta=np.zeros(n_chains, META_DTYPE),
        )


def concat_chains(parts) -> HostChainBatch:
    """parts: iterable of (res_type, bfactor, xyz, title_bytes, meta[1])"""
    parts = list(parts)
    n = len(parts)
    res_off = np.zeros(n + 1, np.uint32)
    atom_off = np.zeros(n + 1, np.uint64)
    title_off = np.zeros(n + 1, np.uint32)
    for i, p in enumerate(parts):
        res_off[i + 1] = res_off[i] + len(p[0])
        atom_off[i + 1] = atom_off[i] + np.uint64(len(p[2]))
        title_off[i + 1] = title_off[i] + len(p[3])
    cat = lambda k, dt, shp: (
        np.ascontiguousarray(np.concatenate([np.asarray(p[k], dt).reshape(shp) for p in parts]))
        if n
        else np.zeros(shp if shp != (-1,) else (0,), dt)
    )
    return HostChainBatch(
        res_off=res_off,
        atom_off=atom_off,
        title_off=title_off,
        res_type=cat(0, np.uint8, (-1,)),
        bfactor=cat(1, np.float32, (-1,)),
        xyz=cat(2, np.float32, (-1, 3)),
        titles=cat(3, np.uint8, (-1,)),
        meta=np.ascontiguousarray(np.concatenate([p[4] for p in parts])) if n else np.zeros(0, META_DTYPE),
    )


def concat_batches(batches) -> HostChainBatch:
    """Concatenate whole batches (vectorised; offsets are re-based)."""
    batches = list(batches)
    if len(batches) == 1:
        return batches[0]
    if not batches:
        return HostChainBatch.empty(0)

    def offs(name, dt):
        out, base = [np.zeros(1, dt)], 0
        for b in batches:
            o = getattr(b, name).astype(np.int64)
            out.append((o[1:] + base).astype(dt))
            base += int(o[-1])
        return np.concatenate(out)

    cat = lambda name: np.ascontiguousarray(np.concatenate([getattr(b, name) for b in batches]))
    return HostChainBatch(
        res_off=offs("res_off", np.uint32), atom_off=offs("atom_off", np.uint64), title_off=offs("title_off", np.uint32),
        res_type=cat("res_type"), bfactor=cat("bfactor"), xyz=cat("xyz"), titles=cat("titles"), meta=cat("meta"),
    )


@dataclass
class HostBlobBatch:
    blob_off: np.ndarray  # uint64 [n+1]
    bytes: np.ndarray  # uint8 [cap]
    status: np.ndarray = field(default=None)

    def __post_init__(self):
        if self.status is None:
            self.status = np.zeros(self.n_chains, np.int32)

    @property
    def n_chains(self) -> int:
        return len(self.blob_off) - 1

    def blob(self, c: int) -> bytes:
        return bytes(self.bytes[int(self.blob_off[c]) : int(self.blob_off[c + 1])])

    def blobs(self):
        return [self.blob(c) for c in range(self.n_chains)]

    def as_struct(self) -> FczBlobBatch:
        s = FczBlobBatch()
        s.n_chains = self.n_chains
        s.mem = FCZ_MEM_HOST
        s.blob_off = _ptr(self.blob_off)
        s.bytes = _ptr(self.bytes)
        s.status = _ptr(self.status)
        s.bytes_cap = len(self.bytes)
        return s

    @staticmethod
    def empty(n_chains: int, cap: int) -> "HostBlobBatch":
        return HostBlobBatch(np.zeros(n_chains + 1, np.uint64), np.zeros(cap, np.uint8))

    @staticmethod
    def from_blobs(blobs) -> "HostBlobBatch":
        blobs = list(blobs)
        off = np.zeros(len(blobs) + 1, np.uint64)
        for i, b in enumerate(blobs):
            off[i + 1] = off[i] + np.uint64(len(b))
        data = np.frombuffer(b"".join(blobs), np.uint8).copy() if blobs else np.zeros(0, np.uint8)
        return HostBlobBatch(off, data)


def encode_bound(n_chains: int, n_res: int, n_atoms: int, n_title: int, anchor_threshold: int) -> int:
    """Same arithmetic as fcz_encode_bound (SURVEY.md Appendix A size formula)."""
    b = max(int(anchor_threshold), 1)
    return 98 * n_chains + 40 * (n_res // b + 2 * n_chains) + n_title + 6 * n_res + n_atoms


@dataclass
class HostTextBatch:
    """Texts (PDB text per chain / extract output per blob), tightly concatenated, host memory."""

    text_off: np.ndarray  # uint64 [n+1]
    bytes: np.ndarray  # uint8 [cap]
    status: np.ndarray = field(default=None)  # int32 [n], written by decode_to_pdb

    def __post_init__(self):
        if self.status is None:
            self.status = np.zeros(self.n_chains, np.int32)

    @property
    def n_chains(self) -> int:
        return len(self.text_off) - 1

    def text(self, c: int) -> bytes:
        return bytes(self.bytes[int(self.text_off[c]) : int(self.text_off[c + 1])])

    def as_struct(self) -> FczTextBatch:
        s = FczTextBatch()
        s.n_chains = self.n_chains
        s.mem = FCZ_MEM_HOST
        s.text_off = _ptr(self.text_off)
        s.bytes = _ptr(self.bytes)
        s.bytes_cap = len(self.bytes)
        s.status = _ptr(self.status)
        return s
