"""foldcomp_b200 -- B200-native FCZ encode/decode engine (drop-in for Foldcomp's per-chain codec).

The Python surface mirrors the reference's CPython module for this path
(/root/reference/foldcomp/foldcomp.cxx:702-709):

    compress(name, pdb_content, *, anchor_residue_threshold=25) -> bytes     (foldcomp.cxx:295-328)
    decompress(fcz_bytes) -> (name, pdb_str)                                 (foldcomp.cxx:222-239)
    open(path, *, ids=None, decompress=True, err_on_missing=False)           (foldcomp.cxx:333-433) -> FoldcompDatabase
get_data() (a per-structure dump of the intermediate angles) is not provided.

Both go through the CUDA engine (include/fcz_engine.h); text parsing/formatting is host code
(pdbio.py).  Batch entry points live in `engine.Engine`.  There is no CPU fallback.
"""
from __future__ import annotations

from . import abi
from .abi import HostBlobBatch, HostChainBatch

__all__ = ["compress", "decompress", "open", "FoldcompDatabase", "error", "Engine", "HostChainBatch", "HostBlobBatch"]


class error(Exception):
    """foldcomp.error (foldcomp.cxx:737-741)."""


_engine = None


def _get_engine():
    global _engine
    if _engine is None:
        from .engine import Engine

        _engine = Engine(0)
    return _engine


def __getattr__(name):
    if name == "Engine":
        from .engine import Engine

        return Engine
    if name in ("open", "FoldcompDatabase"):
        from . import database

        return getattr(database, name)
    raise AttributeError(name)


def compress(name: str, pdb_content: str, *, anchor_residue_threshold: int = abi.DEFAULT_ANCHOR_THRESHOLD) -> bytes:
    from .pdbio import PdbError, parse_pdb_chain

    if not isinstance(anchor_residue_threshold, int):
        raise TypeError("anchor_residue_threshold must be an integer")
    try:
        batch = parse_pdb_chain(pdb_content, name)
    except PdbError as e:
        raise error(str(e)) from None
    eng = _get_engine()
    eng.set_opts(anchor_threshold=anchor_residue_threshold)
    out = eng.encode_host(batch)
    if int(out.status[0]) != abi.FCZ_OK:
        raise error("Error compressing")
    return out.blob(0)


def decompress(fcz: bytes):
    eng = _get_engine()
    data = bytes(fcz)
    blobs = HostBlobBatch.from_blobs([data])
    out = eng.decode_to_pdb_host(blobs)  # decode + PDB text on the GPU (k_dec_* then k_pdb_*)
    if int(out.status[0]) != abi.FCZ_OK:
        raise error("Error decompressing.")
    tl = int.from_bytes(data[24:28], "little")  # CompressedFileHeader.lenTitle; the title follows the anchor indices
    t0 = 76 + 4 * data[12]
    return data[t0 : t0 + tl].decode("latin-1"), out.text(0).decode("latin-1")
