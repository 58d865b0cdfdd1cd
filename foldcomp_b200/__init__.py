"""foldcomp_b200 -- B200-native FCZ encode/decode engine (drop-in for Foldcomp's per-chain codec).

The Python surface mirrors the reference's CPython module for this path
(/root/reference/foldcomp/foldcomp.cxx:702-709):

    compress(name, pdb_content, *, anchor_residue_threshold=25) -> bytes     (foldcomp.cxx:295-328)
    decompress(fcz_bytes) -> (name, pdb_str)                                 (foldcomp.cxx:222-239)
    open(path, *, ids=None, decompress=True, err_on_missing=False)           (foldcomp.cxx:333-433) -> FoldcompDatabase
    get_data(fcz_bytes | pdb_text) -> dict                                   (foldcomp.cxx:497-671)
    split_pdb_by_chain(pdb_text) -> [pdb_text per chain]                     (foldcomp/util.py:1-18)

Both go through the CUDA engine (include/fcz_engine.h); text parsing/formatting is host code
(pdbio.py).  Batch entry points live in `engine.Engine`.  There is no CPU fallback.
"""
from __future__ import annotations

from . import abi
from .abi import HostBlobBatch, HostChainBatch

__all__ = ["compress", "decompress", "get_data", "open", "FoldcompDatabase", "error", "split_pdb_by_chain", "Engine", "HostChainBatch",
           "HostBlobBatch"]


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
    if name == "split_pdb_by_chain":  # pure text helper of the reference package (foldcomp/util.py), same behaviour
        from .pdbio import split_pdb_by_chain

        return split_pdb_by_chain
    if name in ("open", "FoldcompDatabase"):
        from . import database

        return getattr(database, name)
    raise AttributeError(name)


def compress(name: str, pdb_content: str, *, anchor_residue_threshold: int = abi.DEFAULT_ANCHOR_THRESHOLD) -> bytes:
    from . import pdbnative
    from .pdbio import PdbError

    if not isinstance(anchor_residue_threshold, int):
        raise TypeError("anchor_residue_threshold must be an integer")
    # host-side text parsing: the C++ parser of the batch path when it is built, its Python mirror otherwise
    if pdbnative.available():
        parse_pdb_chain = pdbnative.parse_pdb_chain
    else:
        from .pdbio import parse_pdb_chain
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


def _get_data_from_pdb(text: str):
    """getDataFromPDB (foldcomp/foldcomp.cxx:633-671): every ATOM record goes in (no chain check, alternative positions
    kept), Foldcomp::compress runs, and the dict holds the angles BEFORE quantisation plus the input coordinates."""
    from .pdbio import PdbError, _atom_records, canonicalize
    from .tables import NAME1

    try:
        recs = _atom_records(text, one_chain=False, dedup_alt=False)
    except PdbError as e:
        raise ValueError(str(e) + " in PDB file") from None
    batch = canonicalize(recs, "")
    eng = _get_engine()
    ang = eng.backbone_angles_host(batch)  # k_raw_angles: the encoder's own arithmetic, before quantisation
    L = len(ang)
    tors = ang[: L - 1, :3].reshape(-1)
    bond = ang[:, 3:].reshape(-1)[1 : 3 * L - 1]
    return {
        "phi": [float(v) for v in tors[2::3]], "psi": [float(v) for v in tors[0::3]], "omega": [float(v) for v in tors[1::3]],
        "torsion_angles": [float(v) for v in tors], "bond_angles": [float(v) for v in bond],
        "residues": "".join(NAME1[int(c)] for c in batch.res_type),
        "b_factors": [float(v) for v in batch.bfactor],
        "coordinates": [(float(r[5]), float(r[6]), float(r[7])) for r in recs],
    }


def get_data(input):  # noqa: A002 - the reference's parameter name
    """foldcomp.get_data(input) -> dict with phi, psi, omega, torsion_angles, bond_angles, residues, b_factors,
    coordinates (foldcomp/foldcomp.cxx:497-640).  FCZ bytes: the continuised angles of the blob and its decoded
    coordinates (getDataFromFCZ); PDB text: the encoder's angles before quantisation and the input coordinates
    (getDataFromPDB, 633-671).  Both from the GPU engine."""
    data = input.encode("latin-1") if isinstance(input, str) else bytes(input)
    if len(data) == 0:
        raise ValueError("Input is empty")
    if data[:4] != b"FCMP":
        return _get_data_from_pdb(data.decode("latin-1"))
    from .tables import NAME1

    eng = _get_engine()
    blobs = HostBlobBatch.from_blobs([data])
    dec = eng.decode_host(blobs)
    if int(dec.status[0]) != abi.FCZ_OK:
        raise ValueError("Could not decompress FCZ file")
    _, ang = eng.unpack_angles_host(blobs)
    phi, psi, omega = ang[:, 0], ang[:, 1], ang[:, 2]
    n_ca_c, ca_c_n, c_n_ca = ang[:, 3], ang[:, 4], ang[:, 5]
    L = len(ang)
    tors = [float(v) for i in range(L - 1) for v in (psi[i], omega[i], phi[i])]  # src/foldcomp.cpp:788-793
    bond = [float(v) for i in range(L) for v in (ca_c_n[i], c_n_ca[i], n_ca_c[i])]  # src/foldcomp.cpp:799-804
    coords = [tuple(float(x) for x in row) for row in dec.xyz]
    m = dec.meta[0]
    if int(m["has_oxt"]):
        coords.append(tuple(float(x) for x in m["oxt"]))
    return {
        "phi": [float(v) for v in phi], "psi": [float(v) for v in psi], "omega": [float(v) for v in omega],
        "torsion_angles": tors, "bond_angles": bond,
        "residues": "".join(NAME1[int(c)] for c in dec.res_type),
        "b_factors": [float(v) for v in dec.bfactor], "coordinates": coords,
    }
