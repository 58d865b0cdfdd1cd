"""foldcomp-db files written/read in plain Python (test infrastructure): the format of the reference's
src/database_writer.cpp:36-96 -- data file, `.index` (key, offset, length per line), `.lookup` (key, name, file),
`.dbtype` (int32 12)."""
import ctypes as C
import os
import struct
import sys

import helpers as H

GPU_SO = os.path.join(H.ROOT, "foldcomp_b200", "csrc", "libfoldcomp_gpu.so")
PYREF = os.path.join(H.ROOT, "oracle", "_ref", "pyref")


def gpu_host_lib():
    lib = C.CDLL(GPU_SO)
    lib.fczgpu_parse_pdb.restype = C.c_int
    lib.fczgpu_parse_pdb.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32),
                                     C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32]
    lib.fczgpu_parse_float.restype = C.c_float
    lib.fczgpu_parse_float.argtypes = [C.c_char_p, C.c_size_t]
    lib.fczgpu_db_copy.restype = C.c_int
    lib.fczgpu_db_copy.argtypes = [C.c_char_p, C.c_char_p]
    for f in (lib.fczgpu_decompress_db, lib.fczgpu_compress_db):
        f.restype = C.c_int
        f.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_double)]
    return lib


def reference_module():
    """The reference's own CPython module compiled into oracle/_ref/pyref (None when it is not there)."""
    if not os.path.exists(os.path.join(PYREF, "foldcomp.so")):
        return None
    if PYREF not in sys.path:
        sys.path.insert(0, PYREF)
    import foldcomp

    return foldcomp


def write_db(path, entries, nul=True):
    """entries: list of (key, name, payload bytes), written in the given order."""
    off = 0
    with open(path, "wb") as d, open(path + ".index", "w") as ix, open(path + ".lookup", "w") as lk:
        for key, name, data in entries:
            blob = data + (b"\0" if nul else b"")
            d.write(blob)
            ix.write(f"{key}\t{off}\t{len(blob)}\n")
            lk.write(f"{key}\t{name}\t0\n")
            off += len(blob)
    with open(path + ".dbtype", "wb") as t:
        t.write(struct.pack("<i", 12))


def fcz_size(blob: bytes) -> int:
    """Size of an FCZ blob from its header (Foldcomp::getSize, src/foldcomp.cpp:1190-1214), 0 if it is not one."""
    if len(blob) < 76 or blob[:4] != b"FCMP":
        return 0
    L, n_anchor = int.from_bytes(blob[4:6], "little"), blob[12]
    n_sc, title_len = int.from_bytes(blob[16:20], "little"), int.from_bytes(blob[24:28], "little")
    return 76 + 40 * n_anchor + title_len + 13 + 8 * L + n_sc + 8 + L


def read_db(path):
    """-> list of (key, name, payload bytes) in index order; a trailing NUL terminator is stripped -- for an FCZ entry only
    when the header says the blob is one byte shorter (`foldcomp compress --db` writes its blobs WITHOUT a terminator,
    src/main.cpp:510-517, and a blob may end in a zero B-factor byte)."""
    data = open(path, "rb").read()
    names = {}
    if os.path.exists(path + ".lookup"):
        for line in open(path + ".lookup"):
            k, n, _ = line.rstrip("\n").split("\t")
            names[int(k)] = n
    out = []
    for line in open(path + ".index"):
        k, o, n = (int(x) for x in line.split())
        blob = data[o : o + n]
        if blob.endswith(b"\0") and (blob[:4] != b"FCMP" or fcz_size(blob) == len(blob) - 1):
            blob = blob[:-1]
        out.append((k, names.get(k, str(k)), blob))
    return out
