"""foldcomp_b200.open(): the FoldcompDatabase of the reference's CPython module (foldcomp/foldcomp.cxx:36-180, 333-433)
over the GPU engine.

    with foldcomp_b200.open("afdb_swissprot_v4", ids=["AF-P12345-F1-model_v4", ...]) as db:
        for name, pdb in db: ...

Files are the reference's (src/database_reader.cpp): `path` (data), `path.index` (key, offset, length per line),
`path.lookup` (key, name, file).  Entries are decoded in BATCHES: reading entry i decodes entries i .. i+batch-1 with one
fcz_decode_to_pdb call and keeps the texts until the window moves, so that iterating a database runs the GPU on
hundreds of chains per launch instead of one.  Like the reference, one byte (the NUL terminator) is dropped from every
entry's indexed length when the raw bytes are handed out (decompress=False; foldcomp.cxx:66,73).  For decoding the entry
goes to the engine at its full indexed length: an FCZ blob carries its own size in its header, so a terminator is ignored
and a database written WITHOUT terminators -- what `foldcomp compress --db` and `fcz_cli compress-db` write
(src/main.cpp:510-517) -- decodes completely (the reference drops the last B-factor byte of such entries)."""
from __future__ import annotations

import builtins
import mmap
import os
import sys

import numpy as np

from . import abi
from .abi import HostBlobBatch


class FoldcompDatabase:
    def __init__(self, path, ids=None, decompress=True, err_on_missing=False, batch=256):
        from . import _get_engine, error

        self._error = error
        path = os.fspath(path)
        if isinstance(path, bytes):
            path = path.decode()
        if ids is not None and not isinstance(ids, list):
            raise TypeError("user_ids must be a list.")
        if not isinstance(decompress, bool):
            raise TypeError("decompress must be a boolean")
        if not isinstance(err_on_missing, bool):
            raise TypeError("err_on_missing must be a boolean")
        rows = np.loadtxt(path + ".index", dtype=np.int64, ndmin=2) if os.path.getsize(path + ".index") else np.zeros((0, 3), np.int64)
        rows = rows[np.argsort(rows[:, 0], kind="stable")]  # the reference's reader sorts its index by key (database_reader.cpp:109)
        self._keys, self._off, self._len = rows[:, 0], rows[:, 1], rows[:, 2]
        self._f = builtins.open(path, "rb")
        self._mm = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ) if os.path.getsize(path) else None
        self._decompress = decompress
        self._batch = max(int(batch), 1)
        self._win = (0, 0, None, None)  # first position, count, HostTextBatch, blobs
        self._engine = _get_engine() if decompress else None
        self._order = None
        if ids:
            by_name = {}
            with builtins.open(path + ".lookup") as lk:
                for line in lk:
                    cols = line.rstrip("\n").split("\t")
                    if len(cols) >= 2:
                        by_name[cols[1]] = int(cols[0])
            id_of_key = {int(k): i for i, k in enumerate(self._keys)}
            order = []
            for name in ids:
                key = by_name.get(name)
                i = id_of_key.get(key) if key is not None else None
                if i is None:
                    msg = f"Skipping entry {name} which is not in the database."
                    if err_on_missing:
                        self.close()
                        raise KeyError(msg)
                    print(msg, file=sys.stderr)
                    continue
                order.append(i)
            self._order = order

    def __len__(self):
        return len(self._order) if self._order is not None else len(self._keys)

    def _entry(self, pos: int, full: bool = False) -> bytes:
        i = self._order[pos] if self._order is not None else pos
        n = int(self._len[i]) if full else max(int(self._len[i]), 1) - 1
        o = int(self._off[i])
        return self._mm[o : o + n]

    def __getitem__(self, index):
        n = len(self)
        if index < 0:
            index += n
        if index < 0 or index >= n:
            raise IndexError("index out of range")
        if not self._decompress:
            return self._entry(index)
        first, count, texts, blobs = self._win
        if texts is None or not (first <= index < first + count):
            first, count = index, min(self._batch, n - index)
            blobs = [self._entry(p, full=True) for p in range(first, first + count)]
            texts = self._engine.decode_to_pdb_host(HostBlobBatch.from_blobs(blobs))
            self._win = (first, count, texts, blobs)
        j = index - first
        blob = blobs[j]
        if int(texts.status[j]) != abi.FCZ_OK:
            raise self._error("Error decompressing: ")
        tl = int.from_bytes(blob[24:28], "little")
        t0 = 76 + 4 * blob[12]
        return blob[t0 : t0 + tl].decode("latin-1"), texts.text(j).decode("latin-1")

    def close(self):
        if getattr(self, "_mm", None) is not None:
            self._mm.close()
            self._mm = None
        if getattr(self, "_f", None) is not None:
            self._f.close()
            self._f = None
        self._win = (0, 0, None, None)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def open(path, *, ids=None, decompress=True, err_on_missing=False):  # noqa: A001 - the reference's name
    """foldcomp.open(path, *, ids=None, decompress=True, err_on_missing=False) -> FoldcompDatabase (foldcomp.cxx:333-433)."""
    return FoldcompDatabase(path, ids=ids, decompress=decompress, err_on_missing=err_on_missing)
