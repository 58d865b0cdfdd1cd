"""Residue tables of the FCZ format as numpy arrays (host-side logic only: PDB text I/O and the
synthetic data generator).  Generated from the same source as csrc/fcz_tables.h; tests/test_tables.py
checks both against a dump of the reference's own table."""
from __future__ import annotations

import functools

import numpy as np

from . import _tables_gen as g

NUM_CODES = 24
MAX_ATOMS = 14
CODE_PRO = 14
CODE_UNK = 23
NAME1 = "ARNDCQEGHILKMFPSTWYVBZ*X"  # src/utility.h:133-205


class Tables:
    def __init__(self):
        self.natoms = np.array(g.NATOMS, np.int32)
        self.name3 = list(g.NAME3)
        self.code_of = {n: c for c, n in enumerate(self.name3)}
        self.atom_names = [list(a) for a in g.ATOMS]
        self.pred = np.zeros((NUM_CODES, MAX_ATOMS, 3), np.int32)
        self.blen = np.zeros((NUM_CODES, MAX_ATOMS), np.float32)
        self.bang = np.zeros((NUM_CODES, MAX_ATOMS), np.float32)
        self.alt = np.zeros((NUM_CODES, MAX_ATOMS), np.int32)
        for c in range(NUM_CODES):
            for j, s in enumerate(g.ALT[c]):
                self.alt[c, j] = s
            for i, (p0, p1, p2, bl, ba) in enumerate(g.BUILD[c]):
                k = i + 3
                self.pred[c, k] = (p0, p1, p2)
                self.blen[c, k] = np.float32(bl)  # double literal -> float, as the reference stores it
                self.bang[c, k] = np.float32(ba)

    def code(self, name3: str) -> int:
        """getOneLetterCode + convertOneLetterCodeToInt (src/utility.cpp:178-232, 379-459):
        unknown three-letter names map to UNK."""
        return self.code_of.get(name3, CODE_UNK)


@functools.lru_cache(maxsize=1)
def tables() -> Tables:
    return Tables()
