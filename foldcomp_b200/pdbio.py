"""Host-side PDB text <-> canonical SoA conversion (the string-keyed part of the reference that
stays on the CPU).

`parse_pdb_chain` mirrors the fixed-column ATOM parser of the reference's CPython module
(foldcomp/foldcomp.cxx:253-293) followed by removeAlternativePosition
(src/atom_coordinate.cpp:362-370) and then maps every residue onto the fixed atom slots of the
FCZ tables the way the reference resolves atoms by NAME at run time:
  * backbone = atoms named N, CA, C (filterBackbone, src/atom_coordinate.cpp:135-143),
  * a side-chain atom missing from the input reads as (0,0,0) (findFirstAtomCoords,
    src/sidechain.cpp:140-147),
  * B-factor of the residue = tempFactor of its CA (src/foldcomp.cpp:543-547),
  * OXT is recognised only as the very last atom (src/foldcomp.cpp:473-481).
`format_pdb` mirrors writeAtomCoordinatesToPDB (src/atom_coordinate.cpp:220-291).
"""
from __future__ import annotations

import numpy as np

from .abi import META_DTYPE, HostChainBatch, concat_chains
from .tables import CODE_UNK, tables


class PdbError(ValueError):
    pass


def _atom_records(pdb_text: str, one_chain: bool = True, dedup_alt: bool = True):
    """(atom, residue, chain, serial, resnum, x, y, z, bfac) for ATOM lines of ONE chain (compress, foldcomp.cxx:253-293);
    one_chain = dedup_alt = False is get_data's reading of a PDB text (getDataFromPDB, foldcomp.cxx:633-656)."""
    out = []
    chain = None
    for line in pdb_text.splitlines():
        if line[:4] != "ATOM":
            continue
        ch = line[21:22]
        if chain is None:
            chain = ch
        if one_chain and ch != chain:
            raise PdbError("Multiple chains found. Please provide a single chain using 'foldcomp.split_pdb_by_chain'")  # foldcomp.cxx:266-268 (flag 2)
        # a short line or a blank / non-numeric fixed-column field is flag 3 of the C++ parser (parsePdbChain,
        # foldcomp_b200/csrc/fcz_db.cpp: n < 22, n < 61, failed numeric field): the same error here
        if len(line) < 61:
            raise PdbError("Malformed ATOM record")
        try:
            rec = (
                line[12:16].strip(" \t"),
                line[17:20].strip(" \t"),
                ch,
                int(line[6:11]),
                int(line[22:26]),
                np.float32(line[30:38]),
                np.float32(line[38:46]),
                np.float32(line[46:54]),
                np.float32(line[60:66]),
            )
        except ValueError:
            raise PdbError("Malformed ATOM record") from None
        out.append(rec)
    if not out:
        raise PdbError("No ATOM lines found")  # foldcomp.cxx:288-290 (flag 1)
    if not dedup_alt:
        return out
    # removeAlternativePosition: drop an atom whose name equals its predecessor's
    dedup = [out[0]]
    for rec in out[1:]:
        if rec[0] == dedup[-1][0]:
            continue
        dedup.append(rec)
    return dedup


def canonicalize(records, title: str) -> HostChainBatch:
    """One chain of atom records -> a 1-chain HostChainBatch in slot order."""
    tb = tables()
    # split by residue number changes (splitAtomByResidue, src/atom_coordinate.cpp:304-328; the
    # last atom always joins the current residue)
    groups = [[records[0]]]
    for i in range(1, len(records)):
        if i != len(records) - 1 and records[i][4] != records[i - 1][4]:
            groups.append([])
        groups[-1].append(records[i])
    res_type, bfac, xyz = [], [], []
    for g in groups:
        code = tb.code(g[0][1])
        if tb.natoms[code] == 0:
            code = CODE_UNK
        res_type.append(code)
        first = {}
        for rec in g:
            first.setdefault(rec[0], rec)
        for name in tb.atom_names[code]:
            rec = first.get(name)
            xyz.append((rec[5], rec[6], rec[7]) if rec is not None else (0.0, 0.0, 0.0))
        ca = first.get("CA")
        bfac.append(ca[8] if ca is not None else 0.0)
    meta = np.zeros(1, META_DTYPE)
    meta["n_atom"] = len(records) & 0xFFFF
    meta["idx_residue"] = records[0][4] & 0xFFFF
    meta["idx_atom"] = records[0][3] & 0xFFFF
    meta["chain"] = ord(records[0][2]) if records[0][2] else ord(" ")
    last = records[-1]
    if last[0] == "OXT":
        meta["has_oxt"] = 1
        meta["oxt"] = (last[5], last[6], last[7])
    tbytes = np.frombuffer(title.encode("latin-1"), np.uint8)
    return concat_chains(
        [(np.array(res_type, np.uint8), np.array(bfac, np.float32), np.array(xyz, np.float32).reshape(-1, 3), tbytes, meta)]
    )


def parse_pdb_chain(pdb_text: str, title: str) -> HostChainBatch:
    return canonicalize(_atom_records(pdb_text), title)


def _ftoa(v: float, T: int, P: int) -> str:
    """fast_ftoa<T,P> (src/atom_coordinate.cpp:186-218) in float32 arithmetic."""
    n = np.float32(v)
    half = np.float32(0.5) / np.float32(T)
    rounded = np.float32(n + (-half if n < 0 else half))
    integer = int(np.trunc(rounded))
    decimal = int(np.trunc(np.float32(np.float32(rounded - np.float32(integer)) * np.float32(T))))
    s = ""
    if n < 0:
        integer, decimal = abs(integer), abs(decimal)
        s = "-"
    return f"{s}{integer}.{decimal:0{P}d}"


def format_pdb(batch: HostChainBatch, c: int = 0, use_alt_order: bool = False) -> str:
    """PDB text of chain `c` of a decoded batch (atoms already in the order the engine produced)."""
    tb = tables()
    title = batch.title(c)
    lines = []
    if title:
        lines.append("TITLE     %s\n" % title[:70])
        rest, cont = title[70:], 2
        while rest:
            lines.append("TITLE  %3d%s\n" % (cont, rest[:70]))
            rest, cont = rest[70:], cont + 1
    m = batch.meta[c]
    chain = chr(int(m["chain"]))
    serial = int(m["idx_atom"])
    r0, r1 = int(batch.res_off[c]), int(batch.res_off[c + 1])
    a = int(batch.atom_off[c])
    last = None
    for r in range(r0, r1):
        code = int(batch.res_type[r])
        names = tb.atom_names[code]
        if use_alt_order:
            names = [names[s] for s in tb.alt[code][: len(names)]]
        resnum = int(m["idx_residue"]) + (r - r0)
        for name in names:
            x, y, z = batch.xyz[a]
            last = (serial, name, tb.name3[code], resnum, x, y, z, batch.bfactor[r])
            lines.append(_atom_line(last, chain))
            serial += 1
            a += 1
    if int(m["has_oxt"]):
        # Foldcomp::read builds OXT with residue_index = nResidue (src/foldcomp.cpp:958-961)
        code = int(batch.res_type[r1 - 1])
        last = (serial, "OXT", tb.name3[code], r1 - r0, m["oxt"][0], m["oxt"][1], m["oxt"][2], batch.bfactor[r1 - 1])
        lines.append(_atom_line(last, chain))
        serial += 1
    if last is not None:
        lines.append("TER   %5d      %3s %s%4d\n" % (last[0] + 1, last[2], chain, last[3]))
    return "".join(lines)


def _atom_line(rec, chain: str) -> str:
    serial, name, res, resnum, x, y, z, b = rec
    nm = "%-4s" % name if len(name) == 4 else " %-3s" % name
    return "ATOM  %5d %s %3s %s%4d    %8s%8s%8s  1.00%6s          %2s  \n" % (
        serial, nm, res, chain, resnum, _ftoa(x, 1000, 3), _ftoa(y, 1000, 3), _ftoa(z, 1000, 3), _ftoa(b, 100, 2), name[0],
    )


def split_pdb_by_chain(pdb_str: str):
    """foldcomp.split_pdb_by_chain (/root/reference/foldcomp/util.py:1-18): the ATOM records of a PDB text grouped into one
    text per run of equal chain ids (column 22), in file order; every other record is dropped.  A text without ATOM records
    gives one empty string, like the reference."""
    import itertools

    atom_lines = [line for line in pdb_str.splitlines() if line.startswith("ATOM")]
    runs = ["".join(l + "\n" for l in grp) for _, grp in itertools.groupby(atom_lines, key=lambda l: l[21])]
    return runs or [""]
