"""Synthetic protein chains in canonical slot order (BASELINE.json configs 2-5).

Host-side numpy, vectorised across chains.  Residue types are i.i.d. from Swiss-Prot
frequencies (mean 7.82 heavy atoms per residue); the backbone is grown by NeRF with the ideal
bond lengths of the format (src/foldcomp.h:51-54), bond angles ~ N(ideal, 2 deg)
(src/nerf.h:44-47), (phi,psi) from a three-basin Ramachandran mixture, omega ~ N(180, 5 deg) with
0.3 % cis; side chains use the table lengths/angles with uniform torsions (CB is kept
~120 deg from N so that no predecessor triple is collinear); B-factors are uniform in
[30, 98] with two decimals; coordinates are rounded to 0.001 A (PDB precision); every chain ends
with an OXT, is chain 'A', numbered from 1 and titled syn_%07d.
"""
from __future__ import annotations

import numpy as np

from .abi import META_DTYPE, HostChainBatch
from .tables import tables

SEED = 20260925

# Swiss-Prot composition (percent), order ARNDCQEGHILKMFPSTWYV = codes 0..19
_FREQ = np.array([8.25, 5.53, 4.06, 5.45, 1.37, 3.93, 6.75, 7.07, 2.27, 5.96, 9.66, 5.84, 2.42, 3.86, 4.70, 6.56, 5.34, 1.08, 2.92, 6.87])


def _place(a, b, c, length, angle_deg, torsion_deg):
    """Vectorised NeRF placement in float64: a,b,c [...,3]."""
    bc = c - b
    bcn = bc / np.linalg.norm(bc, axis=-1, keepdims=True)
    n = np.cross(b - a, bcn)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    nbc = np.cross(n, bcn)
    shape = a.shape[:-1]
    th = np.broadcast_to(np.deg2rad(angle_deg), shape)
    ta = np.broadcast_to(np.deg2rad(torsion_deg), shape)
    length = np.broadcast_to(length, shape)
    d = np.stack([-length * np.cos(th), length * np.cos(ta) * np.sin(th), length * np.sin(ta) * np.sin(th)], -1)
    return c + bcn * d[..., 0:1] + nbc * d[..., 1:2] + n * d[..., 2:3]


def _dihedral(a, b, c, d):
    b1, b2, b3 = b - a, c - b, d - c
    n1, n2 = np.cross(b1, b2), np.cross(b2, b3)
    m = np.cross(n1, b2 / np.linalg.norm(b2, axis=-1, keepdims=True))
    # sign chosen so that _dihedral(a, b, c, _place(a, b, c, l, th, tau)) == tau
    return -np.rad2deg(np.arctan2((m * n2).sum(-1), (n1 * n2).sum(-1)))


def mixed_lengths(rng: np.random.Generator, n: int, lo: int = 50, hi: int = 2000) -> np.ndarray:
    """Clipped log-normal (median 280, sigma_ln 0.75): the AFDB-UniProt proxy of config 5."""
    return np.clip(np.rint(np.exp(rng.normal(np.log(280.0), 0.75, n))), lo, hi).astype(np.int64)


CHUNK = 500  # chains generated per vectorised pass (bounds the float64 temporaries to ~100 MB)


def generate(n_chains: int, length=350, seed: int = SEED, first_index: int = 0) -> HostChainBatch:
    """`length` is an int (all chains equal) or an array of per-chain lengths.  Chains are generated in
    chunks of CHUNK with a Philox stream keyed by (seed, index of the chunk's first chain)."""
    from .abi import concat_batches

    lens = np.full(n_chains, int(length), np.int64) if np.isscalar(length) else np.asarray(length, np.int64)
    assert len(lens) == n_chains and (n_chains == 0 or lens.min() >= 2)
    if n_chains == 0 or lens.min() == lens.max():
        parts = [_generate_chunk(lens[s : s + CHUNK], seed, first_index + s) for s in range(0, n_chains, CHUNK)]
        return concat_batches(parts)
    # ragged: a chunk costs its LONGEST chain, so chunks are cut from the length-sorted order and the chains put
    # back in the requested order afterwards (titles follow the final position)
    order = np.argsort(lens, kind="stable")
    parts = [_generate_chunk(lens[order[s : s + CHUNK]], seed, first_index + s) for s in range(0, n_chains, CHUNK)]
    inv = np.empty(n_chains, np.int64)
    inv[order] = np.arange(n_chains)
    out = concat_batches(parts).select(inv)
    out.titles = np.frombuffer(b"".join(b"syn_%07d" % (first_index + i) for i in range(n_chains)), np.uint8).copy()
    return out


def _generate_chunk(lens: np.ndarray, seed: int, first_index: int) -> HostChainBatch:
    tb = tables()
    n_chains = len(lens)
    rng = np.random.Generator(np.random.Philox(key=seed + 7919 * first_index))
    Lm = int(lens.max())
    n = n_chains
    types = rng.choice(20, size=(n, Lm), p=_FREQ / _FREQ.sum()).astype(np.uint8)

    # backbone torsions
    basin = rng.choice(3, size=(n, Lm), p=[0.45, 0.35, 0.20])
    phi0 = np.array([-63.0, -120.0, -75.0])[basin]
    psi0 = np.array([-43.0, 130.0, 145.0])[basin]
    phi = phi0 + rng.normal(0, 15.0, (n, Lm))
    psi = psi0 + rng.normal(0, 15.0, (n, Lm))
    omega = 180.0 + rng.normal(0, 5.0, (n, Lm))
    omega = np.where(rng.random((n, Lm)) < 0.003, rng.normal(0, 5.0, (n, Lm)), omega)
    ang_ncac = rng.normal(111.2812, 2.0, (n, Lm))
    ang_cacn = rng.normal(116.6429, 2.0, (n, Lm))
    ang_cnca = rng.normal(121.3822, 2.0, (n, Lm))

    N = np.zeros((n, Lm, 3))
    CA = np.zeros((n, Lm, 3))
    Cc = np.zeros((n, Lm, 3))
    # first residue: a random rigid placement
    N[:, 0] = rng.normal(0, 5.0, (n, 3))
    d1 = rng.normal(0, 1, (n, 3))
    d1 /= np.linalg.norm(d1, axis=1, keepdims=True)
    CA[:, 0] = N[:, 0] + 1.4581 * d1
    helper = N[:, 0] + rng.normal(0, 1, (n, 3)) * 3.0 + 1.0
    Cc[:, 0] = _place(helper, N[:, 0], CA[:, 0], 1.5281, ang_ncac[:, 0], rng.uniform(-180, 180, n))
    for r in range(1, Lm):
        N[:, r] = _place(N[:, r - 1], CA[:, r - 1], Cc[:, r - 1], 1.3311, ang_cacn[:, r - 1], psi[:, r - 1])
        CA[:, r] = _place(CA[:, r - 1], Cc[:, r - 1], N[:, r], 1.4581, ang_cnca[:, r - 1], omega[:, r - 1])
        Cc[:, r] = _place(Cc[:, r - 1], N[:, r], CA[:, r], 1.5281, ang_ncac[:, r], phi[:, r])

    # per-residue atom slots [n, Lm, 14, 3]
    slots = np.zeros((n, Lm, 14, 3))
    slots[:, :, 0], slots[:, :, 1], slots[:, :, 2] = N, CA, Cc
    natoms = tb.natoms[types]  # [n, Lm]
    sc_tor = rng.uniform(-180.0, 180.0, (n, Lm, 14))
    flat = slots.reshape(-1, 14, 3)
    ftypes = types.reshape(-1)
    fnat = natoms.reshape(-1)
    ftor = sc_tor.reshape(-1, 14)
    for k in range(3, 14):
        sel = np.nonzero(fnat > k)[0]
        if len(sel) == 0:
            continue
        t = ftypes[sel]
        p = tb.pred[t, k]  # [m,3]
        a = flat[sel, p[:, 0]]
        b = flat[sel, p[:, 1]]
        c = flat[sel, p[:, 2]]
        tor = ftor[sel, k]
        if k == 4:
            # CB is built from (O, C, CA): keep it ~120 deg away from N around the C-CA axis (L chirality),
            # otherwise a uniform torsion can drop CB onto N and make N-CA-CB collinear, an input on
            # which the reference's own reconstruction is numerically unstable.
            tor = _dihedral(a, b, c, flat[sel, 0]) - 120.0 + rng.normal(0, 5.0, len(sel))
            tor = (tor + 180.0) % 360.0 - 180.0
        flat[sel, k] = _place(a, b, c, tb.blen[t, k].astype(np.float64), tb.bang[t, k].astype(np.float64), tor)
    slots = flat.reshape(n, Lm, 14, 3)

    # pack chains (ragged)
    valid_res = np.arange(Lm)[None, :] < lens[:, None]  # [n, Lm]
    atom_valid = valid_res[:, :, None] & (np.arange(14)[None, None, :] < natoms[:, :, None])
    xyz = np.round(slots[atom_valid], 3).astype(np.float32)
    res_type = types[valid_res]
    bfac = (np.round(rng.uniform(30.0, 98.0, (n, Lm)), 2)).astype(np.float32)[valid_res]
    res_off = np.zeros(n + 1, np.uint32)
    res_off[1:] = np.cumsum(lens)
    atoms_per_chain = (natoms * valid_res).sum(axis=1)
    atom_off = np.zeros(n + 1, np.uint64)
    atom_off[1:] = np.cumsum(atoms_per_chain)

    # OXT from (N, CA, C) of the last residue, opposite to O
    li = lens - 1
    rows = np.arange(n)
    oxt = _place(N[rows, li], CA[rows, li], Cc[rows, li], 1.25, 118.0, sc_tor[rows, li, 3] + 180.0)
    meta = np.zeros(n, META_DTYPE)
    meta["n_atom"] = ((atoms_per_chain + 1) & 0xFFFF).astype(np.uint16)
    meta["idx_residue"] = 1
    meta["idx_atom"] = 1
    meta["chain"] = ord("A")
    meta["has_oxt"] = 1
    meta["oxt"] = np.round(oxt, 3).astype(np.float32)

    titles = b"".join(b"syn_%07d" % (first_index + i) for i in range(n))
    title_off = np.arange(n + 1, dtype=np.uint32) * 11
    return HostChainBatch(
        res_off=res_off,
        atom_off=atom_off,
        title_off=title_off,
        res_type=np.ascontiguousarray(res_type),
        bfactor=np.ascontiguousarray(bfac),
        xyz=np.ascontiguousarray(xyz),
        titles=np.frombuffer(titles, np.uint8).copy(),
        meta=meta,
    )
