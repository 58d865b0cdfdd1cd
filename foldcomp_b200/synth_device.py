"""The synthetic-chain generator of foldcomp_b200/synth.py on the GPU (torch), for the configurations that need millions
of DISTINCT chains (BASELINE.json configs[3]: 2 M chains over 8 GPUs; SURVEY.md 8d "generated on device with the same
generator").  Same model -- Swiss-Prot residue frequencies, NeRF-grown backbone with the format's ideal bond lengths,
N(ideal, 2 deg) bond angles, three-basin Ramachandran mixture, omega ~ N(180, 5 deg) with 0.3 % cis, table side chains
with uniform torsions (CB kept ~120 deg from N), B-factors uniform in [30, 98] with two decimals, coordinates rounded to
0.001 A, OXT, chain 'A', titles syn_%07d -- but torch's device Philox streams instead of numpy's, so the chains are not
the same NUMBERS as synth.generate's.  Equal-length chains only.  Test / benchmark infrastructure: nothing on the
product path imports this."""
from __future__ import annotations

import numpy as np
import torch

from .synth import _FREQ
from .tables import tables


def _norm(v):
    return v / torch.linalg.vector_norm(v, dim=-1, keepdim=True)


def _place(a, b, c, length, angle_deg, torsion_deg):
    bcn = _norm(c - b)
    n = _norm(torch.cross(b - a, bcn, dim=-1))
    nbc = torch.cross(n, bcn, dim=-1)
    th, ta = torch.deg2rad(angle_deg), torch.deg2rad(torsion_deg)
    d0, d1, d2 = -length * torch.cos(th), length * torch.cos(ta) * torch.sin(th), length * torch.sin(ta) * torch.sin(th)
    return c + bcn * d0[..., None] + nbc * d1[..., None] + n * d2[..., None]


def _dihedral(a, b, c, d):
    b1, b2, b3 = b - a, c - b, d - c
    n1, n2 = torch.cross(b1, b2, dim=-1), torch.cross(b2, b3, dim=-1)
    m = torch.cross(n1, _norm(b2), dim=-1)
    return -torch.rad2deg(torch.atan2((m * n2).sum(-1), (n1 * n2).sum(-1)))


def generate_device(n_chains: int, length: int, seed: int, first_index: int, device, chunk: int = 20000):
    """-> dict of device tensors in the canonical layout (res_off int32, atom_off int64, title_off int32, res_type u8,
    bfactor f32, xyz f32 [A,3], titles u8, meta u8 [n,20]) for n_chains chains of `length` residues."""
    tb = tables()
    dev = torch.device(device)
    f64 = torch.float64
    natoms_t = torch.from_numpy(tb.natoms.astype(np.int64)).to(dev)
    pred_t = torch.from_numpy(tb.pred.astype(np.int64)).to(dev)      # [codes, 14, 3]
    blen_t = torch.from_numpy(tb.blen.astype(np.float64)).to(dev)
    bang_t = torch.from_numpy(tb.bang.astype(np.float64)).to(dev)
    cdf = torch.from_numpy(np.cumsum(_FREQ / _FREQ.sum())).to(dev)
    L = int(length)
    out = {k: [] for k in ("res_type", "bfactor", "xyz", "meta", "atoms_per_chain")}
    for s in range(0, n_chains, chunk):
        n = min(chunk, n_chains - s)
        g = torch.Generator(device=dev)
        g.manual_seed(int(seed) * 1000003 + 7919 * (first_index + s))
        U = lambda *shape: torch.rand(*shape, generator=g, device=dev, dtype=f64)
        Nrm = lambda mu, sd, *shape: mu + sd * torch.randn(*shape, generator=g, device=dev, dtype=f64)
        types = torch.searchsorted(cdf, U(n, L)).clamp_(max=19)
        basin = torch.searchsorted(torch.tensor([0.45, 0.80, 1.0], device=dev, dtype=f64), U(n, L)).clamp_(max=2)
        phi = torch.tensor([-63.0, -120.0, -75.0], device=dev, dtype=f64)[basin] + Nrm(0, 15.0, n, L)
        psi = torch.tensor([-43.0, 130.0, 145.0], device=dev, dtype=f64)[basin] + Nrm(0, 15.0, n, L)
        omega = torch.where(U(n, L) < 0.003, Nrm(0, 5.0, n, L), Nrm(180.0, 5.0, n, L))
        ang_ncac, ang_cacn, ang_cnca = Nrm(111.2812, 2.0, n, L), Nrm(116.6429, 2.0, n, L), Nrm(121.3822, 2.0, n, L)
        N = torch.zeros(n, L, 3, device=dev, dtype=f64)
        CA, C = torch.zeros_like(N), torch.zeros_like(N)
        N[:, 0] = Nrm(0, 5.0, n, 3)
        CA[:, 0] = N[:, 0] + 1.4581 * _norm(Nrm(0, 1.0, n, 3))
        helper = N[:, 0] + Nrm(0, 1.0, n, 3) * 3.0 + 1.0
        C[:, 0] = _place(helper, N[:, 0], CA[:, 0], torch.tensor(1.5281, device=dev, dtype=f64), ang_ncac[:, 0], U(n) * 360.0 - 180.0)
        l_cn, l_nca, l_cac = (torch.tensor(v, device=dev, dtype=f64) for v in (1.3311, 1.4581, 1.5281))
        for r in range(1, L):
            N[:, r] = _place(N[:, r - 1], CA[:, r - 1], C[:, r - 1], l_cn, ang_cacn[:, r - 1], psi[:, r - 1])
            CA[:, r] = _place(CA[:, r - 1], C[:, r - 1], N[:, r], l_nca, ang_cnca[:, r - 1], omega[:, r - 1])
            C[:, r] = _place(C[:, r - 1], N[:, r], CA[:, r], l_cac, ang_ncac[:, r], phi[:, r])
        slots = torch.zeros(n * L, 14, 3, device=dev, dtype=f64)
        slots[:, 0], slots[:, 1], slots[:, 2] = N.reshape(-1, 3), CA.reshape(-1, 3), C.reshape(-1, 3)
        ftypes = types.reshape(-1)
        fnat = natoms_t[ftypes]
        sc_tor = U(n * L, 14) * 360.0 - 180.0
        for k in range(3, 14):
            sel = torch.nonzero(fnat > k).squeeze(1)
            if sel.numel() == 0:
                continue
            t = ftypes[sel]
            p = pred_t[t, k]
            a, b, c = slots[sel, p[:, 0]], slots[sel, p[:, 1]], slots[sel, p[:, 2]]
            tor = sc_tor[sel, k]
            if k == 4:  # CB ~120 degrees from N around the C-CA axis (see synth.py)
                tor = _dihedral(a, b, c, slots[sel, 0]) - 120.0 + Nrm(0, 5.0, sel.numel())
                tor = torch.remainder(tor + 180.0, 360.0) - 180.0
            slots[sel, k] = _place(a, b, c, blen_t[t, k], bang_t[t, k], tor)
        atom_valid = torch.arange(14, device=dev)[None, :] < fnat[:, None]
        out["xyz"].append((torch.round(slots[atom_valid] * 1000.0) / 1000.0).to(torch.float32))
        out["res_type"].append(ftypes.to(torch.uint8))
        out["bfactor"].append((torch.round((30.0 + 68.0 * U(n * L)) * 100.0) / 100.0).to(torch.float32))
        apc = fnat.reshape(n, L).sum(1)
        out["atoms_per_chain"].append(apc)
        oxt = _place(N[:, L - 1], CA[:, L - 1], C[:, L - 1], torch.tensor(1.25, device=dev, dtype=f64), torch.tensor(118.0, device=dev, dtype=f64),
                     sc_tor.reshape(n, L, 14)[:, L - 1, 3] + 180.0)
        meta = torch.zeros(n, 20, dtype=torch.uint8, device=dev)
        na = ((apc + 1) & 0xFFFF).to(torch.int32)
        meta[:, 0], meta[:, 1] = (na & 0xFF).to(torch.uint8), (na >> 8).to(torch.uint8)
        meta[:, 2], meta[:, 4] = 1, 1          # idx_residue = idx_atom = 1
        meta[:, 6], meta[:, 7] = ord("A"), 1   # chain, has_oxt
        meta[:, 8:20] = (torch.round(oxt * 1000.0) / 1000.0).to(torch.float32).contiguous().view(torch.uint8).reshape(n, 12)
        out["meta"].append(meta)
        del N, CA, C, slots, sc_tor
    apc = torch.cat(out["atoms_per_chain"])
    atom_off = torch.zeros(n_chains + 1, dtype=torch.int64, device=dev)
    atom_off[1:] = torch.cumsum(apc, 0)
    titles = np.frombuffer(b"".join(b"syn_%07d" % (first_index + i) for i in range(n_chains)), np.uint8)
    return {
        "res_off": (torch.arange(n_chains + 1, device=dev, dtype=torch.int64) * L).to(torch.int32),
        "atom_off": atom_off,
        "title_off": (torch.arange(n_chains + 1, device=dev, dtype=torch.int64) * 11).to(torch.int32),
        "res_type": torch.cat(out["res_type"]), "bfactor": torch.cat(out["bfactor"]), "xyz": torch.cat(out["xyz"]),
        "titles": torch.from_numpy(titles.copy()).to(dev), "meta": torch.cat(out["meta"]),
    }


def generate_device_mixed(lengths, counts, seed: int, device, first_index: int = 0):
    """Chains of several lengths, every one DISTINCT: counts[i] chains of lengths[i] residues each, generated length by
    length (generate_device) and concatenated in that order.  Same dict as generate_device."""
    dev = torch.device(device)
    parts, first = [], int(first_index)
    for L, n in zip(lengths, counts):
        n, L = int(n), int(L)
        if n <= 0:
            continue
        parts.append(generate_device(n, L, seed, first, dev, chunk=max(64, 6_000_000 // L)))
        first += n

    def offsets(key, dt):
        out, base = [torch.zeros(1, dtype=torch.int64, device=dev)], 0
        for p in parts:
            o = p[key].to(torch.int64)
            out.append(o[1:] + base)
            base += int(o[-1].item())
        return torch.cat(out).to(dt)

    return {
        "res_off": offsets("res_off", torch.int32), "atom_off": offsets("atom_off", torch.int64), "title_off": offsets("title_off", torch.int32),
        **{k: torch.cat([p[k] for p in parts]) for k in ("res_type", "bfactor", "xyz", "titles", "meta")},
    }
