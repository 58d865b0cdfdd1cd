"""Python face of the C-ABI engine (include/fcz_engine.h) -- thin ctypes calls, no arithmetic here.

Two kinds of batches:
  * host batches  (`abi.HostChainBatch` / `abi.HostBlobBatch`, numpy): the engine copies in/out;
  * device batches (`DeviceChainBatch` / `DeviceBlobBatch`, torch CUDA tensors): zero-copy, the
    engine only enqueues kernels on the given stream.
PyTorch is used for device memory and streams only.  There is no CPU fallback: if the CUDA library
or a GPU is missing, construction fails.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi
from ._lib import load
from .abi import FczBlobBatch, FczChainBatch, FczOpts, FczSizes, FczTextBatch, HostBlobBatch, HostChainBatch, HostTextBatch


class FczError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"fcz error {code}: {msg}")
        self.code = code


def _torch():
    import torch

    return torch


class DeviceChainBatch:
    """Chains in canonical slot order on one GPU (torch tensors; int32/int64 hold the uint32/uint64 offsets)."""

    def __init__(self, n_chains: int, n_res: int, n_atoms: int, n_title: int, device):
        torch = _torch()
        kw = dict(device=device)
        self.n_chains = n_chains
        self.res_off = torch.zeros(n_chains + 1, dtype=torch.int32, **kw)
        self.atom_off = torch.zeros(n_chains + 1, dtype=torch.int64, **kw)
        self.title_off = torch.zeros(n_chains + 1, dtype=torch.int32, **kw)
        self.res_type = torch.zeros(max(n_res, 1), dtype=torch.uint8, **kw)
        self.bfactor = torch.zeros(max(n_res, 1), dtype=torch.float32, **kw)
        self.xyz = torch.zeros((max(n_atoms, 1), 3), dtype=torch.float32, **kw)
        self.titles = torch.zeros(max(n_title, 1), dtype=torch.uint8, **kw)
        self.meta = torch.zeros((max(n_chains, 1), abi.META_DTYPE.itemsize), dtype=torch.uint8, **kw)
        self.status = torch.zeros(max(n_chains, 1), dtype=torch.int32, **kw)
        self.n_res, self.n_atoms, self.n_title = n_res, n_atoms, n_title

    @staticmethod
    def from_host(b: HostChainBatch, device, pinned: bool = False) -> "DeviceChainBatch":
        torch = _torch()
        d = DeviceChainBatch(b.n_chains, b.n_res, b.n_atoms, len(b.titles), device)

        def put(dst, arr, dt):
            if arr.size == 0:
                return
            src = torch.from_numpy(np.ascontiguousarray(arr).view(dt).reshape(-1))
            dst.view(-1)[: src.numel()].copy_(src)

        put(d.res_off, b.res_off, np.int32)
        put(d.atom_off, b.atom_off, np.int64)
        put(d.title_off, b.title_off, np.int32)
        put(d.res_type, b.res_type, np.uint8)
        put(d.bfactor, b.bfactor, np.float32)
        put(d.xyz, b.xyz, np.float32)
        put(d.titles, b.titles, np.uint8)
        put(d.meta, b.meta.view(np.uint8), np.uint8)
        return d

    def to_host(self) -> HostChainBatch:
        n = self.n_chains
        res_off = self.res_off.cpu().numpy().view(np.uint32).copy()
        atom_off = self.atom_off.cpu().numpy().view(np.uint64).copy()
        title_off = self.title_off.cpu().numpy().view(np.uint32).copy()
        nr, na, nt = int(res_off[n]), int(atom_off[n]), int(title_off[n])
        return HostChainBatch(
            res_off=res_off,
            atom_off=atom_off,
            title_off=title_off,
            res_type=self.res_type[:nr].cpu().numpy().copy(),
            bfactor=self.bfactor[:nr].cpu().numpy().copy(),
            xyz=self.xyz[:na].cpu().numpy().copy(),
            titles=self.titles[:nt].cpu().numpy().copy(),
            meta=self.meta[:n].cpu().numpy().copy().view(abi.META_DTYPE).reshape(-1),
            status=self.status[:n].cpu().numpy().copy(),
        )

    def as_struct(self) -> FczChainBatch:
        s = FczChainBatch()
        s.n_chains = self.n_chains
        s.mem = abi.FCZ_MEM_DEVICE
        s.res_off = self.res_off.data_ptr()
        s.atom_off = self.atom_off.data_ptr()
        s.title_off = self.title_off.data_ptr()
        s.res_type = self.res_type.data_ptr()
        s.bfactor = self.bfactor.data_ptr()
        s.xyz = self.xyz.data_ptr()
        s.titles = self.titles.data_ptr()
        s.meta = self.meta.data_ptr()
        s.status = self.status.data_ptr()
        s.res_cap = self.res_type.numel()
        s.atom_cap = self.xyz.shape[0]
        s.title_cap = self.titles.numel()
        return s

    def nbytes_payload(self) -> int:
        """Bytes of the arrays the encode kernel reads / the decode kernel writes."""
        return self.n_res * 5 + self.n_atoms * 12


class DeviceBlobBatch:
    def __init__(self, n_chains: int, cap: int, device):
        torch = _torch()
        self.n_chains = n_chains
        self.blob_off = torch.zeros(n_chains + 1, dtype=torch.int64, device=device)
        self.bytes = torch.zeros(max(cap, 16), dtype=torch.uint8, device=device)
        self.status = torch.zeros(max(n_chains, 1), dtype=torch.int32, device=device)

    @staticmethod
    def from_host(b: HostBlobBatch, device) -> "DeviceBlobBatch":
        torch = _torch()
        d = DeviceBlobBatch(b.n_chains, int(b.blob_off[-1]), device)
        d.blob_off.copy_(torch.from_numpy(b.blob_off.view(np.int64)))
        nb = int(b.blob_off[-1])
        if nb:
            d.bytes[:nb].copy_(torch.from_numpy(np.ascontiguousarray(b.bytes[:nb])))
        return d

    def to_host(self) -> HostBlobBatch:
        off = self.blob_off.cpu().numpy().view(np.uint64).copy()
        nb = int(off[-1])
        return HostBlobBatch(off, self.bytes[:nb].cpu().numpy().copy(), self.status[: self.n_chains].cpu().numpy().copy())

    def as_struct(self) -> FczBlobBatch:
        s = FczBlobBatch()
        s.n_chains = self.n_chains
        s.mem = abi.FCZ_MEM_DEVICE
        s.blob_off = self.blob_off.data_ptr()
        s.bytes = self.bytes.data_ptr()
        s.status = self.status.data_ptr()
        s.bytes_cap = self.bytes.numel()
        return s


class DeviceTextBatch:
    """Texts on one GPU: text_off (int64 holding uint64) and a byte buffer."""

    def __init__(self, n_chains: int, cap: int, device):
        torch = _torch()
        self.n_chains = n_chains
        self.text_off = torch.zeros(n_chains + 1, dtype=torch.int64, device=device)
        self.bytes = torch.zeros(max(cap, 16), dtype=torch.uint8, device=device)

    def as_struct(self) -> FczTextBatch:
        s = FczTextBatch()
        s.n_chains = self.n_chains
        s.mem = abi.FCZ_MEM_DEVICE
        s.text_off = self.text_off.data_ptr()
        s.bytes = self.bytes.data_ptr()
        s.bytes_cap = self.bytes.numel()
        s.status = None
        return s

    def to_host(self) -> HostTextBatch:
        off = self.text_off.cpu().numpy().view(np.uint64).copy()
        return HostTextBatch(off, self.bytes[: int(off[-1])].cpu().numpy().copy())


class Engine:
    """One engine per GPU (fcz_engine_create).  `stream` is a torch.cuda.Stream or None (engine-owned)."""

    def __init__(self, device: int = 0, anchor_threshold: int = abi.DEFAULT_ANCHOR_THRESHOLD, use_alt_atom_order: bool = False, stream=None):
        self.lib = load()
        self.device = int(device)
        self._opts = FczOpts(int(anchor_threshold), int(bool(use_alt_atom_order)), None, 0)
        if stream is not None:
            self._opts.stream = stream.cuda_stream
        self._stream = stream
        self.h = self.lib.fcz_engine_create(self.device, C.byref(self._opts))
        if not self.h:
            raise FczError(abi.FCZ_E_CUDA, "fcz_engine_create failed (no CUDA device or out of memory); foldcomp_b200 has no CPU fallback")

    def close(self):
        if self.h:
            self.lib.fcz_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc != abi.FCZ_OK:
            raise FczError(rc, (self.lib.fcz_last_error(self.h) or b"").decode() or self.lib.fcz_strerror(rc).decode())

    def set_opts(self, anchor_threshold: int | None = None, use_alt_atom_order: bool | None = None, terminate_blobs: bool | None = None):
        if terminate_blobs is not None:
            self._opts.terminate_blobs = int(bool(terminate_blobs))
        if anchor_threshold is not None:
            self._opts.anchor_threshold = int(anchor_threshold)
        if use_alt_atom_order is not None:
            self._opts.use_alt_atom_order = int(bool(use_alt_atom_order))
        self._check(self.lib.fcz_engine_set_opts(self.h, C.byref(self._opts)))

    @property
    def anchor_threshold(self) -> int:
        return self._opts.anchor_threshold

    def sync(self):
        self._check(self.lib.fcz_engine_sync(self.h))

    def launch_count(self) -> int:
        return int(self.lib.fcz_engine_launch_count(self.h))

    def set_profiling(self, on: bool):
        self._check(self.lib.fcz_engine_set_profiling(self.h, int(bool(on))))

    def get_profile(self) -> abi.FczProfile:
        p = abi.FczProfile()
        self._check(self.lib.fcz_engine_get_profile(self.h, C.byref(p)))
        return p

    # ------------------------------------------------------------------ host batches (e2e path)
    def encode_host(self, batch: HostChainBatch, out: HostBlobBatch | None = None) -> HostBlobBatch:
        if out is None:
            cap = abi.encode_bound(batch.n_chains, batch.n_res, batch.n_atoms, len(batch.titles), self._opts.anchor_threshold)
            out = HostBlobBatch.empty(batch.n_chains, cap)
        sin, sout = batch.as_struct(), out.as_struct()
        self._check(self.lib.fcz_encode_batch(self.h, C.byref(sin), C.byref(sout)))
        return out

    def decode_host(self, blobs: HostBlobBatch, with_titles: bool = True, out: HostChainBatch | None = None) -> HostChainBatch:
        n = blobs.n_chains
        plan = HostChainBatch.empty(n) if out is None else out
        sin, sp = blobs.as_struct(), plan.as_struct()
        sizes = FczSizes()
        self._check(self.lib.fcz_decode_plan(self.h, C.byref(sin), C.byref(sp), C.byref(sizes)))
        if out is None:
            out = HostChainBatch(
                res_off=plan.res_off,
                atom_off=plan.atom_off,
                title_off=plan.title_off,
                res_type=np.zeros(sizes.n_res, np.uint8),
                bfactor=np.zeros(sizes.n_res, np.float32),
                xyz=np.zeros((sizes.n_atoms, 3), np.float32),
                titles=np.zeros(sizes.n_title_bytes, np.uint8),
                meta=np.zeros(n, abi.META_DTYPE),
                status=plan.status,
            )
        so = out.as_struct()
        if not with_titles:
            so.titles = None
        self._check(self.lib.fcz_decode_batch(self.h, C.byref(sin), C.byref(so)))
        return out

    # ------------------------------------------------------------------ device batches (kernel path)
    def encode_device(self, batch: DeviceChainBatch, out: DeviceBlobBatch) -> None:
        sin, sout = batch.as_struct(), out.as_struct()
        self._check(self.lib.fcz_encode_batch(self.h, C.byref(sin), C.byref(sout)))

    def decode_plan_device(self, blobs: DeviceBlobBatch, out: DeviceChainBatch) -> FczSizes:
        sin, so = blobs.as_struct(), out.as_struct()
        sizes = FczSizes()
        self._check(self.lib.fcz_decode_plan(self.h, C.byref(sin), C.byref(so), C.byref(sizes)))
        return sizes

    def decode_device(self, blobs: DeviceBlobBatch, out: DeviceChainBatch, with_titles: bool = True) -> None:
        sin, so = blobs.as_struct(), out.as_struct()
        if not with_titles:
            so.titles = None
        self._check(self.lib.fcz_decode_batch(self.h, C.byref(sin), C.byref(so)))

    # ------------------------------------------------------------------ text
    def pdb_text_host(self, chains: HostChainBatch) -> HostTextBatch:
        """PDB text of every chain of a decoded host batch (writeAtomCoordinatesToPDB, src/atom_coordinate.cpp:220-291)."""
        n = chains.n_chains
        out = HostTextBatch(np.zeros(n + 1, np.uint64), np.zeros(0, np.uint8))
        sin, so = chains.as_struct(), out.as_struct()
        total = C.c_uint64()
        self._check(self.lib.fcz_pdb_text_plan(self.h, C.byref(sin), C.byref(so), C.byref(total)))
        out.bytes = np.zeros(total.value, np.uint8)
        so = out.as_struct()
        self._check(self.lib.fcz_pdb_text_batch(self.h, C.byref(sin), C.byref(so)))
        return out

    def pdb_text_plan_device(self, chains: DeviceChainBatch, out: "DeviceTextBatch") -> int:
        sin, so = chains.as_struct(), out.as_struct()
        total = C.c_uint64()
        self._check(self.lib.fcz_pdb_text_plan(self.h, C.byref(sin), C.byref(so), C.byref(total)))
        return int(total.value)

    def pdb_text_device(self, chains: DeviceChainBatch, out: "DeviceTextBatch") -> None:
        sin, so = chains.as_struct(), out.as_struct()
        self._check(self.lib.fcz_pdb_text_batch(self.h, C.byref(sin), C.byref(so)))

    def decode_to_pdb_host(self, blobs: HostBlobBatch, out: HostTextBatch | None = None) -> HostTextBatch:
        """FCZ blobs -> PDB text in one call (what `foldcomp decompress` does per entry, src/main.cpp:612-689): the decoded
        coordinates stay on the GPU.  out.status holds the per-chain decode status (failed chains give empty text)."""
        n = blobs.n_chains
        if out is None:
            out = HostTextBatch(np.zeros(n + 1, np.uint64), np.zeros(0, np.uint8))
        sin, so = blobs.as_struct(), out.as_struct()
        total = C.c_uint64()
        self._check(self.lib.fcz_decode_to_pdb_plan(self.h, C.byref(sin), C.byref(so), C.byref(total)))
        if len(out.bytes) < total.value:
            out.bytes = np.zeros(total.value, np.uint8)
        so = out.as_struct()
        self._check(self.lib.fcz_decode_to_pdb_batch(self.h, C.byref(sin), C.byref(so)))
        return out

    def extract_host(self, blobs: HostBlobBatch, type_: int, digits: int = 2) -> HostTextBatch:
        """Foldcomp::extract for every blob: type_ 0 = pLDDT (digits 1..4), 1 = sequence (src/foldcomp.cpp:1260-1336)."""
        n = blobs.n_chains
        nb = int(blobs.blob_off[-1]) if n else 0
        out = HostTextBatch(np.zeros(n + 1, np.uint64), np.zeros(6 * nb // 8 + 64, np.uint8))  # <= 6 chars per residue, >= 8 blob bytes per residue
        sin, so = blobs.as_struct(), out.as_struct()
        total = C.c_uint64()
        self._check(self.lib.fcz_extract_batch(self.h, C.byref(sin), int(type_), int(digits), C.byref(so), C.byref(total)))
        return out

    # ------------------------------------------------------------------ PDB text in
    def parse_pdb_device(self, texts: "DeviceTextBatch") -> DeviceChainBatch:
        """Single-chain PDB texts on the GPU -> canonical chains on the GPU (fcz_parse_pdb_plan + fcz_parse_pdb_batch: the
        ATOM parser of foldcomp/foldcomp.cxx:253-293).  status[c] is 0 or FCZ_E_PARSE_*; titles are left empty."""
        n = texts.n_chains
        plan = DeviceChainBatch(n, 0, 0, 0, texts.bytes.device)
        sin, sp = texts.as_struct(), plan.as_struct()
        sizes = FczSizes()
        self._check(self.lib.fcz_parse_pdb_plan(self.h, C.byref(sin), C.byref(sp), C.byref(sizes)))
        out = DeviceChainBatch(n, int(sizes.n_res), int(sizes.n_atoms), 0, texts.bytes.device)
        out.res_off, out.atom_off, out.status = plan.res_off, plan.atom_off, plan.status
        so = out.as_struct()
        self._check(self.lib.fcz_parse_pdb_batch(self.h, C.byref(sin), C.byref(so)))
        return out

    def encode_pdb_text_host(self, texts: HostTextBatch, titles) -> HostBlobBatch:
        """PDB texts (host) -> FCZ blobs (host) in one call, parser and encoder on the GPU (what `foldcomp compress` does per
        entry, src/main.cpp:438-536).  titles: one bytes object per entry.  out.status: parser or encoder status."""
        n = texts.n_chains
        title_off = np.zeros(n + 1, np.uint32)
        title_off[1:] = np.cumsum([len(t) for t in titles], dtype=np.uint64).astype(np.uint32) if n else 0
        tbytes = np.frombuffer(b"".join(titles) + b"\0", np.uint8).copy()
        n_text = int(texts.text_off[-1]) if n else 0
        # FCZ is about 1/40 of its PDB text (15.9 of 634 bytes per residue at -b 25); retry once with the exact size
        out = HostBlobBatch.empty(n, n_text // 16 + 512 * n + int(title_off[-1]) + 1024)
        total = C.c_uint64()
        sin, so = texts.as_struct(), out.as_struct()
        rc = self.lib.fcz_encode_pdb_text_batch(self.h, C.byref(sin), title_off.ctypes.data, tbytes.ctypes.data, C.byref(so), C.byref(total))
        if rc == abi.FCZ_E_CAPACITY:
            out = HostBlobBatch.empty(n, int(total.value) + 64)
            so = out.as_struct()
            rc = self.lib.fcz_encode_pdb_text_batch(self.h, C.byref(sin), title_off.ctypes.data, tbytes.ctypes.data, C.byref(so), C.byref(total))
        self._check(rc)
        return out

    def backbone_angles_host(self, chains: HostChainBatch) -> np.ndarray:
        """[R, 6] per residue: psi, omega, next phi, then the bond angles at N, CA, C -- the encoder's angles before
        quantisation (Foldcomp::preprocess, src/foldcomp.cpp:484-496), zero where the chain end leaves none."""
        out = np.zeros((chains.n_res, 6), np.float32)
        sin = chains.as_struct()
        self._check(self.lib.fcz_backbone_angles_batch(self.h, C.byref(sin), out.ctypes.data))
        return out

    def check_host(self, blobs: HostBlobBatch):
        """(read_status [n], validity [n]): Foldcomp::read + checkValidity for every blob (src/foldcomp.cpp:904-1036,
        1492-1532): read_status 0 / FCZ_E_MAGIC / FCZ_E_TRUNCATED, validity = ValidityError class 0..6 (abi.VALIDITY)."""
        n = blobs.n_chains
        rs, va = np.zeros(n, np.int32), np.zeros(n, np.int32)
        sin = blobs.as_struct()
        self._check(self.lib.fcz_check_batch(self.h, C.byref(sin), rs.ctypes.data, va.ctypes.data))
        return rs, va

    def unpack_angles_host(self, blobs: HostBlobBatch):
        """(res_off [n+1], angles [R, 6]): continuised phi, psi, omega, N-CA-C, CA-C-N, C-N-CA of every residue record."""
        n = blobs.n_chains
        res_off = np.zeros(n + 1, np.uint64)
        sin = blobs.as_struct()
        total = C.c_uint64()
        self._check(self.lib.fcz_unpack_angles_batch(self.h, C.byref(sin), res_off.ctypes.data, None, 0, C.byref(total)))
        ang = np.zeros((total.value, 6), np.float32)
        self._check(self.lib.fcz_unpack_angles_batch(self.h, C.byref(sin), res_off.ctypes.data, ang.ctypes.data, total.value, C.byref(total)))
        return res_off, ang
