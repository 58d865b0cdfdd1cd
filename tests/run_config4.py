#!/usr/bin/env python3
"""BASELINE.json configs[3]: 350-residue chains round-tripped on N GPUs of one box, sharded by chain with no
collective on the data path, ONE merged foldcomp database written by all ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/run_config4.py --chains-per-gpu 250000 --out /tmp/merged_db        # N = 8 -> 2 M chains, 700 M residues

Every rank owns chains [rank * chains_per_gpu, (rank + 1) * chains_per_gpu): 10 000 synthetic chains (the bench.py
generator, rank-specific seed offset) tiled on the device.  Encode runs with opts.terminate_blobs, so the rank's
output IS its slab of the data file; the only exchange is the all_gather of one int64 per rank (the exclusive scan
of slab sizes, foldcomp_b200/shard.py).  Each rank pwrite()s its slab at its offset and sends its index rows to
rank 0, which writes .index / .lookup / .dbtype in key order.  Checks: blobs of every replica identical to the first
one's, the first 10 000 spot-checked against the oracle when it is present, decode round trip within the loss of the
format, the merged file re-read through its index.  Prints one JSON line (rank 0)."""
import argparse
import json
import os
import struct
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from foldcomp_b200 import abi, shard, synth  # noqa: E402
from foldcomp_b200.engine import DeviceBlobBatch, DeviceChainBatch, Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--chains-per-gpu", type=int, default=250000)
ap.add_argument("--base", type=int, default=10000, help="distinct synthetic chains per rank (tiled up to chains-per-gpu)")
ap.add_argument("--out", default="/tmp/fcz_merged_db")
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n, nb = args.chains_per_gpu, min(args.base, args.chains_per_gpu)
reps = (n + nb - 1) // nb
base = synth.generate(nb, 350, seed=synth.SEED, first_index=rank * nb)
L = 350
stream = torch.cuda.Stream(device=dev)
eng = Engine(local, anchor_threshold=25, stream=stream)
eng.set_opts(terminate_blobs=True)


def tile_offsets(off, unit_total, dtype):
    o = torch.from_numpy(off[:-1].astype(np.int64)).to(dev)
    full = (o[None, :] + unit_total * torch.arange(reps, device=dev, dtype=torch.int64)[:, None]).reshape(-1)[:n]
    last = n - (reps - 1) * nb
    end = unit_total * (reps - 1) + int(off[last])
    return torch.cat([full, torch.tensor([end], device=dev, dtype=torch.int64)]).to(dtype), end


with torch.cuda.stream(stream):
    d = DeviceChainBatch(n, 1, 1, 1, dev)
    d.res_off, n_res = tile_offsets(base.res_off, base.n_res, torch.int32)
    d.atom_off, n_atoms = tile_offsets(base.atom_off, base.n_atoms, torch.int64)
    d.title_off, n_title = tile_offsets(base.title_off, len(base.titles), torch.int32)
    last = n - (reps - 1) * nb
    d.res_type = torch.from_numpy(base.res_type).to(dev).repeat(reps)[:n_res].contiguous()
    d.bfactor = torch.from_numpy(base.bfactor).to(dev).repeat(reps)[:n_res].contiguous()
    d.xyz = torch.from_numpy(base.xyz).to(dev).repeat(reps, 1)[:n_atoms].contiguous()
    d.titles = torch.from_numpy(base.titles).to(dev).repeat(reps)[:n_title].contiguous()
    d.meta = torch.from_numpy(base.meta.view(np.uint8).reshape(nb, -1)).to(dev).repeat(reps, 1)[:n].contiguous()
    d.status = torch.zeros(n, dtype=torch.int32, device=dev)
    d.n_res, d.n_atoms, d.n_title = n_res, n_atoms, n_title
    cap = abi.encode_bound(n, n_res, n_atoms, n_title, 25)
    dblob = DeviceBlobBatch(n, cap, dev)
    dout = DeviceChainBatch(n, n_res, n_atoms, n_title, dev)
stream.synchronize()


def step():
    eng.encode_device(d, dblob)
    eng.decode_plan_device(dblob, dout)
    eng.decode_device(dblob, dout)


def barrier():
    stream.synchronize()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()


step()
barrier()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(stream)
for _ in range(args.steps):
    step()
ev1.record(stream)
barrier()
ms = torch.tensor([ev0.elapsed_time(ev1) / args.steps], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)

# ---- checks on the device result
assert int(dblob.status.count_nonzero().item()) == 0 and int(dout.status.count_nonzero().item()) == 0
boff = dblob.blob_off
slab_bytes = int(boff[-1].item())
unit = int(boff[nb].item()) if n > nb else slab_bytes
for k in range(1, reps):
    nk = nb if k < reps - 1 else last
    lo, hi = int(boff[k * nb].item()), int(boff[k * nb + nk].item())
    assert hi - lo == int(boff[nk].item()) and torch.equal(dblob.bytes[lo:hi], dblob.bytes[: hi - lo]), k
first = abi.HostBlobBatch(boff[: nb + 1].cpu().numpy().view(np.uint64).copy(), dblob.bytes[:unit].cpu().numpy().copy())
oracle_checked = 0
try:
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H

    for c in range(0, nb, 1000):
        assert first.blob(c) == H.oracle_encode(base, c, 25) + b"\0", c
        oracle_checked += 1
except ImportError:
    pass
got_xyz = dout.xyz[: base.n_atoms].cpu().numpy()
rt_rmsd = float(np.sqrt(((got_xyz - base.xyz) ** 2).sum(1).mean()))
assert rt_rmsd < 0.2, rt_rmsd

# ---- merged database: exclusive scan of slab sizes, every rank writes its slab at its offset
t0 = time.perf_counter()
if world > 1:
    base_off, total, totals = shard.merged_offsets(slab_bytes)
else:
    base_off, total, totals = 0, slab_bytes, [slab_bytes]
if rank == 0:
    with open(args.out, "wb") as f:
        f.truncate(total)
if world > 1:
    dist.barrier()
host = dblob.bytes[:slab_bytes].cpu().numpy()
fd = os.open(args.out, os.O_WRONLY)
pos = 0
while pos < slab_bytes:
    pos += os.pwrite(fd, memoryview(host[pos : pos + (256 << 20)]), base_off + pos)
os.close(fd)
offs = boff.cpu().numpy().view(np.uint64).astype(np.int64)
keys = rank * n + np.arange(n, dtype=np.int64)
rows = np.stack([keys, base_off + offs[:-1], np.diff(offs)], 1)
if world > 1:
    gathered = [None] * world
    dist.all_gather_object(gathered, rows)
else:
    gathered = [rows]
if rank == 0:
    allrows = np.concatenate(gathered)
    assert np.all(np.diff(allrows[:, 0]) > 0)  # key order
    with open(args.out + ".index", "w") as ix, open(args.out + ".lookup", "w") as lk:
        ix.write("".join(f"{k}\t{o}\t{ln}\n" for k, o, ln in allrows))
        lk.write("".join(f"{k}\tsyn_{k:07d}.fcz\t0\n" for k in allrows[:, 0]))
    with open(args.out + ".dbtype", "wb") as t:
        t.write(struct.pack("<i", 12))
if world > 1:
    dist.barrier()
write_s = time.perf_counter() - t0

# ---- re-read through the index: a sample of entries from every rank's slab equals what that key's generator gives
if rank == 0:
    data = np.memmap(args.out, np.uint8, "r")
    assert len(data) == total
    for r in range(world):
        k = r * n + 7
        key, off, ln = allrows[k]
        assert key == k and data[off + ln - 1] == 0 and bytes(data[off : off + 4]) == b"FCMP"
    mine = allrows[:3]
    for key, off, ln in mine:
        assert bytes(data[off : off + ln]) == first.blob(int(key))
    line = {
        "config": f"BASELINE.json configs[3]: {world * n} synthetic 350-residue chains, round trip sharded over {world} GPU(s), merged db",
        "n_gpus": world, "chains": world * n, "residues": world * n * L, "ms_per_round_trip": float(ms.item()),
        "residues_per_s": world * n * L / (float(ms.item()) * 1e-3), "fcz_bytes_total": int(total), "db_write_s": write_s,
        "roundtrip_rmsd_vs_input": rt_rmsd, "replicas_identical": True, "oracle_spot_checks": oracle_checked,
        "exchange": "one all_gather of an int64 per rank (slab sizes) + index rows to rank 0; no data-path collective",
    }
    print(json.dumps(line), flush=True)
if world > 1:
    dist.destroy_process_group()
eng.close()
