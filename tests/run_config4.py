#!/usr/bin/env python3
"""BASELINE.json configs[3]: 350-residue chains round-tripped on N GPUs of one box, sharded by chain with no
collective on the data path, ONE merged foldcomp database written by all ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/run_config4.py --chains-per-gpu 250000 --out /tmp/merged_db        # N = 8 -> 2 M chains, 700 M residues

Every rank owns chains [rank * chains_per_gpu, (rank + 1) * chains_per_gpu), every one of them DISTINCT: generated on the
device (foldcomp_b200/synth_device.py) from the chain's global index.  Encode runs with opts.terminate_blobs, so the rank's
output IS its slab of the data file; the only exchange is the all_gather of one int64 per rank (the exclusive scan
of slab sizes, foldcomp_b200/shard.py).  Each rank pwrite()s its slab at its offset and sends its index rows to
rank 0, which writes .index / .lookup / .dbtype in key order.  Checks: every chain's status and round trip against its
own input (on the device), the oracle on every 1000th chain of every rank (FCZ bytes identical, decode within tolerance),
the merged file re-read through its index.  Prints one JSON line (rank 0)."""
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
from foldcomp_b200 import abi, shard, synth, synth_device  # noqa: E402
from foldcomp_b200.engine import DeviceBlobBatch, DeviceChainBatch, Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--chains-per-gpu", type=int, default=250000)
ap.add_argument("--out", default="/tmp/fcz_merged_db")
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = args.chains_per_gpu
L = 350
stream = torch.cuda.Stream(device=dev)
eng = Engine(local, anchor_threshold=25, stream=stream)
eng.set_opts(terminate_blobs=True)

# every chain DISTINCT: the synthetic generator on the device (foldcomp_b200/synth_device.py), keyed by the chain's
# global index (rank * chains_per_gpu + i)
t_gen = time.perf_counter()
with torch.cuda.stream(stream):
    g = synth_device.generate_device(n, L, synth.SEED, rank * n, dev)
    d = DeviceChainBatch(n, 1, 1, 1, dev)
    d.res_off, d.atom_off, d.title_off = g["res_off"], g["atom_off"], g["title_off"]
    d.res_type, d.bfactor, d.xyz, d.titles, d.meta = g["res_type"], g["bfactor"], g["xyz"], g["titles"], g["meta"]
    d.status = torch.zeros(n, dtype=torch.int32, device=dev)
    n_res, n_atoms, n_title = int(d.res_off[-1].item()), int(d.atom_off[-1].item()), int(d.title_off[-1].item())
    d.n_res, d.n_atoms, d.n_title = n_res, n_atoms, n_title
    cap = abi.encode_bound(n, n_res, n_atoms, n_title, 25)
    dblob = DeviceBlobBatch(n, cap, dev)
    dout = DeviceChainBatch(n, n_res, n_atoms, n_title, dev)
stream.synchronize()
t_gen = time.perf_counter() - t_gen


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

# ---- checks on the device result: every chain's status, every chain's round trip against its own input (on the device),
# and the oracle on every 1000th chain: FCZ bytes identical, decoded coordinates within the tolerance
assert int(dblob.status.count_nonzero().item()) == 0 and int(dout.status.count_nonzero().item()) == 0
boff = dblob.blob_off
slab_bytes = int(boff[-1].item())
assert torch.equal(dout.res_type[:n_res], d.res_type) and torch.equal(dout.atom_off.to(torch.int64), d.atom_off.to(torch.int64))
d2 = ((dout.xyz[:n_atoms].double() - d.xyz.double()) ** 2).sum(1)
rt_rmsd = float(torch.sqrt(d2.mean()).item())
chain_of_atom = torch.repeat_interleave(torch.arange(n, device=dev), (d.atom_off[1:] - d.atom_off[:-1]).to(torch.int64))
per_chain = torch.sqrt(torch.zeros(n, device=dev, dtype=torch.float64).index_add_(0, chain_of_atom, d2) / (d.atom_off[1:] - d.atom_off[:-1]).double())
worst_chain_rmsd = float(per_chain.max().item())
assert rt_rmsd < 0.1 and worst_chain_rmsd < 0.5, (rt_rmsd, worst_chain_rmsd)
oracle_checked, oracle_bb, oracle_max = 0, 0.0, 0.0
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # the checker (oracle/), test infrastructure

sample = list(range(0, n, 1000))
h_res_off, h_atom_off = d.res_off.cpu().numpy(), d.atom_off.cpu().numpy()
h_boff = boff.cpu().numpy().view(np.uint64).astype(np.int64)
for c in sample:
    r0, r1, a0, a1 = int(h_res_off[c]), int(h_res_off[c + 1]), int(h_atom_off[c]), int(h_atom_off[c + 1])
    one = abi.HostChainBatch(
        res_off=np.array([0, r1 - r0], np.uint32), atom_off=np.array([0, a1 - a0], np.uint64), title_off=np.array([0, 11], np.uint32),
        res_type=d.res_type[r0:r1].cpu().numpy(), bfactor=d.bfactor[r0:r1].cpu().numpy(), xyz=d.xyz[a0:a1].cpu().numpy(),
        titles=d.titles[11 * c : 11 * c + 11].cpu().numpy(), meta=d.meta[c : c + 1].cpu().numpy().view(abi.META_DTYPE).reshape(-1))
    want = H.oracle_encode(one, 0, 25)
    got = bytes(dblob.bytes[int(h_boff[c]) : int(h_boff[c + 1])].cpu().numpy())
    assert got == want + b"\0", (rank, c)
    ref = H.oracle_decode(want)
    mine = dout.xyz[a0:a1].cpu().numpy()
    bbm = H.backbone_mask(ref.res_type)
    oracle_bb = max(oracle_bb, H.rmsd(mine[bbm], ref.xyz[bbm]))
    oracle_max = max(oracle_max, H.max_dev(mine, ref.xyz))
    oracle_checked += 1
assert oracle_bb <= 0.01 and oracle_max <= 0.05, (oracle_bb, oracle_max)
first3 = [bytes(dblob.bytes[int(h_boff[c]) : int(h_boff[c + 1])].cpu().numpy()) for c in range(3)]

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
    for (key, off, ln), blob in zip(allrows[:3], first3):
        assert bytes(data[off : off + ln]) == blob
    line = {
        "config": f"BASELINE.json configs[3]: {world * n} synthetic 350-residue chains, round trip sharded over {world} GPU(s), merged db",
        "n_gpus": world, "chains": world * n, "residues": world * n * L, "ms_per_round_trip": float(ms.item()),
        "residues_per_s": world * n * L / (float(ms.item()) * 1e-3), "fcz_bytes_total": int(total), "db_write_s": write_s,
        "distinct_chains": True, "generator_s": t_gen, "roundtrip_rmsd_vs_input_all_atoms": rt_rmsd, "worst_chain_rmsd_vs_input": worst_chain_rmsd,
        "oracle_checks_per_rank": oracle_checked, "oracle_fcz_bytes_identical": True, "decode_vs_oracle_bb_rmsd_max": oracle_bb,
        "decode_vs_oracle_max_dev": oracle_max,
        "exchange": "one all_gather of an int64 per rank (slab sizes) + index rows to rank 0; no data-path collective",
    }
    print(json.dumps(line), flush=True)
if world > 1:
    dist.destroy_process_group()
eng.close()
