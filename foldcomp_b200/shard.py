"""Multi-GPU sharding helpers (host logic).  Chains are independent units, so a batch shards by
chain with NO collective on the data path; only a merged database needs an exchange:
the exclusive scan of per-rank byte totals, so that every rank knows where its slab of blobs starts.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) of `n_items` for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def balanced_shards(lengths, world: int):
    """Longest-first greedy bin packing on residue counts: list of index arrays, one per rank, each
    in ascending chain order (mixed-length batches, BASELINE.json config 5)."""
    lengths = np.asarray(lengths, np.int64)
    order = np.argsort(-lengths, kind="stable")
    load = np.zeros(world, np.int64)
    bins = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))
        bins[r].append(int(i))
        load[r] += lengths[i]
    return [np.array(sorted(b), np.int64) for b in bins]


def merged_offsets(local_total: int, group=None):
    """(base offset of this rank's slab, grand total, per-rank totals) for a merged output: exclusive
    scan of the per-rank totals via one all_gather of a single int64 (gloo on CPU, NCCL on GPUs)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.tensor([int(local_total)], dtype=torch.int64, device=dev)
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine, group=group)
    totals = [int(t.item()) for t in allv]
    return sum(totals[:rank]), sum(totals), totals


def merged_index(keys, blob_off, base: int):
    """mmseqs-style index rows (key, offset, length incl. the NUL terminator) for this rank's slab of a
    merged foldcomp database (src/database_writer.cpp:75-96); entries are NUL-terminated,
    so entry i sits at base + blob_off[i] + i."""
    blob_off = np.asarray(blob_off, np.int64)
    lens = np.diff(blob_off) + 1
    offs = base + blob_off[:-1] + np.arange(len(lens))
    return [(int(k), int(o), int(l)) for k, o, l in zip(keys, offs, lens)]
