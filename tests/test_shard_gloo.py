"""world_size-2 gloo test of the multi-GPU sharding logic on CPU: each rank encodes its shard (with
the oracle standing in for the GPU engine), the ranks exchange only their byte totals, and the merged
database equals the single-process one byte for byte."""
import os
import socket

import numpy as np
import pytest

import helpers as H
from foldcomp_b200 import shard, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lens = np.array([40, 120, 33, 350, 90, 61, 200, 17, 75])
    batch = synth.generate(len(lens), lens, seed=4242)
    mine = shard.balanced_shards(lens, world)[rank]
    sub = batch.select(mine)
    blobs = H.oracle_encode_batch(sub, 25, 1)
    # NUL-terminated entries: slab size = blob bytes + one terminator per entry
    slab = b"".join(blobs.blob(i) + b"\0" for i in range(sub.n_chains))
    base, total, totals = shard.merged_offsets(len(slab))
    rows = shard.merged_index(mine.tolist(), blobs.blob_off, base)
    q.put((rank, base, total, slab, rows))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_merged_db_matches_single_process():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, base0, total0, slab0, rows0), (_, base1, total1, slab1, rows1) = res
    assert base0 == 0 and base1 == len(slab0) and total0 == total1 == len(slab0) + len(slab1)
    data = slab0 + slab1
    # every chain is found at its index row and equals the single-process encoding
    lens = np.array([40, 120, 33, 350, 90, 61, 200, 17, 75])
    batch = synth.generate(len(lens), lens, seed=4242)
    seen = set()
    for key, off, ln in rows0 + rows1:
        assert data[off + ln - 1 : off + ln] == b"\0"
        assert data[off : off + ln - 1] == H.oracle_encode(batch, key, 25)
        seen.add(key)
    assert seen == set(range(len(lens)))


def test_shard_helpers():
    assert [shard.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    lens = np.array([2000, 50, 60, 70, 1000, 900, 100])
    bins = shard.balanced_shards(lens, 2)
    assert sorted(np.concatenate(bins).tolist()) == list(range(7))
    loads = [int(lens[b].sum()) for b in bins]
    assert abs(loads[0] - loads[1]) <= 200


def test_bench_clock_sampler_without_nvml():
    """bench.py's clock sampler on a box without a GPU: no exception, the JSON object says why there are no samples."""
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    s = bench.ClockSampler("GPU-00000000-0000-0000-0000-000000000000", 0)
    s.mark_begin()
    s.mark_end()
    out = s.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons"} and (out["sm_mhz"] is None or out["sm_mhz"] > 0)
