#!/usr/bin/env python3
"""bench.py -- residues/s for the full FCZ compress -> decompress round trip (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is ONE round trip of the hot path over one batch: fcz_encode_batch -> fcz_decode_plan ->
fcz_decode_batch on BASELINE.json configs[1] (10 000 synthetic single chains of 350 residues,
-b 25), per GPU (weak scaling: every rank owns its own 10 000 chains, no data-path collective).

  value      whole-job residues/s with the batch resident in HBM (device-pointer C ABI), K steps
             between two CUDA events on the engine's stream, max over ranks;
  e2e        the same round trip through the host-pointer C ABI (pinned host buffers; H2D of the
             coordinates and D2H of blobs and decoded coordinates inside the timed region), encode and
             decode on two engines so that their PCIe directions overlap; wall clock;
  roofline   the dominant hot kernel (longest device time per step among k_encode, k_dec_front,
             k_dec_stitch_t, k_dec_back): its algorithmic bytes per launch (SURVEY.md 8d / BASELINE.md 5)
             over its mean launch duration (CUDA events the engine records around each launch on the
             launching stream, same timed region as `value`), against MEASURED_PEAKS.json hbm_gbs;
  cpu_baseline  the reference's own CPU path (oracle/_ref: unmodified sources compiled in place;
             Foldcomp::compress+writeStream+read+decompress per chain under OpenMP) on a bounded
             sample of the same workload, all host cores, rank 0 at N=1 only;
  --impl reference   that CPU path alone, same JSON contract ("impl": "reference").
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CHAINS, LENGTH, ANCHOR = 10000, 350, 25
E2E_PARTS = int(os.environ.get("FCZ_E2E_PARTS", "4"))  # sub-batches per step on the end-to-end path
E2E_PAIRS = int(os.environ.get("FCZ_E2E_PAIRS", "1"))  # (encode engine, decode engine) pairs working on alternate sub-batches
WORKLOAD_NOTE = ""
METRIC = "residues/sec compress+decompress round-trip"
UNIT = "residues/s"


def _len_desc():
    return int(LENGTH) if np.isscalar(LENGTH) else round(float(np.mean(LENGTH)), 1)


def workload_config(n_gpus):
    return {
        "workload": f"BASELINE.json configs[1]: {N_CHAINS} synthetic single chains x {_len_desc()} residues per GPU, -b {ANCHOR}, "
                    "compress -> decompress round trip" + WORKLOAD_NOTE,
        "chains_per_gpu": N_CHAINS, "residues_per_chain": _len_desc(), "anchor_threshold": ANCHOR,
        "parallelism": f"chains sharded over {n_gpus} GPU(s), no collective on the data path",
        "l2": "per step the kernels stream ~346 MB of coordinates in, 56 MB of FCZ and ~343 MB of coordinates out "
              "(> 126 MB L2), so inputs come from HBM every step; no explicit flush",
    }


# ------------------------------------------------------------------------------------ CPU reference


def _cpu_lib():
    """(lib, kind): the unmodified reference compiled in place if present, else the oracle port."""
    ref = os.path.join(ROOT, "oracle", "_ref", "libfoldcomp_ref.so")
    if os.path.exists(ref):
        lib = C.CDLL(ref)
        lib.ref_roundtrip_batch.restype = C.c_int
        lib.ref_roundtrip_batch.argtypes = [C.c_int] + [C.c_void_p] * 7 + [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        return lib, "reference"
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H

    return H, "port"


def cpu_roundtrip(batch, n_chains, threads):
    """Round trip the first n_chains chains of `batch` on the CPU; returns seconds."""
    lib, kind = _cpu_lib()
    sub = batch if n_chains >= batch.n_chains else batch.select(range(n_chains))
    t0 = time.perf_counter()
    if kind == "reference":
        has_oxt = np.ascontiguousarray(sub.meta["has_oxt"])
        oxt = np.ascontiguousarray(sub.meta["oxt"])
        tb, cs = C.c_uint64(), C.c_uint64()
        rc = lib.ref_roundtrip_batch(sub.n_chains, sub.res_off.ctypes.data, sub.atom_off.ctypes.data, sub.res_type.ctypes.data,
                                     sub.xyz.ctypes.data, sub.bfactor.ctypes.data, has_oxt.ctypes.data, oxt.ctypes.data,
                                     ANCHOR, threads, 0, C.byref(tb), C.byref(cs))
        assert rc == 0, rc
    else:
        blobs = lib.oracle_encode_batch(sub, ANCHOR, threads)
        lib.oracle_decode_batch(blobs, False, threads)
    return time.perf_counter() - t0, kind, sub.n_res


def cpu_sample_size(batch, threads, target_s):
    """Calibrate on 128 chains, then size the sample for ~target_s seconds."""
    cpu_roundtrip(batch, 128, threads)  # first call pays library load and thread start-up
    dt, kind, nres = cpu_roundtrip(batch, min(batch.n_chains, 512), threads)
    rate = nres / dt
    n = int(max(128, min(batch.n_chains, rate * target_s / float(np.mean(LENGTH)))))
    return n


def run_reference(args, rank, world, out):
    if rank != 0:
        return
    from foldcomp_b200 import synth

    threads = os.cpu_count() or 1
    batch = synth.generate(N_CHAINS, LENGTH, seed=synth.SEED)  # LENGTH: int or per-chain array
    n = cpu_sample_size(batch, threads, 3.0)
    for _ in range(args.warmup):
        cpu_roundtrip(batch, n, threads)
    t = 0.0
    nres = 0
    kind = "port"
    for _ in range(args.steps):
        dt, kind, r = cpu_roundtrip(batch, n, threads)
        t += dt
        nres += r
    value = nres / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{n} of the {N_CHAINS} chains per step, all host threads (OpenMP over chains), in-memory "
                                   "Foldcomp::compress+writeStream+read+decompress; the timed region includes the adapter's rebuild of "
                                   "vector<AtomCoordinate> from the canonical arrays and back (oracle/ref_shim.cpp), a few percent of a step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=out, flush=True)


# --------------------------------------------------------------------------------------- clocks


class ClockSampler:
    """SM clock and throttle reasons of one GPU while the timed passes run: NVML polled every 2 ms from a thread of this
    process (nvidia_ml_py; an `nvidia-smi -lms` child needs longer to start than the timed region lasts).  The sampler is
    started before the warm-up steps; stop() keeps the samples taken between mark_begin() and mark_end() -- the timed and
    the per-kernel-event passes -- and says how many there were."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_uuid, index):
        import threading

        self.rows, self.err, self._stop = [], None, threading.Event()
        self.t0 = self.t1 = None
        try:
            import pynvml

            pynvml.nvmlInit()
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID(gpu_uuid.encode() if isinstance(gpu_uuid, str) else gpu_uuid)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(int(index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.err = f"NVML unavailable: {e}"
            return

        def run():
            while not self._stop.is_set():
                try:
                    self.rows.append((time.perf_counter(), float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                                      int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))))
                except Exception:  # noqa: BLE001
                    pass
                time.sleep(0.002)

        self.th = threading.Thread(target=run, daemon=True)
        self.th.start()

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.err:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err]}
        self._stop.set()
        self.th.join(timeout=2)
        rows = [r for r in self.rows if self.t0 is not None and self.t1 is not None and self.t0 <= r[0] <= self.t1]
        window = "timed passes"
        if len(rows) < 3:  # a region shorter than a few polls: the warm-up steps before it ran the same kernels
            rows, window = [r for r in self.rows if self.t1 is None or r[0] <= self.t1], "warm-up + timed passes"
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        bits = 0
        for r in rows:
            bits |= r[2]
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(nm for nm, m in self.REASONS if bits & m), "samples": len(rows), "window": window}


# --------------------------------------------------------------------------------------- ours


def bind_rank_to_cores(local_rank, world):
    """Give every rank its own host cores (and, through first touch, its own pinned memory): the cores of the GPU's NUMA node
    when sysfs names one, split among the ranks that share it; otherwise an even slice of the process's affinity set.
    Returns a description for the JSON line.  The ranks of one box otherwise share every core, and their copy threads,
    staging memcpys and pinned pages land wherever the scheduler puts them (e2e efficiency 0.20 at N = 8)."""
    try:
        import torch

        cores = sorted(os.sched_getaffinity(0))
        node, node_cores = -1, None
        try:
            bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
            dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
            dev_id = torch.cuda.get_device_properties(local_rank).pci_device_id
            path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev_id:02x}.0/numa_node"
            node = int(open(path).read().strip())
            if node >= 0:
                txt = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
                node_cores = []
                for part in txt.split(","):
                    a, _, b = part.partition("-")
                    node_cores += list(range(int(a), int(b or a) + 1))
                node_cores = [c for c in node_cores if c in cores]
        except Exception:
            node_cores = None
        pool = node_cores if node_cores else cores
        per = max(1, len(pool) // max(world, 1))
        mine = pool[(local_rank * per) % len(pool):][:per] or pool
        os.sched_setaffinity(0, mine)
        return {"numa_node": node, "cores": f"{mine[0]}-{mine[-1]}", "n_cores": len(mine), "host_cores_total": len(cores)}
    except Exception as ex:  # binding is an optimisation, never a requirement
        return {"error": str(ex)[:80]}


def pinned_like(arr):
    import torch

    t = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1)).pin_memory()
    return t.numpy().view(arr.dtype).reshape(arr.shape), t


def run_ours(args, rank, world, local_rank, out):
    import torch

    from foldcomp_b200 import abi, synth
    from foldcomp_b200.abi import HostBlobBatch, HostChainBatch
    from foldcomp_b200.engine import DeviceBlobBatch, DeviceChainBatch, Engine

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    binding = bind_rank_to_cores(local_rank, world) if os.environ.get("FCZ_BIND", "1") != "0" else {"disabled": True}
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    eng = Engine(local_rank, anchor_threshold=ANCHOR, stream=stream)

    batch = synth.generate(N_CHAINS, LENGTH, seed=synth.SEED, first_index=rank * N_CHAINS)
    n_res, n_atoms, n_title = batch.n_res, batch.n_atoms, len(batch.titles)
    cap = abi.encode_bound(batch.n_chains, n_res, n_atoms, n_title, ANCHOR)

    with torch.cuda.stream(stream):
        dbatch = DeviceChainBatch.from_host(batch, dev)
        dblob = DeviceBlobBatch(batch.n_chains, cap, dev)
        dout = DeviceChainBatch(batch.n_chains, n_res, n_atoms, n_title, dev)
    stream.synchronize()

    def step():
        eng.encode_device(dbatch, dblob)
        eng.decode_plan_device(dblob, dout)
        eng.decode_device(dblob, dout)

    def barrier():
        stream.synchronize()
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()

    try:
        gpu_id = str(torch.cuda.get_device_properties(local_rank).uuid)
        if not gpu_id.startswith("GPU-"):
            gpu_id = "GPU-" + gpu_id
    except Exception:
        gpu_id = str(local_rank)
    sampler = ClockSampler(gpu_id, local_rank)  # polls from here on: the warm-up steps run the same kernels
    # warm-up: the W steps asked for (at least 3), then on until the GPU has been busy for 100 ms -- an idle B200 sits at
    # 120 MHz and a 3 ms warm-up can end before the SM clock has ramped up (seen with eight ranks starting at once)
    n_warm = 0
    t_w = time.perf_counter()
    while n_warm < max(args.warmup, 3) or (time.perf_counter() - t_w < 0.1 and n_warm < 400):
        step()
        n_warm += 1
    barrier()
    # ---- timed region: kernels with HBM-resident inputs
    sampler.mark_begin()
    l0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - l0
    # the same K steps once more with the engine's per-launch events on (they cost a few microseconds per launch, so
    # they stay out of `value`): kernel durations for the roofline
    eng.set_profiling(True)
    eng.get_profile()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record(stream)
    for _ in range(args.steps):
        step()
    ev3.record(stream)
    barrier()
    ms_prof = ev2.elapsed_time(ev3)
    prof = eng.get_profile()
    eng.set_profiling(False)
    sampler.mark_end()
    clocks = sampler.stop()  # sampled over the device-resident timed passes only (the GPU idles between transfers later on)
    ms_t = torch.tensor([ms], dtype=torch.float64, device=dev)
    ms_all = [ms]
    if dist is not None:
        gathered = [torch.zeros_like(ms_t) for _ in range(world)]
        dist.all_gather(gathered, ms_t)
        ms_all = [float(g.item()) for g in gathered]
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_max = float(ms_t.item())
    value = world * n_res * args.steps / (ms_max * 1e-3)

    # sanity on the device result (not a parity test): statuses clean, round trip close to the input
    got = dout.to_host()
    assert not got.status.any() and got.n_atoms == n_atoms
    dev_rt = float(np.sqrt(((got.xyz[: 20000] - batch.xyz[: 20000]) ** 2).sum(1).mean()))
    assert dev_rt < 0.2, dev_rt
    fcz_bytes = int(dblob.blob_off[-1].item())

    # ---- roofline of the dominant kernel: the single kernel with the longest mean launch among the four hot
    # kernels, each timed by CUDA events recorded on its own launch stream inside the timed region above
    enc_bytes = 12 * n_atoms + 5 * n_res + fcz_bytes
    dec_bytes = fcz_bytes + 12 * n_atoms + 4 * n_res
    # algorithmic bytes per kernel (SURVEY.md 8d / BASELINE.md 5): encode reads coordinates+types+B-factors and
    # writes FCZ; decode as a whole reads FCZ and writes coordinates+types+B-factors -- of that, the front kernel
    # owns the FCZ read and the 5 B/residue it emits, the back kernel the coordinates it writes (it re-reads the
    # side-chain bytes); the stitch kernel moves no algorithmic bytes at all (segment scratch only).
    kalg = {"k_encode": enc_bytes, "k_dec_front": fcz_bytes + 5 * n_res, "k_dec_stitch_t": 0, "k_dec_back": 12 * n_atoms + (n_atoms - 3 * n_res)}
    kernels = {}
    for name, kind in abi.PROF_KINDS.items():
        n_l = int(prof.kernel_launches[kind])
        ms_l = float(prof.kernel_ms[kind]) / max(n_l, 1)
        entry = {"ms_per_launch": ms_l, "launches_per_step": n_l / args.steps}
        if name in kalg:
            entry["algorithmic_bytes"] = kalg[name]
            entry["gbs"] = kalg[name] / max(ms_l, 1e-9) / 1e6
        elif name == "encode_span":
            entry["algorithmic_bytes"] = enc_bytes
            entry["gbs"] = enc_bytes / max(ms_l, 1e-9) / 1e6
        elif name == "decode_span":
            entry["algorithmic_bytes"] = dec_bytes
            entry["gbs"] = dec_bytes / max(ms_l, 1e-9) / 1e6
        kernels[name] = entry
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # launches per step > 1 (length tiers) -> total time per step of that kernel decides dominance
    per_step = {k: kernels[k]["ms_per_launch"] * kernels[k]["launches_per_step"] for k in kalg}
    kname = max(per_step, key=per_step.get)
    kms = kernels[kname]["ms_per_launch"]
    kbytes = kalg[kname] / max(kernels[kname]["launches_per_step"], 1.0)
    achieved = kbytes / (kms * 1e-3) / 1e9 if kms > 0 else 0.0
    traffic, traffic_src = None, None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of that kernel: STATIC, from the committed ncu --set full capture of
        # this workload (hardware counters cannot be read by the run itself); traffic_source names the capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        t = tj[kname]
        traffic = int(t["dram_bytes_read"]) + int(t["dram_bytes_write"])
        traffic_src = "static, not measured by this run: " + tj.get("source", "profiles/ncu_traffic.json")
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "ms_per_launch": kms, "algorithmic_bytes_per_launch": kbytes,
        "kernels": kernels, "round_trip_bytes_per_residue": (enc_bytes + dec_bytes) / n_res,
        "round_trip_frac": (enc_bytes + dec_bytes) * args.steps / (ms * 1e-3) / 1e9 / peak,
        "ms_per_step_with_kernel_events": ms_prof / args.steps,
    }

    if args.kernels_only:  # profiling runs (ncu --set full): the device-resident region above is all that is needed
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "ms_per_step": ms_max / args.steps,
                              "roofline": roofline, "gpu_launches": launches, "note": "--kernels-only: no e2e / cpu_baseline legs"}), file=out, flush=True)
        eng.close()
        return

    # ---- e2e: host-pointer C ABI, pinned host buffers, every copy inside the timed region.  The batch goes
    # through in E2E_PARTS sub-batches on TWO engines driven by two host threads: one calls fcz_encode_batch,
    # the other fcz_decode_plan + fcz_decode_batch on the blobs the first produced, so the encode's H2D traffic
    # and the decode's D2H traffic share the full-duplex PCIe link (a compress job and a decompress job side
    # by side, as the reference's CLI runs them under OpenMP).  Every step round-trips every chain; the encode of one
    # sub-batch (or of the next step's first one) runs while the previous one decodes, and a blob buffer is reused
    # only after its decode has finished.
    import threading

    keep = []

    def pin(a):
        v, t = pinned_like(a)
        keep.append(t)
        return v

    def pinned_chain_batch(b):
        return HostChainBatch(pin(b.res_off), pin(b.atom_off), pin(b.title_off), pin(b.res_type), pin(b.bfactor),
                              pin(b.xyz), pin(b.titles), pin(b.meta), pin(np.zeros(b.n_chains, np.int32)))

    def pinned_out_batch(b):
        nt = len(b.titles)
        return HostChainBatch(pin(np.zeros(b.n_chains + 1, np.uint32)), pin(np.zeros(b.n_chains + 1, np.uint64)),
                              pin(np.zeros(b.n_chains + 1, np.uint32)), pin(np.zeros(b.n_res, np.uint8)), pin(np.zeros(b.n_res, np.float32)),
                              pin(np.zeros((b.n_atoms, 3), np.float32)), pin(np.zeros(max(nt, 1), np.uint8)),
                              pin(np.zeros(b.n_chains, abi.META_DTYPE)), pin(np.zeros(b.n_chains, np.int32)))

    def pinned_blob_batch(b):
        c = abi.encode_bound(b.n_chains, b.n_res, b.n_atoms, len(b.titles), ANCHOR)
        return HostBlobBatch(pin(np.zeros(b.n_chains + 1, np.uint64)), pin(np.zeros(c, np.uint8)), pin(np.zeros(b.n_chains, np.int32)))

    bounds = [round(i * batch.n_chains / E2E_PARTS) for i in range(E2E_PARTS + 1)]
    parts = [batch.select(range(bounds[i], bounds[i + 1])) for i in range(E2E_PARTS)]
    h_in = [pinned_chain_batch(p) for p in parts]
    h_out = [pinned_out_batch(p) for p in parts]
    # blobs travel from the encode thread to the decode thread through a ring of at least two pinned buffers, so the
    # encode of the next sub-batch (or of the next step's batch) never waits for the decode of the previous one
    NB = max(2, E2E_PARTS)
    NB = (NB + E2E_PAIRS - 1) // E2E_PAIRS * E2E_PAIRS  # every pair owns its own ring slots
    big = max(parts, key=lambda p: p.n_atoms)
    h_blob = [pinned_blob_batch(big) for _ in range(NB)]
    engs_enc = [Engine(local_rank, anchor_threshold=ANCHOR) for _ in range(E2E_PAIRS)]
    engs_dec = [Engine(local_rank, anchor_threshold=ANCHOR) for _ in range(E2E_PAIRS)]

    def blob_view(slot, j):  # the ring slot, cut to sub-batch j's chain count
        hb_, n_j = h_blob[slot], parts[j].n_chains
        return HostBlobBatch(hb_.blob_off[: n_j + 1], hb_.bytes, hb_.status[:n_j])

    trace = []  # (leg, item, start, end) of every host call of the last run

    def e2e_run(steps):
        ready = [threading.Semaphore(0) for _ in range(NB)]
        free = [threading.Semaphore(1) for _ in range(NB)]
        errs = []
        trace.clear()

        def enc_loop(pair):
            try:
                for i in range(pair, steps * E2E_PARTS, E2E_PAIRS):
                    slot, j = i % NB, i % E2E_PARTS
                    free[slot].acquire()
                    t_a = time.perf_counter()
                    engs_enc[pair].encode_host(h_in[j], blob_view(slot, j))
                    trace.append(("enc", i, t_a, time.perf_counter()))
                    ready[slot].release()
            except Exception as ex:  # surface in the main thread
                errs.append(ex)
                for sem in ready:
                    sem.release()

        def dec_loop(pair):
            try:
                for i in range(pair, steps * E2E_PARTS, E2E_PAIRS):
                    slot, j = i % NB, i % E2E_PARTS
                    ready[slot].acquire()
                    if errs:
                        return
                    t_a = time.perf_counter()
                    engs_dec[pair].decode_host(blob_view(slot, j), out=h_out[j])
                    trace.append(("dec", i, t_a, time.perf_counter()))
                    free[slot].release()
            except Exception as ex:
                errs.append(ex)
                for sem in free:
                    sem.release()

        th = [threading.Thread(target=f, args=(p,)) for p in range(E2E_PAIRS) for f in (enc_loop, dec_loop)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        if errs:
            raise errs[0]
        return dt

    e2e_run(3)
    barrier()
    count_launches = lambda: sum(e_.launch_count() for e_ in engs_enc + engs_dec)
    l1 = count_launches()
    dt_e = e2e_run(args.steps)
    e2e_launches = count_launches() - l1
    if os.environ.get("FCZ_E2E_TRACE") and rank == 0:
        t_base = min(t[2] for t in trace)
        for leg, i, a, b in sorted(trace, key=lambda t: t[2])[: 6 * E2E_PARTS * 2]:
            print(f"[e2e trace] {leg} item {i:3d}  {1e3 * (a - t_base):8.3f} -> {1e3 * (b - t_base):8.3f} ms  ({1e3 * (b - a):.3f} ms)", file=sys.stderr)
    barrier()
    ms_e = torch.tensor([dt_e * 1e3], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(ms_e, op=dist.ReduceOp.MAX)
    e2e_value = world * n_res * args.steps / (float(ms_e.item()) * 1e-3)
    for p, o in zip(parts, h_out):
        assert not o.status.any() and np.array_equal(o.res_type, p.res_type)
        rt = float(np.sqrt(((o.xyz[:5000] - p.xyz[:5000]) ** 2).sum(1).mean()))
        assert rt < 0.2, rt
    nc = batch.n_chains
    # bytes the host-pointer path moves per step (foldcomp_b200/csrc/fcz_engine.cu: encode_host / decode_host)
    h2d = (12 * n_atoms + 5 * n_res + n_title + 20 * nc + 16 * (nc + 1) + 8 * (nc + 1) + 4 * nc) \
        + (fcz_bytes + 24 * (nc + 1) + 8 * nc)
    d2h = fcz_bytes + (12 * n_atoms + 5 * n_res + n_title + 20 * nc + 4 * nc)

    # ---- what the host link gives these very byte counts as PLAIN pinned copies, both directions at once, every rank of
    # the box at the same time (barrier before, max over ranks): the ceiling the end-to-end figure is judged against
    def pcie_ceiling(reps):
        a_h, a_d = torch.empty(h2d, dtype=torch.uint8).pin_memory(), torch.empty(h2d, dtype=torch.uint8, device=dev)
        b_h, b_d = torch.empty(d2h, dtype=torch.uint8).pin_memory(), torch.empty(d2h, dtype=torch.uint8, device=dev)
        s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

        def go(k):
            for _ in range(k):
                with torch.cuda.stream(s1):
                    a_d.copy_(a_h, non_blocking=True)
                with torch.cuda.stream(s2):
                    b_h.copy_(b_d, non_blocking=True)
            s1.synchronize()
            s2.synchronize()

        go(2)
        barrier()
        t0 = time.perf_counter()
        go(reps)
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return max(h2d, d2h) * reps / float(dt.item()) / 1e9

    ceiling_gbs = pcie_ceiling(max(3, min(args.steps, 10)))

    # the same round trip on ONE engine, one call after the other (no encode/decode overlap), for comparison
    hb_all, hblob_all, hout_all = pinned_chain_batch(batch), pinned_blob_batch(batch), pinned_out_batch(batch)

    def serial_step():
        eng.encode_host(hb_all, hblob_all)
        eng.decode_host(hblob_all, out=hout_all)

    for _ in range(2):
        serial_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        serial_step()
    torch.cuda.synchronize(dev)
    e2e_serial = n_res * args.steps / (time.perf_counter() - t0)
    assert not hout_all.status.any() and np.array_equal(hout_all.res_type, batch.res_type)
    for e_ in engs_enc + engs_dec:
        e_.close()

    # ---- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1:
        threads = os.cpu_count() or 1
        n = cpu_sample_size(batch, threads, 12.0)
        dt, kind, r = cpu_roundtrip(batch, n, threads)
        cpu = {"value": r / dt, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": f"first {n} of the {N_CHAINS} chains, one pass, {threads} OpenMP threads over chains, in-memory "
                         "compress+writeStream+read+decompress per chain (no text I/O); includes the adapter's rebuild of "
                         "vector<AtomCoordinate> from the canonical arrays (oracle/ref_shim.cpp)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+f64", "data": "synthetic", "config": workload_config(world), "roofline": roofline,
            "cpu_baseline": cpu, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(ms_e.item()) / args.steps, "gpu_launches": e2e_launches,
                    "pcie_gbs_each_way": max(h2d, d2h) / (float(ms_e.item()) / args.steps) / 1e6,
                    "pcie_ceiling_gbs": ceiling_gbs,
                    "frac_of_pcie_ceiling": max(h2d, d2h) / (float(ms_e.item()) / args.steps) / 1e6 / ceiling_gbs,
                    "pcie_ceiling_how": "per rank, GB/s each way: the step's H2D and D2H byte counts as plain pinned cudaMemcpyAsync "
                                        "on two streams, both directions and all ranks of the box at once, max over ranks",
                    "host_binding": binding,
                    "how": f"host-pointer C ABI, pinned buffers, {E2E_PARTS} sub-batches per step, {E2E_PAIRS} pair(s) of engines (one "
                           "encoding, one decoding) driven by their own host threads (H2D of encode overlaps D2H of decode); wall clock "
                           "around K steps with a device synchronize on both sides, max over ranks",
                    "serial_one_engine": {"value": world * e2e_serial, "unit": UNIT,
                                          "how": "one engine, fcz_encode_batch then fcz_decode_batch on the whole batch, no overlap"}},
            "gpu_launches": launches, "fcz_bytes_per_step": fcz_bytes, "roundtrip_rmsd_vs_input": dev_rt,
            "warmup_steps_run": n_warm, "ms_per_step_by_rank": [m / args.steps for m in ms_all],
        }
        print(json.dumps(line), file=out, flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


def _quiet_stdout():
    """Libraries (NCCL's version banner) print to fd 1; the contract is ONE JSON line on stdout.  Route fd 1 to
    stderr for the run and return a file on the real stdout for the JSON line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    out = _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kernels-only", action="store_true", help="profiling aid: stop after the device-resident timed region")
    ap.add_argument("--lengths", default="fixed", choices=["fixed", "mixed"],
                    help="fixed = the headline workload (BASELINE.json configs[1]); mixed = the same number of chains with the "
                         "clipped log-normal lengths of configs[4] (AFDB proxy, 50..2000 residues) -- a secondary measurement")
    args = ap.parse_args()
    global LENGTH, WORKLOAD_NOTE
    from foldcomp_b200 import synth
    if args.lengths == "mixed":
        LENGTH = synth.mixed_lengths(np.random.default_rng(synth.SEED), N_CHAINS)
        WORKLOAD_NOTE = " [--lengths mixed: log-normal lengths 50..2000, median 280 -- NOT the headline workload]"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world, out)
    else:
        run_ours(args, rank, world, local_rank, out)


if __name__ == "__main__":
    main()
