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
             coordinates and D2H of blobs and decoded coordinates inside the timed region);
  roofline   the slower of the two hot kernels (k_encode / k_decode): algorithmic bytes per launch
             (SURVEY.md 8d / BASELINE.md 5) over its mean device time (CUDA events recorded by the
             engine around each launch, same timed region), against MEASURED_PEAKS.json hbm_gbs;
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
                                   "Foldcomp::compress+writeStream+read+decompress"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=out, flush=True)


# --------------------------------------------------------------------------------------- clocks


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_id):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_id), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------- ours


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

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # ---- timed region: kernels with HBM-resident inputs
    try:
        gpu_id = str(torch.cuda.get_device_properties(local_rank).uuid)
        if not gpu_id.startswith("GPU-"):
            gpu_id = "GPU-" + gpu_id
    except Exception:
        gpu_id = str(local_rank)
    sampler = ClockSampler(gpu_id)
    eng.set_profiling(True)
    eng.get_profile()
    l0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    prof = eng.get_profile()
    eng.set_profiling(False)
    launches = eng.launch_count() - l0
    clocks = sampler.stop()
    ms_t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_max = float(ms_t.item())
    value = world * n_res * args.steps / (ms_max * 1e-3)

    # sanity on the device result (not a parity test): statuses clean, round trip close to the input
    got = dout.to_host()
    assert not got.status.any() and got.n_atoms == n_atoms
    dev_rt = float(np.sqrt(((got.xyz[: 20000] - batch.xyz[: 20000]) ** 2).sum(1).mean()))
    assert dev_rt < 0.2, dev_rt
    fcz_bytes = int(dblob.blob_off[-1].item())

    # ---- roofline of the dominant kernel
    enc_ms = prof.encode_kernel_ms / max(prof.encode_launches, 1)  # one span per call: all tier launches, fork to join
    dec_ms = prof.decode_kernel_ms / max(prof.decode_launches, 1)
    enc_bytes = 12 * n_atoms + 5 * n_res + fcz_bytes
    dec_bytes = fcz_bytes + 12 * n_atoms + 4 * n_res
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    kname, kms, kbytes = ("k_encode", enc_ms, enc_bytes) if enc_ms >= dec_ms else ("k_decode", dec_ms, dec_bytes)
    achieved = kbytes / (kms * 1e-3) / 1e9 if kms > 0 else 0.0
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of that kernel from the committed ncu --set full capture
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[kname]
        traffic = int(t["dram_bytes_read"]) + int(t["dram_bytes_write"])
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "peak_source": peak_src, "ms_per_launch": kms, "algorithmic_bytes_per_launch": kbytes,
        "kernels": {"k_encode": {"ms_per_launch": enc_ms, "algorithmic_bytes": enc_bytes, "gbs": enc_bytes / max(enc_ms, 1e-9) / 1e6},
                    "k_decode": {"ms_per_launch": dec_ms, "algorithmic_bytes": dec_bytes, "gbs": dec_bytes / max(dec_ms, 1e-9) / 1e6}},
        "round_trip_bytes_per_residue": (enc_bytes + dec_bytes) / n_res,
    }

    # ---- e2e: host-pointer C ABI, pinned buffers, copies inside the timed region
    keep = []

    def pin(a):
        v, t = pinned_like(a)
        keep.append(t)
        return v

    hb = HostChainBatch(pin(batch.res_off), pin(batch.atom_off), pin(batch.title_off), pin(batch.res_type), pin(batch.bfactor),
                        pin(batch.xyz), pin(batch.titles), pin(batch.meta), pin(np.zeros(batch.n_chains, np.int32)))
    hblob = HostBlobBatch(pin(np.zeros(batch.n_chains + 1, np.uint64)), pin(np.zeros(cap, np.uint8)), pin(np.zeros(batch.n_chains, np.int32)))
    hout = HostChainBatch(pin(np.zeros(batch.n_chains + 1, np.uint32)), pin(np.zeros(batch.n_chains + 1, np.uint64)),
                          pin(np.zeros(batch.n_chains + 1, np.uint32)), pin(np.zeros(n_res, np.uint8)), pin(np.zeros(n_res, np.float32)),
                          pin(np.zeros((n_atoms, 3), np.float32)), pin(np.zeros(max(n_title, 1), np.uint8)), pin(np.zeros(batch.n_chains, abi.META_DTYPE)),
                          pin(np.zeros(batch.n_chains, np.int32)))

    def e2e_step():
        eng.encode_host(hb, hblob)
        eng.decode_host(hblob, out=hout)

    for _ in range(3):
        e2e_step()
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        e2e_step()
    ev1.record(stream)
    barrier()
    ms_e = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(ms_e, op=dist.ReduceOp.MAX)
    e2e_value = world * n_res * args.steps / (float(ms_e.item()) * 1e-3)
    assert not hout.status.any() and np.array_equal(hout.res_type, batch.res_type)
    nc = batch.n_chains
    # bytes the host-pointer path moves per step (foldcomp_b200/csrc/fcz_engine.cu: encode_host / decode_host)
    h2d = (12 * n_atoms + 5 * n_res + n_title + 20 * nc + 16 * (nc + 1) + 8 * (nc + 1) + 4 * nc) \
        + (fcz_bytes + 24 * (nc + 1) + 8 * nc)
    d2h = fcz_bytes + (12 * n_atoms + 5 * n_res + n_title + 20 * nc + 4 * nc)

    # ---- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1:
        threads = os.cpu_count() or 1
        n = cpu_sample_size(batch, threads, 12.0)
        dt, kind, r = cpu_roundtrip(batch, n, threads)
        cpu = {"value": r / dt, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": f"first {n} of the {N_CHAINS} chains, one pass, {threads} OpenMP threads over chains, in-memory "
                         "compress+writeStream+read+decompress per chain (no text I/O)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+f64", "data": "synthetic", "config": workload_config(world), "roofline": roofline,
            "cpu_baseline": cpu, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(ms_e.item()) / args.steps},
            "gpu_launches": launches, "fcz_bytes_per_step": fcz_bytes, "roundtrip_rmsd_vs_input": dev_rt,
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
