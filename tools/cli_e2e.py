#!/usr/bin/env python3
"""CLI end-to-end baseline: the reference's own CLI
(integration/_build/foldcomp_ref: unmodified sources) with `-t <all cores> --db` next to this repo's batched CLI (fcz_cli
compress-db / decompress-db) on the SAME foldcomp databases of N synthetic 350-residue chains -- PDB text in, FCZ out and
back.  Everything from and to files; the text database is produced once by the GPU emitter (that run is the timed
`ours decompress`).  Prints one JSON line.

    python tools/cli_e2e.py [--chains 20000] [--dir /tmp/cli_e2e]"""
import argparse, json, os, shutil, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=20000)
ap.add_argument("--dir", default="/tmp/cli_e2e")
ap.add_argument("--skip-ours", action="store_true", help="reference legs only (no GPU needed; needs a PDB-text db made earlier)")
args = ap.parse_args()
REF = os.path.join(ROOT, "integration", "_build", "foldcomp_ref")
OURS = os.path.join(ROOT, "foldcomp_b200", "csrc", "fcz_cli")
os.makedirs(args.dir, exist_ok=True)
P = lambda n: os.path.join(args.dir, n)
threads = os.cpu_count() or 1
L = 350


def timed(cmd):
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise SystemExit(f"{cmd}: rc {r.returncode}\n{r.stdout[-500:]}\n{r.stderr[-500:]}")
    return dt, r


def db_bytes(path):
    return sum(os.path.getsize(path + s) for s in ("", ".index", ".lookup") if os.path.exists(path + s))


out = {"chains": args.chains, "residues": args.chains * L, "host_threads": threads}
if not args.skip_ours:
    import torch
    from foldcomp_b200 import abi, synth, synth_device
    from foldcomp_b200.engine import Engine
    import dbutil

    # 1. the FCZ database: chains generated on the device, encoded by the engine, written with the reference's layout
    dev = torch.device("cuda:0")
    g = synth_device.generate_device(args.chains, L, synth.SEED, 0, dev)
    hb = abi.HostChainBatch(res_off=g["res_off"].cpu().numpy().astype(np.uint32), atom_off=g["atom_off"].cpu().numpy().astype(np.uint64),
                            title_off=g["title_off"].cpu().numpy().astype(np.uint32), res_type=g["res_type"].cpu().numpy(),
                            bfactor=g["bfactor"].cpu().numpy(), xyz=g["xyz"].cpu().numpy(), titles=g["titles"].cpu().numpy(),
                            meta=g["meta"].cpu().numpy().view(abi.META_DTYPE).reshape(-1))
    with Engine(0) as eng:
        blobs = eng.encode_host(hb)
    dbutil.write_db(P("fcz_db"), [(c, "syn_%07d" % c, blobs.blob(c)) for c in range(args.chains)])
    del g, hb
    # 2. ours: FCZ db -> PDB-text db (GPU decode + text emitter), then PDB-text db -> FCZ db (GPU parser + GPU encode)
    dt, r = timed([OURS, "decompress-db", P("fcz_db"), P("pdb_db")])
    out["ours_decompress_db"] = {"seconds": dt, "residues_per_s": args.chains * L / dt, "stderr": r.stderr.strip()[-200:]}
    os.environ["FCZ_DB_WRITE"] = "mmap"  # A/B of the text writer: stores through a shared mapping instead of pwrite
    dt, r = timed([OURS, "decompress-db", P("fcz_db"), P("pdb_db_pwrite")])
    del os.environ["FCZ_DB_WRITE"]
    out["ours_decompress_db_mmap"] = {"seconds": dt, "residues_per_s": args.chains * L / dt, "stderr": r.stderr.strip()[-200:]}
    for sfx in ("", ".index", ".lookup", ".dbtype"):
        os.remove(P("pdb_db_pwrite") + sfx)
    dt, r = timed([OURS, "compress-db", P("pdb_db"), P("fcz_db_ours")])
    out["ours_compress_db"] = {"seconds": dt, "residues_per_s": args.chains * L / dt, "stderr": r.stderr.strip()[-200:]}
    os.environ["FCZ_HOST_PARSER"] = "1"  # A/B: the per-entry host parser (OpenMP) instead of the GPU parser
    dt, r = timed([OURS, "compress-db", P("pdb_db"), P("fcz_db_ours_hostparse")])
    del os.environ["FCZ_HOST_PARSER"]
    out["ours_compress_db_host_parser"] = {"seconds": dt, "residues_per_s": args.chains * L / dt, "stderr": r.stderr.strip()[-200:]}
    out["pdb_text_db_bytes"] = db_bytes(P("pdb_db"))
    out["fcz_db_bytes"] = db_bytes(P("fcz_db"))
# 3. the reference CLI on the same two databases, all host threads
for d in ("pdb_db_ref", "fcz_db_ref"):
    for s in ("", ".index", ".lookup", ".dbtype"):
        if os.path.exists(P(d) + s):
            os.remove(P(d) + s)
dt, r = timed([REF, "decompress", "-t", str(threads), "-y", "--db", P("fcz_db"), P("pdb_db_ref")])
out["reference_decompress_db"] = {"seconds": dt, "residues_per_s": args.chains * L / dt}
dt, r = timed([REF, "compress", "-t", str(threads), "-y", "--db", P("pdb_db"), P("fcz_db_ref")])
out["reference_compress_db"] = {"seconds": dt, "residues_per_s": args.chains * L / dt}
if not args.skip_ours:
    # same answers: the FCZ databases hold the same blobs (padding bytes masked), the text databases the same entries
    import helpers as H

    stem = lambda n: n.split(".")[0]
    a = {stem(n): b for _, n, b in dbutil.read_db(P("fcz_db_ours"))}
    b = {stem(n): x for _, n, x in dbutil.read_db(P("fcz_db_ref"))}
    same = sum(1 for k in a if k in b and H.masked(a[k]) == H.masked(b[k]))
    out["compress_outputs_identical"] = f"{same} of {len(a)} entries (reference has {len(b)})"
    out["speedup_compress"] = out["ours_compress_db"]["residues_per_s"] / out["reference_compress_db"]["residues_per_s"]
    out["speedup_decompress"] = out["ours_decompress_db"]["residues_per_s"] / out["reference_decompress_db"]["residues_per_s"]
print(json.dumps(out))
shutil.rmtree(args.dir, ignore_errors=True)
