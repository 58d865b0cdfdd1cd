#!/usr/bin/env python3
"""Print the bench line's headline numbers and the ncu launch list (per-kernel max/min) from gpurun_out/."""
import json, csv, collections, os
try:
    j = json.load(open("gpurun_out/bench.json"))
    print("value %.3f G res/s  step %.4f ms  kernels %s  e2e %.1f M res/s" % (j["value"] / 1e9, j["ms_per_step"],
          {k: round(v["ms_per_launch"] * v.get("launches_per_step", 1), 4) for k, v in j["roofline"]["kernels"].items()}, j["e2e"]["value"] / 1e6))
    print("roofline", j["roofline"]["kernel"], round(j["roofline"]["frac"], 4), "e2e serial %.1f M" % (j["e2e"].get("serial_one_engine", {}).get("value", 0) / 1e6),
          "pcie %.1f GB/s" % j["e2e"].get("pcie_gbs_each_way", 0), "cpu", j.get("cpu_baseline"))
except Exception as ex:
    print("bench.json:", ex)
if os.path.exists("gpurun_out/launches.csv"):
    rows = [r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r) > 5]
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        agg.setdefault(r[ki].split("(")[0], []).append(v)
    for k, v in agg.items():
        if k.startswith("k_"):
            print(f"{k[:40]:40s} n={len(v):3d} max={max(v)/1e3:9.1f} us  min={min(v)/1e3:9.1f} us")
