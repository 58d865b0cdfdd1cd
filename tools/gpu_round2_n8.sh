#!/bin/bash
# Round-2 visit of an N-GPU box (default 8): the bench at N ranks with and without per-rank core binding (end-to-end
# scaling + in-run PCIe ceiling), BASELINE.json configs[3] at full size (N x 250 000 distinct chains, merged database).
N=${N:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
(lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)"; nproc; free -g | head -2; for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -q 0x10de $d/vendor 2>/dev/null; then echo "$d numa $(cat $d/numa_node)"; fi; done) >> gpurun_out/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench exit $?" >> gpurun_out/bench_n$N.err
[ "${NOBIND:-0}" = "1" ] && FCZ_BIND=0 timeout 600 $TR bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_n${N}_nobind.json 2> gpurun_out/bench_n${N}_nobind.err
timeout 900 $TR tests/run_config4.py --chains-per-gpu ${CHAINS:-250000} --out /tmp/fcz_merged_db > gpurun_out/config4_n$N.json 2> gpurun_out/config4_n$N.err
echo "config4 exit $?" >> gpurun_out/config4_n$N.err
ls -la /tmp/fcz_merged_db* >> gpurun_out/config4_n$N.err 2>&1
python - <<PY
import json
for f in ("gpurun_out/bench_n$N.json", "gpurun_out/bench_n${N}_nobind.json"):
    try:
        d = json.load(open(f)); e = d["e2e"]
        print(f, "value", d["value"], "e2e", e["value"], "link", e["pcie_gbs_each_way"], "ceiling", e["pcie_ceiling_gbs"], "frac", e["frac_of_pcie_ceiling"], e.get("host_binding"))
    except Exception as ex:
        print(f, "ERR", ex)
PY
cat gpurun_out/config4_n$N.json; tail -3 gpurun_out/config4_n$N.err; tail -2 gpurun_out/bench_n$N.err; cat gpurun_out/topo.txt | tail -25
