mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "host_api or mixed_lengths or edge or long_chains or titles" 2>&1 | tail -5 > gpurun_out/pytest_sub.log
for p in 1 2 4; do
FCZ_E2E_PARTS=$p timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_p$p.json 2> gpurun_out/bench_p$p.err
done
cat gpurun_out/pytest_sub.log
python - <<'PY'
import json
for p in (1, 2, 4):
    try:
        j = json.load(open(f"gpurun_out/bench_p{p}.json"))
        print(p, "value %.3f G" % (j["value"] / 1e9), "e2e %.1f M" % (j["e2e"]["value"] / 1e6), "serial %.1f M" % (j["e2e"]["serial_one_engine"]["value"] / 1e6), "pcie %.1f" % j["e2e"]["pcie_gbs_each_way"], "cpu %.2f M" % (j["cpu_baseline"]["value"] / 1e6))
    except Exception as ex:
        print(p, ex, open(f"gpurun_out/bench_p{p}.err").read()[-800:])
PY
