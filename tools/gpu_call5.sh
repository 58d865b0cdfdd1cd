mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py::test_terminated_blobs_are_a_db_slab tests/test_gpu_parity.py::test_edge_cases -x -q 2>&1 | tail -5 > gpurun_out/pytest_sub.log
for p in 1 2; do
FCZ_E2E_PARTS=$p timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_p$p.json 2> gpurun_out/bench_p$p.err
done
timeout 600 python tools/config4.py --chains-per-gpu 100000 --out /tmp/merged_db > gpurun_out/config4_n1.json 2> gpurun_out/config4_n1.err
cat gpurun_out/pytest_sub.log; cat gpurun_out/config4_n1.json; tail -3 gpurun_out/config4_n1.err
python - <<'PY'
import json
for p in (1, 2):
    try:
        j = json.load(open(f"gpurun_out/bench_p{p}.json"))
        print(p, "value %.3f G" % (j["value"] / 1e9), "e2e %.1f M" % (j["e2e"]["value"] / 1e6), "serial %.1f M" % (j["e2e"]["serial_one_engine"]["value"] / 1e6), "pcie %.1f" % j["e2e"]["pcie_gbs_each_way"], "cpu %.2f M" % (j["cpu_baseline"]["value"] / 1e6))
    except Exception as ex:
        print(p, ex, open(f"gpurun_out/bench_p{p}.err").read()[-800:])
PY
