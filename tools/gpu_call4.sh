mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_text.py tests/test_gpu_db.py tests/test_gpu_parity.py::test_edge_cases tests/test_gpu_parity.py::test_roundtrip_uniform_350_host_api tests/test_gpu_parity.py::test_mixed_lengths_anchor_sweep -x -q 2>&1 | tail -5 > gpurun_out/pytest_sub.log
timeout 300 python tools/bench_text.py > gpurun_out/bench_text.json 2> gpurun_out/bench_text.err
for p in 4 8; do
FCZ_E2E_PARTS=$p timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_p$p.json 2> gpurun_out/bench_p$p.err
done
N_CHAINS=10000 STEPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_pdb_emit' -s 3 -c 1 -o gpurun_out/prof_text -f python tools/bench_text.py > gpurun_out/ncu_text.log 2>&1
cat gpurun_out/pytest_sub.log; cat gpurun_out/bench_text.json
python - <<'PY'
import json
for p in (4, 8):
    try:
        j = json.load(open(f"gpurun_out/bench_p{p}.json"))
        print(p, "value %.3f G" % (j["value"] / 1e9), "e2e %.1f M" % (j["e2e"]["value"] / 1e6), "serial %.1f M" % (j["e2e"]["serial_one_engine"]["value"] / 1e6), "pcie %.1f" % j["e2e"]["pcie_gbs_each_way"], "cpu %.2f M" % (j["cpu_baseline"]["value"] / 1e6))
    except Exception as ex:
        print(p, ex)
PY
