mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_text.py tests/test_gpu_parity.py::test_edge_cases -x -q -rA 2>&1 | tail -30 > gpurun_out/pytest_gpu_text.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu_text.log
timeout 300 python tools/bench_text.py > gpurun_out/bench_text.json 2> gpurun_out/bench_text.err
N_CHAINS=10000 STEPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_pdb_emit|k_pdb_plan' -s 6 -c 2 -o gpurun_out/prof_text -f python tools/bench_text.py > gpurun_out/ncu_text.log 2>&1
tail -12 gpurun_out/pytest_gpu_text.log; cat gpurun_out/bench_text.json; tail -3 gpurun_out/bench_text.err
