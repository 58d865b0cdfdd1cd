mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -rA 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_text.py > gpurun_out/bench_text.json 2> gpurun_out/bench_text.err
N_CHAINS=10000 STEPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_pdb_emit|k_pdb_plan' -s 6 -c 2 -o gpurun_out/prof_text -f python tools/bench_text.py > gpurun_out/ncu_text.log 2>&1
grep -E "passed|failed|error|Error" gpurun_out/pytest_gpu.log | tail -8; cat gpurun_out/bench_text.json; tail -3 gpurun_out/bench_text.err
