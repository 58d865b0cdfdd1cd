mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 compute-sanitizer --tool memcheck --target-processes application-only --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_text.py -x -q -k "edge or golden or unk or titles or extract or fused or terminated or failed_chains" > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck exit: $?" >> gpurun_out/sanitize_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --target-processes application-only --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck exit: $?" >> gpurun_out/sanitize_racecheck.log
grep -E "ERROR SUMMARY|exit:|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitize_memcheck.log gpurun_out/sanitize_racecheck.log | tail -10
grep -E "Invalid|hazard" gpurun_out/sanitize_memcheck.log gpurun_out/sanitize_racecheck.log | sort | uniq -c | sort -rn | head -10
