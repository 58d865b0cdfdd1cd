mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
# memcheck: edge cases of the codec and text kernels (round 1) + the round-2 kernels (k_check, k_parse_*, k_raw_angles)
timeout 1200 compute-sanitizer --tool memcheck --target-processes application-only --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_text.py tests/test_check.py tests/test_gpu_parse.py tests/test_getdata.py -q -m gpu -k "edge or golden or unk or titles or extract or fused or terminated or failed_chains or degenerate or check or device_parser or one_call or backbone_angles" > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck exit: $?" >> gpurun_out/sanitize_memcheck.log
# racecheck: one small encode + decode (smoke) and the parser's block-level passes
timeout 600 compute-sanitizer --tool racecheck --target-processes application-only --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck exit: $?" >> gpurun_out/sanitize_racecheck.log
timeout 900 compute-sanitizer --tool racecheck --target-processes application-only --error-exitcode 9 python -m pytest tests/test_gpu_parse.py -q -m gpu -k "device_parser" > gpurun_out/sanitize_racecheck_parse.log 2>&1
echo "racecheck (parser) exit: $?" >> gpurun_out/sanitize_racecheck_parse.log
grep -E "ERROR SUMMARY|exit:|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitize_memcheck.log gpurun_out/sanitize_racecheck.log gpurun_out/sanitize_racecheck_parse.log | tail -14
grep -E "Invalid|hazard" gpurun_out/sanitize_memcheck.log gpurun_out/sanitize_racecheck.log gpurun_out/sanitize_racecheck_parse.log | sort | uniq -c | sort -rn | head -10
