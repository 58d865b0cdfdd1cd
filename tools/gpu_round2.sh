#!/bin/bash
# Round-2 GPU-box visit.  Stages selected by env (default: tests + bench + A/B); everything bounded by `timeout`.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
if [ "${TESTS:-1}" = "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -rA ${PYTEST_ARGS:-} 2>&1 | tail -120 > gpurun_out/pytest_gpu.log
  echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
  echo "smoke exit: $?" >> gpurun_out/smoke.log
fi
if [ "${BENCH:-1}" = "1" ]; then
  timeout 600 python bench.py --steps ${STEPS:-20} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit: $?" >> gpurun_out/bench.err
fi
if [ -n "${AB:-}" ]; then
  VARIANTS="$AB" timeout 1200 python tools/kernel_ab.py > gpurun_out/ab.jsonl 2> gpurun_out/ab.err
fi
if [ -n "${AB_MIXED:-}" ]; then
  LENGTHS=mixed VARIANTS="$AB_MIXED" timeout 1200 python tools/kernel_ab.py > gpurun_out/ab_mixed.jsonl 2> gpurun_out/ab_mixed.err
fi
if [ "${REFARM:-0}" = "1" ]; then
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
fi
if [ "${NCU:-0}" = "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-k_encode|k_dec_front|k_dec_stitch_t|k_dec_back}" -s ${NCU_SKIP:-12} -c ${NCU_COUNT:-4} \
      -o gpurun_out/prof_step -f python bench.py --kernels-only --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
fi
if [ -n "${EXTRA:-}" ]; then
  bash -c "$EXTRA" > gpurun_out/extra.log 2>&1
fi
tail -8 gpurun_out/pytest_gpu.log 2>/dev/null; tail -2 gpurun_out/smoke.log 2>/dev/null; cat gpurun_out/bench.json 2>/dev/null | cut -c1-1500; tail -3 gpurun_out/bench.err 2>/dev/null
cat gpurun_out/ab.jsonl 2>/dev/null | cut -c1-600; tail -3 gpurun_out/ab.err 2>/dev/null
exit 0
