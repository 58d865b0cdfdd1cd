#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), PCIe probe, ncu launch list of the bench command and one
# ncu --set full capture of one step's four hot kernels.  Everything bounded by `timeout`.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
if [ "${TESTS:-1}" = "1" ]; then
timeout 1200 python -m pytest tests -m gpu -x -q -rA -s 2>&1 | tail -80 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke.log
fi
timeout 120 python tools/pcie_probe.py > gpurun_out/pcie.json 2> gpurun_out/pcie.err
timeout 600 python bench.py --steps ${STEPS:-20} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit: $?" >> gpurun_out/bench.err
if [ "${REFARM:-1}" = "1" ]; then
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
fi
if [ "${NCU:-1}" = "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
  # one step's hot kernels after 3 warm-up steps: 4 matching launches per step (one length tier)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_encode|k_dec_front|k_dec_stitch_t|k_dec_back' -s 12 -c 4 \
      -o gpurun_out/prof_step -f python bench.py --kernels-only --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
fi
if [ "${TEXT:-1}" = "1" ]; then
  timeout 300 python tools/bench_text.py > gpurun_out/bench_text.json 2> gpurun_out/bench_text.err
  if [ "${NCU:-1}" = "1" ]; then
    N_CHAINS=10000 STEPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_pdb_emit|k_pdb_plan' -s 6 -c 2 \
        -o gpurun_out/prof_text -f python tools/bench_text.py > gpurun_out/ncu_text.log 2>&1
  fi
fi
if [ "${MIXED:-0}" = "1" ]; then
  timeout 600 python bench.py --lengths mixed --steps 10 --warmup 3 > gpurun_out/bench_mixed.json 2> gpurun_out/bench_mixed.err
fi
tail -6 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/pcie.json; python tools/summ.py; cat gpurun_out/bench_text.json; tail -2 gpurun_out/bench_text.err; tail -3 gpurun_out/bench.err
