#!/bin/bash
# ncu --set full capture of the two hot kernels (one launch each, after warm-up) -> gpurun_out/prof_*.ncu-rep
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_encode -s 4 -c 1 -o gpurun_out/prof_encode -f \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full_encode.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_decode -s 4 -c 1 -o gpurun_out/prof_decode -f \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full_decode.log 2>&1
ls -la gpurun_out/*.ncu-rep
