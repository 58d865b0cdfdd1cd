mkdir -p gpurun_out
for p in 1 4; do
FCZ_E2E_TRACE=1 FCZ_E2E_PARTS=$p timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_p$p.json 2> gpurun_out/bench_p$p.err
grep "e2e trace" gpurun_out/bench_p$p.err | head -28
done
