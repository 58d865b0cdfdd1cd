mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/config4.py --chains-per-gpu 250000 --out /tmp/merged_db > gpurun_out/config4_n2.json 2> gpurun_out/config4_n2.err
cat gpurun_out/bench_n2.json | cut -c1-300; tail -2 gpurun_out/bench_n2.err; cat gpurun_out/bench_ref_n2.json | cut -c1-200; cat gpurun_out/config4_n2.json; tail -3 gpurun_out/config4_n2.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/bench_n2.json"))
print("N=2 value %.3f G e2e %.1f M" % (j["value"] / 1e9, j["e2e"]["value"] / 1e6))
PY
