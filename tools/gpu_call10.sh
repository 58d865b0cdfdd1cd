mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "host_api or mixed_lengths or edge or long_chains or titles or golden" 2>&1 | tail -3
timeout 300 python tools/overlap_probe.py
