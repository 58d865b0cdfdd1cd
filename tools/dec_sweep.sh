#!/bin/bash
# decode sub-batch size sweep (residues per sub-batch); prints enc/dec kernel ms per setting
mkdir -p gpurun_out
for r in ${SWEEP:-130000 207200 262000 414400 525000 1050000 4000000}; do
  FCZ_DEC_SUB_RESIDUES=$r timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('$r', round(j['value']/1e9,3), round(j['ms_per_step'],4), {k:round(v['ms_per_launch'],4) for k,v in j['roofline']['kernels'].items()}, round(j['e2e']['value']/1e6,1))" 
done | tee gpurun_out/dec_sweep.txt
