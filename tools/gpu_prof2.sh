#!/bin/bash
# launch list + ncu --set full of the decode kernels (whole-batch phases unless FCZ_DEC_SUB_RESIDUES is set)
mkdir -p gpurun_out
export FCZ_DEC_SUB_RESIDUES=${FCZ_DEC_SUB_RESIDUES:-4000000}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
for k in ${KERNELS:-k_dec_front k_dec_back k_dec_stitch_soa}; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o gpurun_out/prof_$k -f \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full_$k.log 2>&1
done
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    agg.setdefault(r[ki].split('(')[0],[]).append(v)
for k,v in agg.items(): print(f"{k:40s} n={len(v):3d} mean={sum(v)/len(v)/1e3:9.1f} us  last={v[-1]/1e3:9.1f} us")
PY
