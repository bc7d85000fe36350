#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (splat)"; timeout 900 python -m pytest tests/test_gpu_splat.py -m gpu -x -q 2>&1 | tail -8
for v in 1 2; do
  echo "== bench FLDR_SPLAT_COLS=$v"
  FLDR_SPLAT_COLS=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_v$v.txt
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_v$v.txt').read())
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3))
for k,v in d['breakdown'].items():
    if k.startswith('splat'): print(f"  {k:16s} {v['ms_per_call']*1000:9.1f} us  {v['GBps']:8.1f} GB/s  {v['frac_of_peak']:.3f}")
PY
done
for v in 1 2; do
  FLDR_SPLAT_COLS=$v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:splat -c 60 --csv --log-file gpurun_out/l_v$v.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/l_v$v.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
print("variant $v")
for r in rows[-24:]:
    print(f"  {float(r[vi].replace(',',''))/1e3:9.1f} us  grid {r[gi]:>16s}  {r[ki][:60]}")
PY
done
