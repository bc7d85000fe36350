#!/bin/bash
# ncu launch list of one bench step + full capture of selected kernels
mkdir -p gpurun_out
KREGEX=${1:-splat_scatter}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); 
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    a=agg.setdefault(r[ki],[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:25]:
    print(f"{a[1]/1e3:10.1f} us total {a[0]:4d} launches {a[1]/a[0]/1e3:9.1f} us avg {100*a[1]/tot:5.1f}%  {k[:90]}")
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 2 -c 2 -o gpurun_out/prof_$KREGEX -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
