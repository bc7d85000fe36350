#!/bin/bash
# usage: gpu_iter2.sh <pytest target or "none"> [ENV=VAL ...]   -> tests + bench + per-kernel ncu list, with env applied
mkdir -p gpurun_out
T=$1; shift
for kv in "$@"; do export "$kv"; done
if [ "$T" != "none" ]; then echo "== pytest"; timeout 900 python -m pytest $T -m gpu -x -q 2>&1 | tail -6; fi
echo "== bench ($*)"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench.txt
python - <<PY
import json
d=json.loads(open('gpurun_out/bench.txt').read())
print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1))
for k,v in d['breakdown'].items(): print(f"  {k:16s} {v['ms_per_call']*1000:9.1f} us  {v['GBps']:8.1f} GB/s  {v['frac_of_peak']:.3f}")
PY
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"splat" -c 60 --csv --log-file gpurun_out/l.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/l.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size'); mi=hdr.index('Metric Name'); ii=hdr.index('ID')
d=collections.OrderedDict()
for r in rows[1:]:
    e=d.setdefault(r[ii],{'k':r[ki],'g':r[gi]}); e[r[mi]]=float(r[vi].replace(',',''))
seen=0
for e in d.values():
    if e.get('gpu__time_duration.sum',0) < 12e3: continue
    seen+=1
    if seen>10: break
    print(f"  {e.get('gpu__time_duration.sum',0)/1e3:8.1f} us inst {e.get('smsp__inst_executed.sum',0)/1e6:7.2f}M dramR {e.get('dram__bytes_read.sum',0)/1e6:7.1f}MB W {e.get('dram__bytes_write.sum',0)/1e6:7.1f}MB grid {e['g']:>15s} {e['k'][:48]}")
PY
