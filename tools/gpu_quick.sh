#!/bin/bash
# quick GPU iteration: corr+splat parity tests then a short bench (no CPU baseline)
mkdir -p gpurun_out
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
echo "== bench" ; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench.txt
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench.txt') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1))
    for k,v in d['breakdown'].items(): print(f"  {k:16s} {v['ms_per_call']*1000:9.1f} us  {v['GBps']:8.1f} GB/s  {v['frac_of_peak']:.3f}")
PY
