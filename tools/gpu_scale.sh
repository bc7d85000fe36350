#!/bin/bash
# bench.py at N GPUs (argument), short legs only; the line lands in gpurun_out/scale_n$N.json
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -12 > gpurun_out/topo_n$N.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-fldrnet --no-ref-gpu --no-cpu-baseline > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
tail -c 300 gpurun_out/scale_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/scale_n$N.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "n_gpus", "ms_per_step")}, d["e2e"], d.get("e2e_frames", {}).get("value"), d.get("strong_scaling", {}).get("value"))
PY
