#!/bin/bash
# first GPU call: parity tests, design microbenchmarks, a short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
ls /root/reference > gpurun_out/ref_exists.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== microbench" ; timeout 300 ./tools/microbench 2>&1 | tee gpurun_out/microbench.txt
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.txt
echo "== bench" ; timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench.txt
