#!/bin/bash
mkdir -p gpurun_out
export FLDR_SPLAT_STREAM=1
echo "== pytest streaming"; timeout 600 python -m pytest tests/test_gpu_splat.py -m gpu -x -q -k "stream or overflow or 4k" 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench.txt; python tools/show_bench.py | head -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:splat_stream -s 0 -c 1 -o gpurun_out/prof_stream -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
