#!/bin/bash
# warp row: parity tests, then timing of bwarp / splat_metric at the 4K image shape against the torch path
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_warp.py -m gpu -x -q 2>&1 | tail -8
timeout 300 python tools/warp_probe.py 2>&1 | tee gpurun_out/warp_probe.txt
